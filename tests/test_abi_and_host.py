"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol declared in
include/smm_b200.h, fails loudly without a GPU (no CPU fallback), and the host mirror builds the
config the way MAlgoBGP does (AlgoBGP.jl:505-538)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "smm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(smm):
    L = smm.lib()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/smm_b200.h but not exported"
    assert set(names) == set(smm.EXPORTS)
    assert L.smm_abi_version() == 1


def test_struct_layout_matches_header(smm):
    """sizeof(smm_bgp_config) as ctypes sees it == as the C compiler sees it"""
    import subprocess, tempfile
    from smm_jl_b200._abi import smm_bgp_config, smm_trace_view, smm_counters
    code = ('#include <stdio.h>\n#include <stddef.h>\n#include "smm_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n",'
            'sizeof(smm_bgp_config),sizeof(smm_trace_view),sizeof(smm_counters),offsetof(smm_bgp_config,nccl_id),'
            'offsetof(smm_bgp_config,seed_algo));return 0;}')
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(code)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")], check=True)
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split()
    assert [int(x) for x in out] == [ctypes.sizeof(smm_bgp_config), ctypes.sizeof(smm_trace_view), ctypes.sizeof(smm_counters),
                                     smm_bgp_config.nccl_id.offset, smm_bgp_config.seed_algo.offset]


def test_no_cpu_fallback(smm):
    """without a CUDA device the product fails loudly instead of computing on the host"""
    from smm_jl_b200 import configs
    if smm.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(smm.SMMError) as e:
        smm.BGPHandle(configs.c1_serial_normal(5))
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    with pytest.raises(smm.SMMError):
        smm.debug_normals(1, 0, 0, 0, 8)


def test_argument_errors_before_any_device_work(smm):
    from smm_jl_b200 import configs
    from smm_jl_b200._abi import SMM_E_ARG, SMM_E_UNSUPPORTED_SHAPE
    bad = configs.mvnormal(4, 4, n_params=5, batch_size=2)
    with pytest.raises(smm.SMMError) as e:
        smm.BGPHandle(bad)
    assert e.value.code == SMM_E_UNSUPPORTED_SHAPE
    bad = configs.c1_serial_normal(5)
    bad.ub = [3.0, -30.0]
    with pytest.raises(smm.SMMError) as e:
        smm.BGPHandle(bad)
    assert e.value.code == SMM_E_ARG
    bad = configs.mvnormal(6, 4, world_size=4, rank=1)
    with pytest.raises(smm.SMMError) as e:
        smm.BGPHandle(bad)
    assert e.value.code == SMM_E_UNSUPPORTED_SHAPE


def test_host_uacc_stream_matches_oracle(smm, oracle):
    got = smm.acc_uniforms(12, 2, 1, 50)
    want = np.array([oracle.acc_uniform(12, 2, it) for it in range(1, 51)])
    np.testing.assert_array_equal(got, want)


def test_product_does_not_import_the_oracle():
    """the product path must not route through oracle/ (or any CPU implementation)"""
    pkg = os.path.join(ROOT, "smm_jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in txt and "oracle_np" not in txt and "libsmm_oracle" not in txt, f
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


def test_mprob_builders_and_config():
    """test/test_MProb.jl:29-47 + MAlgoBGP constructor defaults (AlgoBGP.jl:505-538)"""
    from smm_jl_b200 import api
    m = api.MProb()
    api.addSampledParam(m, {"p1": [0.2, -3, 3], "p2": [-0.2, -2, 2]})
    api.addMoment(m, {"mu1": {"value": -1.0, "weight": 1.0}, "mu2": {"value": 1.0, "weight": 1.0}})
    api.addEvalFunc(m, api.objfunc_norm)
    assert api.ps2s_names(m) == ["p1", "p2"] and api.ms_names(m) == ["mu1", "mu2"] and api.ps_names(m) == ["p1", "p2"]
    algo = api.MAlgoBGP(m)                       # default opts: 3 chains, i == 0 (test_algoBGP.jl:14-28)
    assert algo.i == 0 and algo["N"] == 3
    cfg = algo._cfg
    np.testing.assert_allclose(cfg.sigma0, 0.05 * np.array([1.0, 1.5, 2.0]))   # range(1, maxtemp=2, length=3)
    assert cfg.batch_size == 2 and cfg.smpl_iters == 1000 and cfg.sigma_update_steps == 10
    opts = {"N": 3, "maxiter": 20, "maxtemp": 5, "smpl_iters": 1000, "parallel": False, "min_improve": [0.0] * 3,
            "acc_tuners": [20, 2, 1.0], "coverage": 0.02, "animate": False}
    cfg = api.MAlgoBGP(m, opts)._cfg
    np.testing.assert_allclose(cfg.sigma0, 0.05 * np.array([1.0, 3.0, 5.0]))
    np.testing.assert_allclose(cfg.acc_tuner, [20, 2, 1])
    np.testing.assert_allclose(api.mapto_ab(api.mapto_01([0.2, -0.2], cfg.lb, cfg.ub), cfg.lb, cfg.ub), [0.2, -0.2])
    # an arbitrary host function cannot run on the device: no CPU fallback
    api.addEvalFunc(m, lambda ev: ev)
    with pytest.raises(NotImplementedError):
        api.MAlgoBGP(m, opts)
    # non-sampled parameters break proposal() upstream; rejected here
    m2 = api.MProb()
    api.addParam(m2, "fixed", 1.0)
    api.addSampledParam(m2, "p1", 0.2, -3, 3)
    api.addMoment(m2, "mu1", 0.0)
    api.addEvalFunc(m2, api.objfunc_norm)
    with pytest.raises(ValueError):
        api.MAlgoBGP(m2, {"N": 1, "maxiter": 3})


def test_eval_accessors():
    """test/test_Eval.jl:30-64"""
    from smm_jl_b200 import api
    m = api.MProb()
    api.addSampledParam(m, "a", 0.3, -1, 1)
    api.addSampledParam(m, "b", -0.9, -2, 2)
    api.addMoment(m, "mu1", 0.0, 0.5)
    api.addMoment(m, "mu2", 1.0, 2.0)
    ev = api.Eval(m)
    assert ev.status == -1 and ev.value == -1.0 and ev.prob == 0.0 and not ev.accepted     # Eval.jl:82-106
    assert api.param(ev, "a") == 0.3 and list(api.param(ev)) == [0.3, -0.9] and api.paramd(ev) == {"a": 0.3, "b": -0.9}
    assert api.dataMoment(ev, "mu2") == 1.0 and list(api.dataMomentW(ev)) == [0.5, 2.0]
    api.setMoments(ev, {"mu1": 0.1, "mu2": 0.9})
    api.setMoments(ev, "mu1", 0.2)
    api.setValue(ev, 3.5)
    assert ev.simMoments == {"mu1": 0.2, "mu2": 0.9} and ev.value == 3.5
    ev2 = api.Eval(m, {"a": 0.0, "b": 0.0})
    assert ev2 != ev and api.Eval(m) == api.Eval(m)


def test_temperature_ladder_and_shards():
    from smm_jl_b200.configs import temperature_ladder
    from smm_jl_b200.dist import shard_chains, owner_of, local_index
    np.testing.assert_allclose(temperature_ladder(3, 5), [1, 3, 5])
    np.testing.assert_allclose(temperature_ladder(1, 5), [1])
    assert shard_chains(1024, 8, 7)[:3].tolist() == [7, 15, 23] and len(shard_chains(1024, 8, 0)) == 128
    assert owner_of(129, 1024, 8) == 1 and local_index(129, 8) == 16
    with pytest.raises(ValueError):
        shard_chains(10, 4, 0)


def _toy_problem(api, n_params=2, objective=None):
    m = api.MProb()
    for k in range(n_params):
        api.addSampledParam(m, f"p{k}", 0.1, -1.0, 1.0)
    for k in range(n_params):
        api.addMoment(m, f"m{k}", 0.0, 1.0)
    api.addEvalFunc(m, objective or api.objfunc_norm)
    return m


def test_exchange_mode_selection_and_fallback(monkeypatch):
    """the host mirror picks the barrier-free persistent kernel for every world size, the multi-launch kernels for
    the panel objective / more than 32 parameters, never overrides an explicit request, and falls back to mode 0 when
    the library says the persistent kernel does not fit the shape (SMM_E_UNSUPPORTED_SHAPE)"""
    from smm_jl_b200 import api, _lib
    from smm_jl_b200._abi import SMM_E_UNSUPPORTED_SHAPE, SMM_E_CUDA
    m = _toy_problem(api)
    assert api._exchange_mode(m, {}, 2) == 2
    assert api._exchange_mode(m, {"world_size": 8}, 2) == 3      # several GPUs: flag-in-data hand-over
    assert api._exchange_mode(m, {"exchange_mode": 1, "world_size": 8}, 2) == 1
    assert api._exchange_mode(m, {}, 40) == 0
    assert api._exchange_mode(_toy_problem(api, 2, api.objfunc_panel), {}, 2) == 0

    created = []

    class FakeHandle:
        def __init__(self, cfg):
            created.append(cfg.exchange_mode)
            if cfg.exchange_mode != 0:
                raise _lib.SMMError(SMM_E_UNSUPPORTED_SHAPE, "persistent kernel does not fit on an SM")

        def close(self):
            pass

    monkeypatch.setattr(_lib, "BGPHandle", FakeHandle)
    algo = api.MAlgoBGP(m, {"N": 4, "maxiter": 5})
    algo._handle()
    assert created == [2, 0] and algo._cfg.exchange_mode == 0
    created.clear()
    algo = api.MAlgoBGP(m, {"N": 4, "maxiter": 5, "exchange_mode": 1})      # explicit: the error surfaces
    with pytest.raises(_lib.SMMError):
        algo._handle()
    assert created == [1]

    class Broken(FakeHandle):
        def __init__(self, cfg):
            raise _lib.SMMError(SMM_E_CUDA, "no such CUDA device")

    monkeypatch.setattr(_lib, "BGPHandle", Broken)
    with pytest.raises(_lib.SMMError):                                        # other failures are never retried
        api.MAlgoBGP(m, {"N": 4, "maxiter": 5})._handle()


def test_library_holds_both_builds_of_the_persistent_kernel():
    """smm_kernels.cu is compiled twice (smm_jl_b200/build.py): plain, and with -DSMM_LL_TU into smm::ll for
    exchange_mode 3.  Both launchers and both kernels must be in the shared library (no GPU needed to check)."""
    import subprocess
    from smm_jl_b200 import _lib
    _lib.lib()
    syms = subprocess.run(["nm", "-C", _lib.lib_path()], capture_output=True, text=True).stdout
    assert "smm::launch_persistent(" in syms and "smm::ll::launch_persistent(" in syms
    assert "smm::bgp_persistent_kernel<true>" in syms and "smm::ll::bgp_persistent_kernel<true>" in syms
