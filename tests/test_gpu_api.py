"""The reference-facing host API on the GPU: MProb / MAlgoBGP / run! / history / summary / restart!"""
import os

import numpy as np
import pytest

from smm_jl_b200 import api, configs
from tests.parity import assert_trace_parity

pytestmark = pytest.mark.gpu


def serial_normal_problem():
    # Examples.jl:373-446 (snorm_impl, npar = 2)
    m = api.MProb()
    api.addSampledParam(m, {"p1": [0.2, -3, 3], "p2": [-0.2, -20, 20]})
    api.addMoment(m, {"name": ["mu1", "mu2"], "value": [-1.0, 10.0], "weight": [1.0, 1.0]})
    api.addEvalFunc(m, api.objfunc_norm)
    opts = {"N": 3, "maxiter": 20, "maxtemp": 5, "coverage": 0.02, "smpl_iters": 1000, "parallel": False,
            "min_improve": [0.0] * 3, "acc_tuners": [20, 2, 1.0], "animate": False, "seed": 12}
    return m, opts


def test_serial_normal_runs(oracle):
    """test/test_algoBGP.jl:30-38: serialNormal(2,20): history is 20 x 9, o.i == 20"""
    m, opts = serial_normal_problem()
    MA = api.MAlgoBGP(m, opts)
    api.run(MA)
    assert MA.i == 20 and len(MA.chains) == 3
    h = api.history(MA.chains[0])
    assert h.shape == (20, 9)
    assert list(h.columns) == ["iter", "value", "accepted", "curr_val", "best_val", "prob", "exchanged", "p1", "p2"]
    # the same run through the oracle
    ref = oracle.run(configs.c1_serial_normal(20), 20)
    c = MA.chains[0]
    assert_trace_parity(MA._trace, ref.trace)
    assert c.id == 1 and c.iter == 20 and c.accept_rate == ref.accept_rate[0] and c.sigma == ref.sigma[0]
    np.testing.assert_array_equal(c.probs_acc, [oracle.acc_uniform(12, 0, it) for it in range(1, 21)])
    ev = c.evals[0]
    assert ev.accepted and ev.prob == 1.0 and ev.status == 1 and list(ev.params) == ["p1", "p2"]
    assert api.param(ev, "p1") == 0.2 and ev.simMoments["mu1"] == pytest.approx(ref.trace.sim_moments[0, 0, 0], rel=1e-9)
    s = api.summary(MA)
    assert list(s.columns) == ["id", "acc_rate", "perc_exchanged", "exchanged_most_with", "best_val"] and len(s) == 3
    v, idx = api.best(c)
    assert v == pytest.approx(ref.trace.value[:, 0].min(), rel=1e-9) and idx == int(np.argmin(ref.trace.value[:, 0])) + 1
    assert set(api.mean(c)) == {"p1", "p2"} and api.CI(c)["p1"].shape == (2,)
    assert len(api.allAccepted(c)) == int(ref.trace.accepted[:, 0].sum())
    MA.close()


def test_recover_the_mean():
    """test/test_algoBGP.jl:57-121: 2 chains x 200 iterations, median of chain 1 within 1.0 of (-1, 1)"""
    m = api.MProb()
    api.addSampledParam(m, {"p1": [0.2, -3, 3], "p2": [-0.2, -2, 2]})
    api.addMoment(m, {"name": ["mu1", "mu2"], "value": [-1.0, 1.0], "weight": [1.0, 1.0]})
    api.addEvalFunc(m, api.objfunc_norm)
    opts = {"N": 2, "maxiter": 200, "maxtemp": 5, "sigma_update_steps": 201, "sigma_adjust_by": 0.01, "smpl_iters": 1000,
            "parallel": True, "min_improve": [0.0, 0.0], "acc_tuners": [5, 1.0], "seed": 1234}
    MA = api.MAlgoBGP(m, opts)
    api.run(MA)
    med = api.median(MA.chains[0])
    assert abs(med["p1"] + 1.0) < 1.0 and abs(med["p2"] - 1.0) < 1.0
    MA.close()


def test_objfunc_norm_eval(oracle):
    """test/test_objfunc.jl:22-29 through evaluateObjective on the device"""
    m = api.MProb()
    api.addSampledParam(m, {"p1": [0.0, -3, 3], "p2": [0.0, -3, 3]})
    api.addMoment(m, {"mu1": {"value": 0.0, "weight": 1.0}, "mu2": {"value": 0.0, "weight": 1.0}})
    api.addEvalFunc(m, api.objfunc_norm)
    ev = api.evaluateObjective(m, {"p1": 0.0, "p2": 0.0})
    assert ev.status == 1
    assert abs(ev.simMoments["mu1"] - ev.dataMoments["mu1"]) < 0.1 and abs(ev.simMoments["mu2"]) < 0.1
    ev2 = api.evaluateObjective(m, {"p1": 0.0, "p2": 0.0}, noseed=True, rep=3)
    assert ev2.simMoments != ev.simMoments and ev2.options["noseed"]
    api.addEvalFunc(m, api.Testobj_fails)
    ev3 = api.evaluateObjective(m, {"p1": 0.0, "p2": 0.0})
    assert ev3.status == -2 and ev3.value == -1.0 and len(ev3.simMoments) == 0     # mprob.jl:183-186


def test_save_read_restart(tmp_path, oracle):
    """test/test_AlgoAbstract.jl:36-62 (readMalgo == in-memory) and the restart test the reference meant to
    write (test_algoBGP.jl:198-313): restart!(algo, k) == a straight run of maxiter + k"""
    m, opts = serial_normal_problem()
    fn = os.path.join(tmp_path, "algo.pkl")
    opts = dict(opts, maxiter=10, save_frequency=5, filename=fn)
    MA = api.MAlgoBGP(m, opts)
    api.run(MA)
    MB = api.readMalgo(fn)
    assert MB.i == MA.i == 10
    for a, b in zip(MA.chains, MB.chains):
        for f in ("best_id", "best_val", "curr_val", "accepted", "exchanged", "probs_acc"):
            np.testing.assert_array_equal(getattr(a, f), getattr(b, f))
        assert a.sigma == b.sigma and a.accept_rate == b.accept_rate
        assert a.evals[9] == b.evals[9]
    api.restart(MB, 15)
    assert MB.i == 25 and MB["maxiter"] == 25
    ref = oracle.run(configs.c1_serial_normal(25), 25)
    _ = MB.chains
    assert_trace_parity(MB._trace, ref.trace)
    MA.close()
    MB.close()


def test_checkpoint_blob_is_bound_to_its_ensemble(tmp_path, smm):
    """a state blob names the ensemble it belongs to (chains, world, rank, seeds): importing it anywhere else fails
    instead of resuming the wrong chains; save() after close() raises instead of writing an empty trace"""
    cfg = configs.mvnormal(16, 12, n_sim=300)
    with smm.BGPHandle(cfg) as h:
        h.step(6)
        blob = h.export_state()
    with smm.BGPHandle(configs.mvnormal(16, 12, n_sim=300, seed_algo=7)) as h2:
        with pytest.raises(smm.SMMError, match="streams"):
            h2.import_state(blob)
    with smm.BGPHandle(configs.mvnormal(16, 12, n_sim=300, seed_sim=99)) as h2:
        with pytest.raises(smm.SMMError, match="streams"):
            h2.import_state(blob)
    with smm.BGPHandle(configs.mvnormal(32, 12, n_sim=300)) as h2:
        with pytest.raises(smm.SMMError):
            h2.import_state(blob)
    with smm.BGPHandle(configs.mvnormal(16, 40, n_sim=300)) as h2:     # a longer run of the same ensemble: fine
        h2.import_state(blob)
        assert h2.iteration == 6
    m, opts = serial_normal_problem()
    MA = api.MAlgoBGP(m, dict(opts, maxiter=5))
    api.run(MA)
    MA.close()
    with pytest.raises(RuntimeError, match="closed"):
        api.save(MA, os.path.join(tmp_path, "late.pkl"))


@pytest.mark.parametrize("mode,window", [(0, 0), (1, 0), (1, 7), (1, 50), (0, 13), (3, 0), (3, 7)])
def test_streaming_run_equals_step_and_read(smm, mode, window):
    """smm_bgp_run (windows computed while the previous window's rows travel to the host) = step + read_trace"""
    import numpy as np
    from smm_jl_b200 import configs
    n = 150
    cfg = configs.mvnormal(24, n, exchange_mode=mode)
    with smm.BGPHandle(cfg) as h:
        h.step(n)
        want = h.read_trace(1, n)
    buf = smm.PinnedTrace.acquire(n, 24, cfg.n_params, cfg.n_moments)
    with smm.BGPHandle(cfg) as h:
        h.run(40, into=buf, window=window)            # rows 0..39 = iterations 1..40
        first = {f: np.array(getattr(buf, f)[:40]) for f in buf.FLOAT_FIELDS + buf.INT_FIELDS}
        h.run(n - 40, into=buf, window=window)        # rows 0..109 = iterations 41..150
        assert h.iteration == n
        for f in buf.FLOAT_FIELDS + buf.INT_FIELDS:
            w = getattr(want, f)
            a, b = first[f], getattr(buf, f)[: n - 40]
            if w.dtype.kind == "f":
                assert np.array_equal(a.view(np.uint64), w[:40].view(np.uint64)), f
                assert np.array_equal(b.view(np.uint64), w[40:].view(np.uint64)), f
            else:
                assert np.array_equal(a, w[:40]) and np.array_equal(b, w[40:]), f
    buf.release()


def test_run_serves_chains_from_the_streamed_trace(smm, oracle):
    from smm_jl_b200 import api, configs
    from oracle import oracle_lib  # noqa: F401
    import numpy as np
    m = api.MProb()
    api.addSampledParam(m, "p1", 0.2, -3.0, 3.0)
    api.addSampledParam(m, "p2", -0.2, -20.0, 20.0)
    api.addMoment(m, "mu1", -1.0, 1.0)
    api.addMoment(m, "mu2", 10.0, 1.0)
    api.addEvalFunc(m, api.objfunc_norm)
    opts = {"N": 3, "maxiter": 60, "maxtemp": 5.0, "sigma": 0.05, "acc_tuners": [20.0, 2.0, 1.0], "min_improve": [0.0] * 3,
            "seed": 12, "exchange_mode": 1}
    algo = api.MAlgoBGP(m, opts)
    api.run(algo)
    assert algo._streamed is not None and algo.i == 60
    ref = oracle.run(configs.c1_serial_normal(60), 60)
    for c in range(3):
        ch = algo.chains[c]
        assert np.array_equal(ch.accepted, ref.trace.accepted[:, c].astype(bool))
        assert np.array_equal(ch.exchanged, ref.trace.exchanged[:, c])
        np.testing.assert_allclose(ch.curr_val, ref.trace.curr_val[:, c], rtol=1e-9)
    algo.close()
    assert algo._streamed is None


@pytest.mark.parametrize("n_iter", [60, 9000])
def test_device_side_statistics_equal_the_host_ones(n_iter, smm):
    """mean / median / CI / summary of an algorithm are reduced on the device (smm_bgp_accepted_stats,
    smm_bgp_chain_summary; AlgoBGP.jl:174-206) and equal what the host computes from the full trace -- with a trace
    short enough for the shared-memory sort and one that needs the global scratch (> 8192 accepted draws possible)"""
    cfg = configs.mvnormal(12, n_iter, 4, n_sim=30, exchange_mode=2 if n_iter < 100 else 0)
    m = api.MProb()
    for k in range(4):
        api.addSampledParam(m, f"p{k + 1}", cfg.init[k], cfg.lb[k], cfg.ub[k])
    for k in range(8):
        api.addMoment(m, f"m{k + 1}", cfg.data_mom[k], cfg.data_w[k])
    api.addEvalFunc(m, api.objfunc_norm_mv)
    api.addEvalFuncOpts(m, {"n_sim": 30})
    MA = api.MAlgoBGP(m, {"N": 12, "maxiter": n_iter, "maxtemp": 5.0, "acc_tuners": list(np.asarray(cfg.acc_tuner)),
                          "min_improve": [0.0] * 12, "smpl_iters": 100000, "exchange_mode": cfg.exchange_mode,
                          "sigma_adjust_by": 0.0})   # constant sigma: the adaptation grows it without bound over 9000 iterations
    api.computeNextIteration(MA, n_iter)
    dev_mean, dev_med, dev_ci, dev_sum = api.mean(MA), api.median(MA), api.CI(MA, 0.9), api.summary(MA)
    cnt, _, _ = MA._handle().accepted_stats(())
    for ic, c in enumerate(MA.chains):
        assert cnt[ic] == int(c.accepted.sum())
        hm, hmed, hci, hs = api.mean(c), api.median(c), api.CI(c, 0.9), api.summary(c)
        for k in hm:
            np.testing.assert_allclose(dev_mean[ic][k], hm[k], rtol=1e-12, atol=1e-14)
            np.testing.assert_allclose(dev_med[ic][k], hmed[k], rtol=1e-13, atol=0)
            np.testing.assert_allclose(dev_ci[ic][k], hci[k], rtol=1e-13, atol=0)
        row = dev_sum.iloc[ic] if hasattr(dev_sum, "iloc") else dev_sum[ic]
        assert int(row["id"]) == hs["id"] and int(row["exchanged_most_with"]) == hs["exchanged_most_with"]
        assert float(row["best_val"]) == hs["best_val"] and float(row["acc_rate"]) == hs["acc_rate"]
        np.testing.assert_allclose(float(row["perc_exchanged"]), hs["perc_exchanged"], rtol=1e-15)
    MA.close()
