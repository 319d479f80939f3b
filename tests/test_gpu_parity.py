"""GPU parity tests proper: the CUDA path through the C ABI vs the CPU oracle on the same inputs."""
import numpy as np
import pytest

from smm_jl_b200 import configs
from smm_jl_b200._abi import SMM_OBJ_FAILS
from tests.parity import assert_trace_parity, max_rel_err

pytestmark = pytest.mark.gpu


def run_gpu(smm, cfg, n):
    with smm.BGPHandle(cfg) as h:
        h.step(n)
        tr = h.read_trace(1, n)
        sigma, acc = h.chain_state()
        ctr = h.counters()
    return tr, sigma, acc, ctr


def test_normals_bit_exact(smm, oracle):
    """the stream transform is bit-identical on host and device"""
    for (seed, k, c2, c3) in [(1234, 0, 0, 1 << 28), (1234, 7, 0, 1 << 28), (99, 3, 17, (2 << 28) | 5)]:
        g = smm.debug_normals(seed, k, c2, c3, 20000)
        o = oracle.normals(seed, k, c2, c3, 20000)
        assert np.array_equal(g.view(np.uint64), o.view(np.uint64))


@pytest.mark.parametrize("niter", [1, 2, 50])
def test_c1_serial_normal_parity(smm, oracle, niter):
    cfg = configs.c1_serial_normal(niter)
    tr, sigma, acc, ctr = run_gpu(smm, cfg, niter)
    ref = oracle.run(cfg, niter)
    assert_trace_parity(tr, ref.trace)
    np.testing.assert_array_equal(sigma, ref.sigma)
    np.testing.assert_allclose(acc, ref.accept_rate, rtol=0, atol=0)
    assert ctr["swaps"] == ref.swaps
    assert ctr["proposal_attempts"] == ref.attempts


def test_params_bit_exact(smm, oracle):
    """proposals are pure functions of the streams: parameter traces must agree to the bit"""
    cfg = configs.c1_serial_normal(100)
    tr, *_ = run_gpu(smm, cfg, 100)
    ref = oracle.run(cfg, 100)
    assert np.array_equal(tr.params.view(np.uint64), ref.trace.params.view(np.uint64))


@pytest.mark.parametrize("n_chains,n_split", [(16, 0), (16, 1), (16, 5), (64, 0)])
def test_mvnormal_parity(smm, oracle, n_chains, n_split):
    niter = 40
    cfg = configs.mvnormal(n_chains, niter, n_split=n_split)
    tr, sigma, acc, ctr = run_gpu(smm, cfg, niter)
    ref = oracle.run(cfg, niter, n_threads=8)
    assert_trace_parity(tr, ref.trace)
    np.testing.assert_array_equal(sigma, ref.sigma)
    assert ctr["swaps"] == ref.swaps
    assert max_rel_err(tr, ref.trace) < 1e-9


def test_eval_batch_parity(smm, oracle):
    cfg = configs.mvnormal(4, 4)
    rng = np.random.default_rng(3)
    params = rng.uniform(-3, 3, size=(33, 8))
    with smm.BGPHandle(cfg) as h:
        v, m, s = h.eval_batch(params)
        v2, m2, s2 = h.eval_batch(params, noseed=1, rep0=7)
    ov, om, os_ = oracle.eval_batch(cfg, params)
    ov2, om2, os2 = oracle.eval_batch(cfg, params, noseed=1, rep0=7)
    np.testing.assert_allclose(v, ov, rtol=1e-9)
    np.testing.assert_allclose(m, om, rtol=1e-9, atol=1e-12)
    np.testing.assert_array_equal(s, os_)
    np.testing.assert_allclose(v2, ov2, rtol=1e-9)
    np.testing.assert_allclose(m2, om2, rtol=1e-9, atol=1e-12)
    assert not np.allclose(v, v2)


def test_pairs_schedule(smm, oracle):
    """device Pairs[iter] = oracle's sample; the level order is a valid parallel schedule"""
    cfg = configs.mvnormal(64, 8)
    with smm.BGPHandle(cfg) as h:
        for it in (2, 3, 7):
            ij, off, nlev = h.debug_pairs(it)
            want = oracle.pairs(cfg.seed_algo, it, 64)
            assert sorted(map(tuple, ij.tolist())) == sorted(map(tuple, want.tolist()))
            assert off[0] == 0 and off[-1] == len(ij)
            # inside a level no chain appears twice
            for l in range(nlev):
                seg = ij[off[l]:off[l + 1]].reshape(-1)
                assert len(set(seg.tolist())) == len(seg)
            # pairs sharing a chain keep their sample order across levels
            pos = {tuple(p): t for t, p in enumerate(map(tuple, ij.tolist()))}
            order = [pos[tuple(p)] for p in want.tolist()]
            last = {}
            for t, p in zip(order, want.tolist()):
                for c in p:
                    assert last.get(c, -1) < t
                    last[c] = t


def test_failing_objective(smm, oracle):
    cfg = configs.c1_serial_normal(12, objective_id=SMM_OBJ_FAILS)
    tr, *_ = run_gpu(smm, cfg, 12)
    ref = oracle.run(cfg, 12)
    assert_trace_parity(tr, ref.trace)
    assert (tr.status[0] == 1).all() and (tr.status[1:][tr.exchanged[1:] == 0] == -2).all()


def test_batch_size_one(smm, oracle):
    cfg = configs.mvnormal(8, 30, batch_size=1)
    tr, *_ = run_gpu(smm, cfg, 30)
    ref = oracle.run(cfg, 30, n_threads=8)
    assert_trace_parity(tr, ref.trace)
