"""GPU parity of the dynamic-panel objective (BASELINE config C4) through the C ABI: the bare objective
(`smm_bgp_eval_batch`) and full BGP runs against the CPU oracle, plus size-independent properties at the full
C4 shape (T = 50, N_ind = 5000, K = 8): scheduling invariance to the bit, value 0 at the data-generating
point, replicas of one parameter vector give bit-equal values."""
import numpy as np
import pytest

from smm_jl_b200 import configs
from tests.parity import assert_trace_parity, max_rel_err

pytestmark = pytest.mark.gpu


def oracle_data_moments(oracle, K, T, NI):
    return configs.panel_data_moments(lambda c, p: oracle.eval_batch(c, p), K, T, NI)


@pytest.mark.parametrize("K,T,NI", [(1, 7, 5), (2, 9, 33), (3, 12, 100), (5, 20, 64), (8, 50, 257), (8, 7, 31), (16, 8, 40)])
def test_panel_eval_batch_parity(smm, oracle, K, T, NI):
    """K = 8 takes the register-resident kernel, every other K the run-time-K instantiation"""
    cfg = configs.dynamic_panel(1, 1, K, T, NI)
    lb, ub = configs.panel_box(K)
    rng = np.random.default_rng(K * 1000 + T)
    params = lb + rng.uniform(0.0, 1.0, size=(9, lb.size)) * (ub - lb)
    params[0] = lb          # corners of the box: the bounds the exact accumulators are sized for
    params[1] = ub
    with smm.BGPHandle(cfg) as h:
        v, m, s = h.eval_batch(params)
        v2, m2, s2 = h.eval_batch(params, noseed=1, rep0=5)
    ov, om, os_ = oracle.eval_batch(cfg, params, n_threads=8)
    ov2, om2, _ = oracle.eval_batch(cfg, params, noseed=1, rep0=5, n_threads=8)
    np.testing.assert_array_equal(s, os_)
    np.testing.assert_allclose(m, om, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(v, ov, rtol=1e-6)          # north_star tolerance
    np.testing.assert_allclose(v, ov, rtol=1e-9)          # what the implementation actually delivers
    np.testing.assert_allclose(m2, om2, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(v2, ov2, rtol=1e-9)
    assert not np.allclose(v, v2)


def test_panel_eval_rejects_params_outside_the_box(smm):
    cfg = configs.dynamic_panel(1, 1, 2, 8, 16)
    lb, ub = configs.panel_box(2)
    p = 0.5 * (lb + ub)
    p[0] = 0.99
    with smm.BGPHandle(cfg) as h:
        with pytest.raises(smm.SMMError) as e:
            h.eval_batch(p)
    assert e.value.code == -1


def test_panel_needs_stationary_box(smm):
    cfg = configs.dynamic_panel(1, 1, 2, 8, 16)
    cfg.ub = np.array(cfg.ub, dtype=float)
    cfg.ub[0] = 1.0
    with pytest.raises(smm.SMMError) as e:
        smm.BGPHandle(cfg)
    assert e.value.code == -4


@pytest.mark.parametrize("K,T,NI,n_chains,niter", [(2, 9, 100, 6, 30), (8, 12, 200, 8, 25), (3, 8, 40, 3, 40)])
def test_panel_chain_parity(smm, oracle, K, T, NI, n_chains, niter):
    dm = oracle_data_moments(oracle, K, T, NI)
    cfg = configs.dynamic_panel(n_chains, niter, K, T, NI, data_mom=dm, sigma_update_steps=5)
    with smm.BGPHandle(cfg) as h:
        h.step(niter // 2)
        h.step(niter - niter // 2)
        tr = h.read_trace(1, niter)
        sigma, acc = h.chain_state()
        ctr = h.counters()
    ref = oracle.run(cfg, niter, n_threads=8)
    assert_trace_parity(tr, ref.trace)
    np.testing.assert_array_equal(sigma, ref.sigma)
    assert ctr["swaps"] == ref.swaps
    assert ctr["proposal_attempts"] == ref.attempts
    assert np.array_equal(tr.params.view(np.uint64), ref.trace.params.view(np.uint64))
    assert max_rel_err(tr, ref.trace) < 1e-8


def test_panel_c4_shape_against_oracle(smm, oracle):
    """full C4 panel (K = 8, T = 50, 5000 individuals): bare objective at four points + a short 8-chain run"""
    K, T, NI = 8, 50, 5000
    lb, ub = configs.panel_box(K)
    dm_gpu = configs.panel_data_moments_gpu(K, T, NI)
    dm = oracle_data_moments(oracle, K, T, NI)
    np.testing.assert_allclose(dm_gpu, dm, rtol=1e-10, atol=1e-12)
    cfg = configs.dynamic_panel(8, 6, K, T, NI, data_mom=dm)
    rng = np.random.default_rng(4)
    params = lb + rng.uniform(0.0, 1.0, size=(4, lb.size)) * (ub - lb)
    with smm.BGPHandle(cfg) as h:
        v, m, s = h.eval_batch(params)
        h.step(6)
        tr = h.read_trace(1, 6)
    ov, om, os_ = oracle.eval_batch(cfg, params, n_threads=4)
    np.testing.assert_allclose(m, om, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(v, ov, rtol=1e-9)
    ref = oracle.run(cfg, 6, n_threads=8)
    assert_trace_parity(tr, ref.trace)


def test_panel_c4_properties(smm):
    """size-independent properties at the full C4 shape, no oracle involved"""
    K, T, NI = 8, 50, 5000
    lb, ub = configs.panel_box(K)
    theta0 = 0.5 * (lb + ub)
    dm = configs.panel_data_moments_gpu(K, T, NI)
    # (1) at the data-generating point with the data-generating shocks the distance is exactly 0
    cfg0 = configs.dynamic_panel(1, 1, K, T, NI, data_mom=dm, seed_sim=4321)
    with smm.BGPHandle(cfg0) as h:
        v, m, _ = h.eval_batch(theta0)
    assert v[0] == 0.0 and np.array_equal(m[0], dm)
    # (2) the pooled sums are exact integers: any work distribution gives the same bits
    rng = np.random.default_rng(11)
    params = lb + rng.uniform(0.0, 1.0, size=(24, lb.size)) * (ub - lb)
    params[12:] = params[:12]          # (3) replicas of a parameter vector: bit-equal values (ties stay ties)
    outs = []
    for n_split in (1, 7, 0):           # n_split caps the CTA count of the panel kernel: 1, 7, one full wave
        cfg = configs.dynamic_panel(1, 1, K, T, NI, data_mom=dm, n_split=n_split)
        with smm.BGPHandle(cfg) as h:
            outs.append(h.eval_batch(params))
    for v, m, s in outs[1:]:
        assert np.array_equal(v.view(np.uint64), outs[0][0].view(np.uint64))
        assert np.array_equal(m.view(np.uint64), outs[0][1].view(np.uint64))
    v = outs[0][0]
    assert np.array_equal(v[:12].view(np.uint64), v[12:].view(np.uint64))
    assert np.all(v > 0) and np.all(np.isfinite(v))
    # (4) a run is independent of how step() is chunked and of the CTA count
    traces = []
    for n_split, chunks in ((0, (12,)), (5, (1, 4, 7))):
        cfg = configs.dynamic_panel(16, 12, K, T, NI, data_mom=dm, n_split=n_split)
        with smm.BGPHandle(cfg) as h:
            for c in chunks:
                h.step(c)
            traces.append(h.read_trace(1, 12))
    for f in traces[0].FLOAT_FIELDS + traces[0].INT_FIELDS:
        a, b = getattr(traces[0], f), getattr(traces[1], f)
        assert np.array_equal(a, b, equal_nan=True) if a.dtype.kind == "f" else np.array_equal(a, b), f
    # iteration 1 evaluates every chain at the same point: equal values, hence no swap at iteration 2 from ties
    assert len(set(traces[0].value[0].tolist())) == 1
