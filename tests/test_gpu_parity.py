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


def test_ziggurat_normals_bit_exact(smm, oracle):
    """the simulator stream of the MvNormal objectives (fast path, wedges, base strip, tails, retries: 0.8 % of 600 000
    draws leave the fast path) is bit-identical on host and device"""
    for (seed, k, c2, c3) in [(1234, 0, 0, 1 << 28), (1234, 7, 0, 1 << 28), (99, 3, 17, (1 << 28) | 5)]:
        g = smm.debug_zig_normals(seed, k, c2, c3, 200000)
        o = oracle.zig_normals(seed, k, c2, c3, 200000)
        assert np.array_equal(g.view(np.uint64), o.view(np.uint64))
        assert np.abs(g).max() > 3.8520461503683912      # the tail branch was exercised


@pytest.mark.parametrize("n_sim", [1, 2, 3, 4, 95, 96, 97, 1001, 4096])
def test_deferred_ziggurat_matches_sequential_definition(smm, oracle, n_sim):
    """the kernels resolve rejected fast-path candidates later, in warp-sized batches, and patch the exact integer
    accumulators; the totals must equal the oracle's draw-by-draw evaluation for ragged draw counts, in both modes"""
    rng = np.random.default_rng(n_sim)
    for P in (1, 3, 8):
        cfg = configs.mvnormal(1, 1, n_params=P, n_sim=max(n_sim, 2))
        p = rng.uniform(-3, 3, (24, P))
        for noseed in (0, 1):
            with smm.BGPHandle(cfg) as h:
                v, m, st = h.eval_batch(p, noseed=noseed, rep0=5)
            vo, mo, so = oracle.eval_batch(cfg, p, noseed=noseed, rep0=5, n_threads=4)
            np.testing.assert_array_equal(st, so)
            # the fixed-point grids (2^-40 for a block's sum of x, 2^-34 for its sum of x^2 at these shapes) bound the absolute error of a moment; a sample variance
            # of two or three draws can be tiny, so the bound is absolute here (north star: 1e-6 relative)
            np.testing.assert_allclose(m, mo, rtol=1e-9, atol=1e-9)
            np.testing.assert_allclose(v, vo, rtol=1e-7, atol=1e-9)


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


def test_golden_vectors_on_gpu(smm):
    """the committed fixtures (tests/golden/make_golden.py) without the oracle in the loop"""
    import os
    from smm_jl_b200._abi import Trace
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "bgp_golden.npz"))
    assert np.array_equal(smm.debug_normals(1234, 0, 0, 1 << 28, 64).view(np.uint64), g["normals_bits"])
    assert np.array_equal(smm.debug_zig_normals(1234, 0, 0, 1 << 28, 4096).view(np.uint64), g["zig_normals_bits"])
    for tag, cfg, n in (("c1", configs.c1_serial_normal(40), 40), ("mv", configs.mvnormal(8, 10), 10)):
        tr, sigma, *_ = run_gpu(smm, cfg, n)
        for f in Trace.INT_FIELDS:
            np.testing.assert_array_equal(getattr(tr, f), g[f"{tag}_{f}"], err_msg=f)
        for f in Trace.FLOAT_FIELDS:
            np.testing.assert_allclose(getattr(tr, f), g[f"{tag}_{f}"], rtol=1e-6, atol=1e-12, err_msg=f)
        np.testing.assert_array_equal(sigma, g[f"{tag}_sigma"])


def test_step_in_pieces_equals_one_call(smm):
    """run! iteration by iteration (the reference's loop) == n iterations enqueued at once"""
    cfg = configs.mvnormal(16, 30)
    tr_a, *_ = run_gpu(smm, cfg, 30)
    with smm.BGPHandle(cfg) as h:
        for n in (1, 1, 5, 13, 10):
            h.step(n)
        tr_b = h.read_trace(1, 30)
    for f in tr_a.FLOAT_FIELDS + tr_a.INT_FIELDS:
        assert np.array_equal(getattr(tr_a, f), getattr(tr_b, f), equal_nan=True), f


def test_checkpoint_restart_is_bit_exact(smm):
    """save / readMalgo / restart! (AlgoAbstract.jl:83-102, AlgoBGP.jl:804-884): a run resumed from an
    exported state equals the straight run bit for bit (all randomness is counter-indexed)"""
    cfg = configs.mvnormal(16, 40)
    tr_a, sig_a, acc_a, _ = run_gpu(smm, cfg, 40)
    with smm.BGPHandle(cfg) as h:
        h.step(17)
        blob = h.export_state()
    with smm.BGPHandle(configs.mvnormal(16, 40)) as h2:
        h2.import_state(blob)
        assert h2.iteration == 17
        h2.step(23)
        tr_b = h2.read_trace(1, 40)
        sig_b, acc_b = h2.chain_state()
    for f in tr_a.FLOAT_FIELDS + tr_a.INT_FIELDS:
        assert np.array_equal(getattr(tr_a, f), getattr(tr_b, f), equal_nan=True), f
    assert np.array_equal(sig_a, sig_b) and np.array_equal(acc_a, acc_b)


def test_large_config_properties(smm):
    """BASELINE C2 at full size (256 chains x 10 000 draws): size-independent properties"""
    n = 60
    cfg = configs.mvnormal(256, n)
    tr, sigma, acc, ctr = run_gpu(smm, cfg, n)
    # running-minimum / index bookkeeping (set_eval!)
    np.testing.assert_array_equal(tr.best_val, np.minimum.accumulate(tr.value, axis=0))
    rows = tr.best_id - 1
    np.testing.assert_array_equal(np.take_along_axis(tr.value, rows, axis=0), tr.best_val)
    # curr_val follows accepted values
    want = np.where(tr.accepted[1:] == 1, tr.value[1:], tr.curr_val[:-1])
    np.testing.assert_array_equal(tr.curr_val[1:], want)
    # common random numbers: simulated means are affine in the parameters, variances are constant
    zbar = tr.sim_moments[:, :, :8] - tr.params
    assert np.ptp(zbar, axis=(0, 1)).max() < 1e-12
    assert np.ptp(tr.sim_moments[:, :, 8:], axis=(0, 1)).max() < 1e-9
    # value is the weighted distance of the stored moments
    d = (tr.sim_moments - np.asarray(cfg.data_mom)) / np.asarray(cfg.data_w)
    np.testing.assert_allclose(tr.value, (d * d).mean(axis=2), rtol=1e-12)
    # exchanges are symmetric, partner ids valid, and swap accepted records
    ex = tr.exchanged
    assert (ex[0] == 0).all() and ex.max() <= 256
    its, cs = np.nonzero(ex)
    assert len(its) > 0 and (tr.accepted[its, cs] == 1).all()
    # a chain that swapped exactly once points at a partner that points back
    assert ctr["swaps"] > 0 and ctr["evaluations"] == 256 * n
    # proposals stay inside the box
    assert (tr.params >= -3).all() and (tr.params <= 3).all()
    assert (sigma > 0).all() and ((acc >= 0) & (acc <= 1)).all()


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_c2_full_size_against_the_oracle(smm, oracle, mode):
    """BASELINE C2 at FULL size (256 chains x 10 000 draws, the benchmarked configuration) in every exchange mode the
    bench can run, compared entry by entry with the oracle: bookkeeping exact, floats to 1e-6 relative"""
    import os
    n = 60
    cfg = configs.mvnormal(256, n, exchange_mode=mode)
    tr, sigma, acc, ctr = run_gpu(smm, cfg, n)
    ref = oracle.run(configs.mvnormal(256, n), n, n_threads=os.cpu_count() or 1)
    assert_trace_parity(tr, ref.trace)
    assert np.array_equal(tr.params.view(np.uint64), ref.trace.params.view(np.uint64))
    np.testing.assert_array_equal(sigma, ref.sigma)
    np.testing.assert_array_equal(acc, ref.accept_rate)
    assert ctr["swaps"] == ref.swaps and ctr["proposal_attempts"] == ref.attempts
    assert max_rel_err(tr, ref.trace) < 1e-9


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_zero_and_negative_weights_on_device(smm, oracle, mode):
    """The reference's own fixture carries zero and negative moment weights (test/include/test-include.jl:78): a zero
    weight divides by zero (ObjExamples.jl:97) -> value = +Inf for every evaluation, a negative one is squared away.
    From iteration 2 on, exp(acc_tuner * (Inf - Inf)) = NaN -> rejected with status -1 (AlgoBGP.jl:350-353), and the
    exchange compares Inf - Inf = NaN > min_improve -> never swaps (:688)."""
    n = 30
    cfg = configs.c1_serial_normal(n, data_w=[0.0, -1.0], exchange_mode=mode)
    tr, sigma, acc, ctr = run_gpu(smm, cfg, n)
    ref = oracle.run(cfg, n)
    assert_trace_parity(tr, ref.trace)
    np.testing.assert_array_equal(sigma, ref.sigma)
    np.testing.assert_array_equal(acc, ref.accept_rate)
    assert np.isinf(tr.value).all() and (tr.status[1:] == -1).all() and (tr.prob[1:] == 0).all()
    assert (tr.accepted[0] == 1).all() and (tr.accepted[1:] == 0).all() and ctr["swaps"] == 0


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_nonfinite_branches_of_accept_reject_on_device(smm, oracle, mode):
    """A weight of 1e-155 makes ((sim - data) / w)^2 overflow unless |sim - data| < 0.13, so chains wander between
    finite (~1e300) and infinite objective values and doAcceptReject! takes every branch (AlgoBGP.jl:336-367) on the
    device: Inf -> Inf = NaN prob -> status -1; Inf -> finite = forced accept with prob 1 (:355-359); finite -> Inf =
    prob exp(-Inf) = 0, status 1; finite -> finite = the ordinary Metropolis test."""
    n = 60
    kw = dict(n_params=2, data_w=[1e-155, 1.0, 1.0, 1.0], n_sim=2000)
    tr, sigma, acc, ctr = run_gpu(smm, configs.mvnormal(8, n, exchange_mode=mode, **kw), n)
    ref = oracle.run(configs.mvnormal(8, n, **kw), n, n_threads=4)
    assert_trace_parity(tr, ref.trace)
    np.testing.assert_array_equal(sigma, ref.sigma)
    np.testing.assert_array_equal(acc, ref.accept_rate)
    assert ctr["swaps"] == ref.swaps
    # the branches were really taken
    assert (tr.status == -1).sum() > 10
    forced = (~np.isfinite(tr.curr_val[:-1])) & np.isfinite(tr.value[1:]) & (tr.accepted[1:] == 1) & (tr.exchanged[1:] == 0)
    assert forced.sum() >= 1 and (tr.prob[1:][forced] == 1.0).all()
    assert ((tr.prob[1:] == 0) & (tr.status[1:] == 1) & np.isinf(tr.value[1:])).sum() > 10


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_nan_objective_is_the_reference_error(smm, oracle, mode):
    """`eval_new.value >= 0 || error(...)` (AlgoBGP.jl:341) also fires for NaN: a NaN data moment makes every value
    NaN, iteration 1 is accepted unconditionally (:327-333), iteration 2 aborts the run -- on both sides"""
    from smm_jl_b200._abi import SMM_E_NEGATIVE_OBJECTIVE
    cfg = configs.c1_serial_normal(5, data_mom=[float("nan"), 10.0], exchange_mode=mode)
    with pytest.raises(smm.SMMError) as e:
        with smm.BGPHandle(cfg) as h:
            h.step(5)
    assert e.value.code == SMM_E_NEGATIVE_OBJECTIVE
    with pytest.raises(oracle.OracleError) as e2:
        oracle.run(cfg, 5)
    assert e2.value.code == SMM_E_NEGATIVE_OBJECTIVE


def test_sampler_exhaustion_is_an_error(smm, oracle):
    """AlgoBGP.jl:409: `error("no draw in support ...")` with a single batch"""
    from smm_jl_b200._abi import SMM_E_SAMPLER_EXHAUSTED
    cfg = configs.mvnormal(4, 6, sigma0=[50.0] * 4, smpl_iters=3)
    with pytest.raises(smm.SMMError) as e:
        with smm.BGPHandle(cfg) as h:
            h.step(6)
    assert e.value.code == SMM_E_SAMPLER_EXHAUSTED
    with pytest.raises(oracle.OracleError) as e2:
        oracle.run(cfg, 6)
    assert e2.value.code == SMM_E_SAMPLER_EXHAUSTED
    # several batches: the failure is swallowed and the parameter falls to its lower bound (:445-452)
    cfg = configs.mvnormal(4, 6, sigma0=[50.0] * 4, smpl_iters=3, batch_size=1)
    tr, *_ = run_gpu(smm, cfg, 6)
    ref = oracle.run(cfg, 6)
    assert_trace_parity(tr, ref.trace)
    assert (tr.params[1:] == -3.0).any()


def test_odd_draw_count_and_other_dims(smm, oracle):
    for P, S in ((2, 9999), (5, 1001), (16, 2000)):
        cfg = configs.mvnormal(6, 8, n_params=P, n_sim=S)
        tr, *_ = run_gpu(smm, cfg, 8)
        ref = oracle.run(cfg, 8, n_threads=4)
        assert_trace_parity(tr, ref.trace)
    cfg = configs.normal_means(6, 8, n_params=18, batch_size=1, sigma0=[0.001] * 6)   # snorm_18 shape (Examples.jl:232)
    tr, *_ = run_gpu(smm, cfg, 8)
    assert_trace_parity(tr, oracle.run(cfg, 8, n_threads=4).trace)


def test_noseed_run(smm, oracle):
    cfg = configs.mvnormal(8, 12, noseed=1)
    tr, *_ = run_gpu(smm, cfg, 12)
    assert_trace_parity(tr, oracle.run(cfg, 12, n_threads=4).trace)


def test_slow_objective(smm, oracle):
    import time
    cfg = configs.slow_normal(8, 6, slow_seconds=0.02)
    t0 = time.time()
    tr, *_ = run_gpu(smm, cfg, 6)
    assert time.time() - t0 >= 6 * 0.02
    cfg_fast = configs.slow_normal(8, 6, slow_seconds=0.0)
    assert_trace_parity(tr, oracle.run(cfg_fast, 6).trace)
