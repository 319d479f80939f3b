"""exchange_mode 1 and 2: the persistent cooperative kernel (with grid barriers / barrier-free) must reproduce the
oracle (and therefore the multi-launch path) exactly like mode 0 does."""
import numpy as np
import pytest

from smm_jl_b200 import configs
from smm_jl_b200._abi import SMM_OBJ_FAILS
from tests.parity import assert_trace_parity

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[1, 2, 3], ids=["barriers", "dataflow", "flag-in-data"])
def pmode(request):
    """exchange_mode 1 = persistent kernel with grid barriers, 2 = barrier-free (one completion counter per rank),
    3 = barrier-free with flag-in-data words for the values / proposal centres"""
    return request.param


def run_gpu(smm, cfg, n, pieces=None):
    with smm.BGPHandle(cfg) as h:
        for k in (pieces or [n]):
            h.step(k)
        tr = h.read_trace(1, n)
        sigma, acc = h.chain_state()
        ctr = h.counters()
    return tr, sigma, acc, ctr


@pytest.mark.parametrize("niter", [1, 2, 3, 50])
def test_c1_persistent(smm, pmode, oracle, niter):
    cfg = configs.c1_serial_normal(niter, exchange_mode=pmode)
    tr, sigma, acc, ctr = run_gpu(smm, cfg, niter)
    ref = oracle.run(cfg, niter)
    assert_trace_parity(tr, ref.trace)
    np.testing.assert_array_equal(sigma, ref.sigma)
    np.testing.assert_array_equal(acc, ref.accept_rate)
    assert ctr["swaps"] == ref.swaps and ctr["proposal_attempts"] == ref.attempts


@pytest.mark.parametrize("n_chains,n_split", [(16, 0), (64, 0), (64, 40), (300, 0), (1, 0), (1024, 0)])
def test_mvnormal_persistent(smm, pmode, oracle, n_chains, n_split):
    niter = 30
    cfg = configs.mvnormal(n_chains, niter, exchange_mode=pmode, n_split=n_split, n_sim=2000)
    tr, sigma, acc, ctr = run_gpu(smm, cfg, niter)
    ref = oracle.run(cfg, niter, n_threads=8)
    assert_trace_parity(tr, ref.trace)
    np.testing.assert_array_equal(sigma, ref.sigma)
    assert ctr["swaps"] == ref.swaps


def test_persistent_equals_multilaunch_bitwise_params(smm, pmode):
    """both modes share the device functions and the order-invariant accumulators: traces are bit-identical"""
    a, *_ = run_gpu(smm, configs.mvnormal(64, 40, exchange_mode=0), 40)
    b, *_ = run_gpu(smm, configs.mvnormal(64, 40, exchange_mode=pmode), 40)
    # order-invariant accumulation: how the draw space is cut into CTAs does not change a single bit
    for f in a.INT_FIELDS + a.FLOAT_FIELDS:
        assert np.array_equal(getattr(a, f), getattr(b, f), equal_nan=True), f


def test_persistent_in_pieces_and_across_chunks(smm, pmode, oracle):
    """several launches (each up to 128 iterations) chain up exactly; 300 iterations cross two chunk borders"""
    n = 300
    cfg = configs.c1_serial_normal(n, exchange_mode=pmode, n_sim=500)
    tr_a, *_ = run_gpu(smm, cfg, n)
    tr_b, *_ = run_gpu(smm, cfg, n, pieces=[1, 1, 2, 127, 129, 40])
    for f in tr_a.FLOAT_FIELDS + tr_a.INT_FIELDS:
        assert np.array_equal(getattr(tr_a, f), getattr(tr_b, f), equal_nan=True), f
    assert_trace_parity(tr_a, oracle.run(cfg, n).trace)


def test_persistent_checkpoint(smm, pmode):
    cfg = configs.mvnormal(16, 40, exchange_mode=pmode, n_sim=1000)
    tr_a, sig_a, *_ = run_gpu(smm, cfg, 40)
    with smm.BGPHandle(cfg) as h:
        h.step(17)
        blob = h.export_state()
    with smm.BGPHandle(configs.mvnormal(16, 40, exchange_mode=pmode, n_sim=1000)) as h2:
        h2.import_state(blob)
        h2.step(23)
        tr_b = h2.read_trace(1, 40)
        sig_b, _ = h2.chain_state()
    for f in tr_a.FLOAT_FIELDS + tr_a.INT_FIELDS:
        assert np.array_equal(getattr(tr_a, f), getattr(tr_b, f), equal_nan=True), f
    assert np.array_equal(sig_a, sig_b)


def test_persistent_other_objectives(smm, pmode, oracle):
    cfg = configs.c1_serial_normal(12, objective_id=SMM_OBJ_FAILS, exchange_mode=pmode)
    tr, *_ = run_gpu(smm, cfg, 12)
    assert_trace_parity(tr, oracle.run(cfg, 12).trace)
    cfg = configs.mvnormal(8, 20, batch_size=1, exchange_mode=pmode, n_sim=999)
    tr, *_ = run_gpu(smm, cfg, 20)
    assert_trace_parity(tr, oracle.run(cfg, 20, n_threads=4).trace)
    cfg = configs.mvnormal(8, 12, noseed=1, exchange_mode=pmode, n_sim=1500)
    tr, *_ = run_gpu(smm, cfg, 12)
    assert_trace_parity(tr, oracle.run(cfg, 12, n_threads=4).trace)
    cfg = configs.slow_normal(8, 6, slow_seconds=0.01, exchange_mode=pmode)
    tr, *_ = run_gpu(smm, cfg, 6)
    assert_trace_parity(tr, oracle.run(configs.slow_normal(8, 6, slow_seconds=0.0), 6).trace)


@pytest.mark.parametrize("n_params,n_sim", [(1, 777), (3, 1000), (5, 4001), (16, 512), (20, 999), (32, 300)])
def test_persistent_ragged_shapes(smm, pmode, oracle, n_params, n_sim):
    """rows that do not tile a warp (masked lanes), odd draw counts, units that end inside a step: the deferred
    ziggurat queue and the exact accumulators must still reproduce the oracle's sequential sums"""
    niter = 12
    cfg = configs.mvnormal(12, niter, n_params=n_params, exchange_mode=pmode, n_sim=n_sim)
    tr, sigma, acc, ctr = run_gpu(smm, cfg, niter)
    ref = oracle.run(cfg, niter, n_threads=8)
    assert_trace_parity(tr, ref.trace)
    np.testing.assert_array_equal(sigma, ref.sigma)
