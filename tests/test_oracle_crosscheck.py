"""Pin the C++ oracle (and the stream header it shares with the kernels) against the independent
numpy / pure-Python re-derivation in oracle/oracle_np.py, and against the committed golden vectors."""
import os

import numpy as np
import pytest

from oracle import oracle_np
from smm_jl_b200 import configs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_philox_numpy_vs_cpp(oracle):
    rng = np.random.default_rng(0)
    for _ in range(20):
        c = [int(v) for v in rng.integers(0, 2 ** 32, 4)]
        k = [int(v) for v in rng.integers(0, 2 ** 32, 2)]
        got = oracle.philox(c, k)
        want = [int(v) for v in oracle_np.philox(*c, *k)]
        assert [int(v) for v in got] == want


def test_normals_numpy_vs_cpp(oracle):
    z = oracle.normals(1234, 3, 0, 1 << 28, 4000)                      # Box-Muller (proposals, panel)
    want = oracle_np.sim_normals(1234, 3, 8000, transform="bm")
    np.testing.assert_allclose(z, want, rtol=0, atol=4e-15)


def test_ziggurat_numpy_vs_cpp(oracle):
    """the simulator stream of the MvNormal objectives: header (fma-only exp/log, generated tables) against the
    numpy re-derivation (tables from mpmath, libm exp/log).  The fast path must agree to the bit."""
    n_blocks = 70_000
    z = oracle.zig_normals(1234, 5, 0, 1 << 28, n_blocks)
    want = oracle_np.sim_normals(1234, 5, 3 * n_blocks)
    np.testing.assert_allclose(z, want, rtol=0, atol=2e-15)
    assert np.mean(z == want) > 0.9995
    W, KH, F, R = oracle_np.zig_tables()
    assert abs(R - 3.8520461503683912) < 1e-11   # the 512-layer solution of the closing condition (edge perturbed < 2^-40)


def test_streams_numpy_vs_cpp(oracle):
    for (seed, chain, it) in [(12, 0, 1), (12, 2, 77), (20261017, 255, 1000)]:
        assert oracle.acc_uniform(seed, chain, it) == oracle_np.acc_uniform(seed, chain, it)
    for N in (2, 3, 9, 40):
        assert oracle.pairs(99, 5, N).tolist() == [list(p) for p in oracle_np.pair_sample(99, 5, N)]


def test_objective_numpy_vs_cpp(oracle):
    cfg = configs.mvnormal(2, 2, n_sim=2000)
    p = np.array([[0.3, -0.2, 1.0, 0.0, -1.5, 2.0, 0.1, -0.1]])
    v, m, _ = oracle.eval_batch(cfg, p)
    v_np, m_np = oracle_np.objective(cfg, p[0])
    np.testing.assert_allclose(v[0], v_np, rtol=1e-11)
    np.testing.assert_allclose(m[0], m_np, rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize("K,T,NI", [(1, 7, 5), (2, 9, 33), (3, 12, 64), (8, 50, 40)])
def test_panel_objective_numpy_vs_cpp(oracle, K, T, NI):
    """the C++ panel simulator (two-pass centred moments over the materialised panel) against numpy"""
    cfg = configs.dynamic_panel(1, 1, K, T, NI)
    lb, ub = configs.panel_box(K)
    rng = np.random.default_rng(K * 100 + T)
    for noseed in (0, 1):
        p = lb + rng.uniform(0.05, 0.95, lb.size) * (ub - lb)
        v, m, st = oracle.eval_batch(cfg, p[None, :], noseed=noseed, rep0=7)
        v_np, m_np = oracle_np.panel_objective(cfg, p, uid=0, rep=7, noseed=noseed)
        assert st[0] == 1
        np.testing.assert_allclose(m[0], m_np, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(v[0], v_np, rtol=1e-10)


def test_panel_full_algorithm_numpy_vs_cpp(oracle):
    K, T, NI = 2, 8, 24
    dm = configs.panel_data_moments(lambda c, p: oracle.eval_batch(c, p), K, T, NI)
    cfg = configs.dynamic_panel(3, 10, K, T, NI, data_mom=dm, sigma_update_steps=3)
    r = oracle.run(cfg, 10)
    w = oracle_np.run(cfg, 10)
    for f in ("accepted", "status", "exchanged", "best_id"):
        np.testing.assert_array_equal(getattr(r.trace, f), w[f], err_msg=f)
    for f in ("value", "prob", "curr_val", "best_val", "params", "sim_moments"):
        np.testing.assert_allclose(getattr(r.trace, f), w[f], rtol=1e-8, atol=1e-11, err_msg=f)


@pytest.mark.parametrize("name", ["c1", "mv_batch1"])
def test_full_algorithm_numpy_vs_cpp(oracle, name):
    if name == "c1":
        cfg, n = configs.c1_serial_normal(25, n_sim=400), 25
    else:
        cfg, n = configs.mvnormal(4, 12, n_params=4, n_sim=300, batch_size=1, sigma_update_steps=3), 12
    r = oracle.run(cfg, n)
    w = oracle_np.run(cfg, n)
    for f in ("accepted", "status", "exchanged", "best_id"):
        np.testing.assert_array_equal(getattr(r.trace, f), w[f], err_msg=f)
    for f in ("value", "prob", "curr_val", "best_val", "params", "sim_moments"):
        np.testing.assert_allclose(getattr(r.trace, f), w[f], rtol=1e-9, atol=1e-12, err_msg=f)
    np.testing.assert_allclose(r.sigma, w["sigma"], rtol=1e-14)


def test_golden_vectors(oracle):
    """committed fixtures (tests/golden/make_golden.py): regression pins of the stream + algorithm"""
    g = np.load(os.path.join(GOLDEN, "bgp_golden.npz"))
    np.testing.assert_array_equal(oracle.normals(1234, 0, 0, 1 << 28, 64).view(np.uint64), g["normals_bits"])
    np.testing.assert_array_equal(oracle.zig_normals(1234, 0, 0, 1 << 28, 4096).view(np.uint64), g["zig_normals_bits"])
    cfg = configs.c1_serial_normal(40)
    tr = oracle.run(cfg, 40).trace
    for f in tr.INT_FIELDS:
        np.testing.assert_array_equal(getattr(tr, f), g["c1_" + f], err_msg=f)
    for f in tr.FLOAT_FIELDS:
        np.testing.assert_allclose(getattr(tr, f), g["c1_" + f], rtol=1e-12, atol=0, err_msg=f)
    cfg = configs.mvnormal(8, 10)
    tr = oracle.run(cfg, 10).trace
    for f in tr.INT_FIELDS:
        np.testing.assert_array_equal(getattr(tr, f), g["mv_" + f], err_msg=f)
    for f in tr.FLOAT_FIELDS:
        np.testing.assert_allclose(getattr(tr, f), g["mv_" + f], rtol=1e-12, atol=0, err_msg=f)
