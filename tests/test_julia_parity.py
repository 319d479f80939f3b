"""The route to a REFERENCE pin of parity: julia/parity_harness.jl runs the real SMM.jl
`computeNextIteration!` on the four injected streams dumped by tools/dump_streams.py (committed under
tests/golden/julia/<case>/) and writes `<case>/julia_trace/`.  When that directory exists the oracle (here) and the CUDA
path (-m gpu) are compared with it: bookkeeping bit-exact, floats within 1e-6 relative (BASELINE.json north_star).  Julia
is not installable in the build image, so until someone with Julia commits the traces those comparisons are skipped --
what always runs is (a) the dumps are exactly what the stream definitions give, (b) the comparison path itself, fed with
a trace written in the harness's format by the independent numpy re-derivation of the algorithm (oracle/oracle_np.py)."""
import os

import numpy as np
import pytest

from tests import julia_trace
from tests.parity import assert_trace_parity
from tools import dump_streams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "julia")
CASES = list(dump_streams.cases().keys())
HOWTO = ("no Julia trace for this case: run `julia --project=<SMM.jl checkout> julia/parity_harness.jl "
         "tests/golden/julia/{case}` (Julia is not installable in the build image) and commit {case}/julia_trace/")


@pytest.mark.parametrize("case", CASES)
def test_committed_stream_dumps_are_the_stream_definitions(case, tmp_path, oracle):
    """tools/dump_streams.py is deterministic and the committed files are its output: the harness reads exactly the
    streams the kernels and the oracle use"""
    cfg, objective = dump_streams.cases()[case]
    out = os.path.join(tmp_path, case)
    dump_streams.dump(case, cfg, objective, out)
    names = sorted(os.listdir(out))
    assert names == sorted(n for n in os.listdir(os.path.join(GOLDEN, case)) if n != "julia_trace")
    for n in names:
        with open(os.path.join(out, n), "rb") as a, open(os.path.join(GOLDEN, case, n), "rb") as b:
            assert a.read() == b.read(), n
    # and an independent derivation of a few elements (numpy / pure Python Philox + transforms)
    from oracle import oracle_np as onp
    N, P, I, S = cfg.n_chains, cfg.n_params, cfg.max_iter, cfg.n_sim
    A = dump_streams.N_ATTEMPTS
    zsim = np.fromfile(os.path.join(out, "zsim.f64")).reshape(P, S)
    assert np.array_equal(zsim[P - 1], onp.sim_normals(cfg.seed_sim, P - 1, S))
    zprop = np.fromfile(os.path.join(out, "zprop.f64")).reshape(N, I, A, P)
    for (c, it, a, k) in ((0, 2, 0, 0), (N - 1, I, A - 1, P - 1), (1, 7, 3, 1)):
        assert zprop[c, it - 1, a, k] == pytest.approx(onp.prop_normal(cfg.seed_algo, c, it, a, k), rel=1e-14)
    uacc = np.fromfile(os.path.join(out, "uacc.f64")).reshape(N, I)
    assert uacc[N - 1, I - 1] == onp.acc_uniform(cfg.seed_algo, N - 1, I)
    pairs = np.fromfile(os.path.join(out, "pairs.i32"), dtype=np.int32).reshape(I, -1, 2)
    assert pairs[I - 1].tolist() == [[i + 1, j + 1] for i, j in onp.pair_sample(cfg.seed_algo, I, N)]
    assert not pairs[0].any()


def test_comparison_path_on_a_trace_in_the_harness_format(tmp_path, oracle):
    """the reader and the comparison, end to end, on a trace written in the harness's layout by the independent numpy
    restatement of the algorithm (a stand-in for the Julia run, 30 iterations of the 8-chain case)"""
    from oracle import oracle_np as onp
    from smm_jl_b200._abi import Trace
    cfg, _ = dump_streams.cases()["mvnormal_8chains"]
    n = 30
    r = onp.run(cfg, n)
    tr = Trace(n, cfg.n_chains, cfg.n_params, cfg.n_moments)
    for f in Trace.FLOAT_FIELDS + Trace.INT_FIELDS:
        setattr(tr, f, np.asarray(r[f]).astype(getattr(tr, f).dtype))
    d = os.path.join(tmp_path, "julia_trace")
    julia_trace.store(d, tr, r["sigma"], np.zeros(cfg.n_chains), julia="numpy stand-in")
    got, sigma, _ = julia_trace.load(d, cfg.n_chains, n, cfg.n_params, cfg.n_moments)
    ref = oracle.run(cfg, n)
    assert_trace_parity(ref.trace, got)
    np.testing.assert_allclose(ref.sigma, sigma, rtol=1e-12)


def _julia_trace(case):
    d = os.path.join(GOLDEN, case, "julia_trace")
    if not os.path.exists(os.path.join(d, "value.f64")):
        pytest.skip(HOWTO.format(case=case))
    cfg, _ = dump_streams.cases()[case]
    return cfg, julia_trace.load(d, cfg.n_chains, cfg.max_iter, cfg.n_params, cfg.n_moments)


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_the_real_smm_jl(case, oracle):
    cfg, (want, sigma, acc) = _julia_trace(case)
    ref = oracle.run(cfg, cfg.max_iter)
    assert_trace_parity(ref.trace, want)
    np.testing.assert_allclose(ref.sigma, sigma, rtol=1e-12)
    np.testing.assert_allclose(ref.accept_rate, acc, rtol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("mode", [0, 2])
def test_cuda_path_matches_the_real_smm_jl(case, mode, smm):
    import copy
    cfg, (want, sigma, acc) = _julia_trace(case)
    c = copy.copy(cfg)
    c.exchange_mode = mode
    with smm.BGPHandle(c) as h:
        h.step(c.max_iter)
        got = h.read_trace(1, c.max_iter)
        s, a = h.chain_state()
    assert_trace_parity(got, want)
    np.testing.assert_allclose(s, sigma, rtol=1e-12)
    np.testing.assert_allclose(a, acc, rtol=1e-12)


def test_stream_files_as_the_harness_indexes_them(oracle, monkeypatch):
    """The harness reads the dumped arrays column-major with the dimensions reversed and indexes them 1-based as
    ZPROP[k, attempt, iter, chain], UACC[iter, chain], PAIRS[1:2, t, iter], ZSIM[s, k].  Here the numpy restatement of the
    algorithm runs on the FILES through exactly those index expressions (Fortran-order views, 1-based arithmetic) and
    must reproduce the oracle's trace: a dimension or off-by-one mistake in the dump or in the harness's indexing shows
    up here, without Julia."""
    from oracle import oracle_np as onp
    case = "mvnormal_8chains"
    cfg, _ = dump_streams.cases()[case]
    d = os.path.join(GOLDEN, case)
    meta = julia_trace.read_meta(os.path.join(d, "meta.txt"))
    N, P, S, I, A, NS = (int(meta[k]) for k in ("n_chains", "n_params", "n_sim", "n_iter", "n_attempts", "n_pairs"))
    ZSIM = np.fromfile(os.path.join(d, "zsim.f64")).reshape((S, P), order="F")
    ZPROP = np.fromfile(os.path.join(d, "zprop.f64")).reshape((P, A, I, N), order="F")
    UACC = np.fromfile(os.path.join(d, "uacc.f64")).reshape((I, N), order="F")
    PAIRS = np.fromfile(os.path.join(d, "pairs.i32"), dtype=np.int32).reshape((2, NS, I), order="F")
    j = lambda i: i - 1                                   # Julia's 1-based index -> numpy
    monkeypatch.setattr(onp, "prop_normal", lambda seed, c, it, a, k: float(ZPROP[j(k + 1), j(a + 1), j(it), j(c + 1)]))
    monkeypatch.setattr(onp, "acc_uniform", lambda seed, c, it: float(UACC[j(it), j(c + 1)]))
    monkeypatch.setattr(onp, "pair_sample", lambda seed, it, n: [(int(PAIRS[0, t, j(it)]) - 1, int(PAIRS[1, t, j(it)]) - 1) for t in range(NS)])
    monkeypatch.setattr(onp, "sim_normals", lambda seed, k, s, *a, **kw: ZSIM[:, j(k + 1)].copy())
    n = 40
    r = onp.run(cfg, n)
    ref = oracle.run(cfg, n)
    for f in ("accepted", "status", "exchanged", "best_id"):
        assert np.array_equal(np.asarray(r[f]), getattr(ref.trace, f)), f
    assert np.array_equal(np.asarray(r["params"]), ref.trace.params)
    np.testing.assert_allclose(np.asarray(r["value"]), ref.trace.value, rtol=1e-9)
    np.testing.assert_allclose(r["sigma"], ref.sigma, rtol=1e-12)
