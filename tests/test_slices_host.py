"""Host logic of doSlices / optSlices / FD_gradient / getSigma / get_stdErrors (slices.jl, econometrics.jl) with the
oracle as the batch evaluator -- the GPU tests (tests/test_gpu_slices.py) run the same calls on the device."""
import numpy as np

from smm_jl_b200 import slices
from tests.slices_common import oracle_evaluator, serial_normal_problem


def test_do_slices_structure_and_linearity(oracle):
    m = serial_normal_problem()
    ev = oracle_evaluator(oracle)
    s = slices.doSlices(m, 7, evaluator=ev)
    assert set(s.res) == {"p1", "p2"} and all(len(v) == 7 for v in s.res.values())
    d = s.get("p1", "value")
    assert np.all(np.diff(d["x"]) > 0) and d["x"][0] == -3.0 and d["x"][-1] == 3.0
    # objfunc_norm with common random numbers: the simulated mean is the parameter plus a constant (ObjExamples.jl:76-79)
    mu = s.get("p1", "mu1")
    np.testing.assert_allclose(np.diff(mu["y"]), np.diff(mu["x"]), rtol=0, atol=1e-9)
    other = s.get("p1", "mu2")                      # slicing p1 leaves the second moment where p2's initial value puts it
    assert np.ptp(other["y"]) < 1e-9
    # the value along the slice is the quadratic it should be: minimal at the grid point closest to the data moment
    assert d["x"][np.argmin(d["y"])] == -1.0
    assert s.p0 == m.initial_value and list(s.m0) == ["mu1", "mu2"]


def test_opt_slices_converges_to_the_data_moments(oracle):
    m = serial_normal_problem()
    out = slices.optSlices(m, 9, tol=1e-3, update=0.4, evaluator=oracle_evaluator(oracle))
    best = out["best"]
    assert abs(best["p"]["p1"] + 1.0) < 0.1 and abs(best["p"]["p2"] - 10.0) < 0.3
    assert best["value"] < 0.05 and out["iterations"] >= 2
    rows = out["history"]
    assert {"iter", "param", "val_idx", "p1", "p2", "value"} <= set(rows[0]) and len(rows) == out["iterations"] * 2 * 9
    assert out["ranges"]["p1"]["ub"] - out["ranges"]["p1"]["lb"] < 6.0      # the search ranges shrank


def test_fd_gradient_is_the_identity_for_the_normal_means(oracle):
    m = serial_normal_problem()
    p = {"p1": 0.3, "p2": -1.0}
    for method in ("forward", "central"):
        J = slices.FD_gradient(m, p, diff_method=method, evaluator=oracle_evaluator(oracle))
        np.testing.assert_allclose(J, np.eye(2), rtol=0, atol=1e-8)
    J = slices.FD_gradient(m, p, use_range=False, step_perc=0.05, evaluator=oracle_evaluator(oracle))
    np.testing.assert_allclose(J, np.eye(2), rtol=0, atol=1e-8)


def test_sigma_and_standard_errors(oracle):
    m = serial_normal_problem()
    p = {"p1": -1.0, "p2": 10.0}
    ev = oracle_evaluator(oracle)
    S = slices.getSigma(m, p, 60, evaluator=ev)
    assert S.shape == (2, 2) and np.allclose(S, S.T) and np.all(np.linalg.eigvalsh(S) > 0)
    # variance of a mean of 10 000 unit normals: 1e-4 (sampling error of a variance from 60 draws: ~18 %)
    assert np.all(np.abs(np.diag(S) / 1e-4 - 1) < 0.6)
    S2 = slices.getSigma(m, p, 60, rep0=1000, evaluator=ev)            # other repetitions: other shocks
    assert not np.allclose(S, S2)
    se = slices.get_stdErrors(m, p, reps=40, evaluator=ev)
    assert list(se) == ["p1", "p2"] and all(0.005 < v < 0.02 for v in se.values())


# ---- the reference's own test/test_slices.jl, with the oracle as the evaluator -------------------------------------
def _problem_2x2():
    from smm_jl_b200 import api
    m = api.MProb()
    api.addSampledParam(m, {"p1": [0.0, -2, 2], "p2": [0.0, -2, 2]})
    api.addMoment(m, {"name": ["mu1", "mu2"], "value": [0.0, 0.0], "weight": [0.7, 0.3]})
    api.addEvalFunc(m, api.objfunc_norm)
    return m


def test_slices_read_and_write(oracle, tmp_path):
    """test_slices.jl:15-37: doSlices(mprob, 3), save, load, get(:p1, :mu1) / get(:p2, :value) survive the round trip"""
    sl = slices.doSlices(_problem_2x2(), 3, evaluator=oracle_evaluator(oracle))
    fn = str(tmp_path / "slices.pkl")
    slices.save(sl, fn)
    disk = slices.load(fn)["s"]
    for p, what in (("p1", "mu1"), ("p2", "value")):
        a, b = sl.get(p, what), disk.get(p, what)
        assert np.array_equal(a["x"], b["x"]) and np.array_equal(a["y"], b["y"])


def test_slices_survive_a_failing_objective(oracle):
    """test_slices.jl:39-60: doSlices(mprob_fail, 30) with Testobj_fails completes; as upstream the failed evaluations are
    stored as status -2 records: value -1.0, no moments (mprob.jl:181-186)"""
    from smm_jl_b200 import api
    m = _problem_2x2()
    m.objfunc = api.Testobj_fails
    sl = slices.doSlices(m, 30, evaluator=oracle_evaluator(oracle))
    assert all(len(v) == 30 for v in sl.res.values())
    assert all(e["value"] == -1.0 and len(e["moments"]) == 0 for v in sl.res.values() for e in v.values())
    assert np.isnan(sl.get("p1", "mu1")["y"]).all() and (sl.get("p1", "value")["y"] == -1.0).all()


def test_naive_coordinate_descent_works(oracle):
    """test_slices.jl:62-65 / Examples.jl:210-230 (snorm_6_taxi): optSlices(mprob, 3, tol, update = 0.4) ends within 0.5
    (Euclidean norm) of the data moments"""
    from smm_jl_b200 import api
    m = api.MProb()
    api.addSampledParam(m, {"p1": [0.2, -3, 3], "p2": [-0.2, -2, 2], "p3": [-0.3, -2, 2], "p4": [-0.4, -2, 2],
                            "p5": [0.3, -2, 2], "p6": [0.4, -2, 2]})
    vals = [-1.0, 1.0, 0.5, -0.5, 0.7, -0.7]
    api.addMoment(m, {"name": [f"mu{i + 1}" for i in range(6)], "value": vals, "weight": [1.0] * 6})
    api.addEvalFunc(m, api.objfunc_norm)
    s = slices.optSlices(m, 3, tol=0.1, update=0.4, evaluator=oracle_evaluator(oracle))
    assert np.linalg.norm(np.array(vals) - np.array(list(s["best"]["p"].values()))) < 0.5
