"""doSlices / FD_gradient / getSigma on the device (one objective_kernel launch per batch) against the same calls
evaluated by the oracle."""
import numpy as np
import pytest

from smm_jl_b200 import api
from tests.slices_common import oracle_evaluator, serial_normal_problem

pytestmark = pytest.mark.gpu


def test_do_slices_on_the_device(oracle):
    m = serial_normal_problem()
    s = api.doSlices(m, 9)
    r = api.doSlices(m, 9, evaluator=oracle_evaluator(oracle))
    for p in ("p1", "p2"):
        for what in ("value", "mu1", "mu2"):
            a, b = s.get(p, what), r.get(p, what)
            np.testing.assert_array_equal(a["x"], b["x"])
            np.testing.assert_allclose(a["y"], b["y"], rtol=1e-9, atol=1e-12)


def test_gradient_and_sigma_on_the_device(oracle):
    m = serial_normal_problem()
    p = {"p1": 0.3, "p2": -1.0}
    J = api.FD_gradient(m, p)
    np.testing.assert_allclose(J, api.FD_gradient(m, p, evaluator=oracle_evaluator(oracle)), rtol=0, atol=1e-9)
    np.testing.assert_allclose(J, np.eye(2), rtol=0, atol=1e-8)
    S = api.getSigma(m, p, 40)
    np.testing.assert_allclose(S, api.getSigma(m, p, 40, evaluator=oracle_evaluator(oracle)), rtol=1e-6, atol=1e-12)
    se = api.get_stdErrors(m, p, reps=40)
    assert all(0.005 < v < 0.02 for v in se.values())


def test_opt_slices_on_the_device():
    m = serial_normal_problem()
    out = api.optSlices(m, 9, tol=1e-3, update=0.4)
    assert abs(out["best"]["p"]["p1"] + 1.0) < 0.1 and abs(out["best"]["p"]["p2"] - 10.0) < 0.3
