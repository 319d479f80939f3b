"""Comparison helpers for the parity tests: floats within RTOL (north_star: 1e-6 relative),
integer / boolean bookkeeping bit-exact."""
import numpy as np

from smm_jl_b200._abi import Trace

RTOL = 1e-6  # BASELINE.json north_star: "objective values within 1e-6 relative"
ATOL = 1e-12


def first_divergence(a: Trace, b: Trace):
    """(iteration, chain, field) of the first differing integer/bool entry, or None."""
    for f in Trace.INT_FIELDS:
        x, y = getattr(a, f), getattr(b, f)
        if not np.array_equal(x, y):
            idx = np.argwhere(x != y)[0]
            return int(idx[0]) + 1, int(idx[1]), f, x[tuple(idx)], y[tuple(idx)]
    return None


def assert_trace_parity(got: Trace, want: Trace, rtol=RTOL, atol=ATOL):
    div = first_divergence(got, want)
    assert div is None, f"bookkeeping diverges first at (iter, chain, field, got, want) = {div}"
    for f in Trace.FLOAT_FIELDS:
        x, y = getattr(got, f), getattr(want, f)
        assert x.shape == y.shape, f
        np.testing.assert_allclose(x, y, rtol=rtol, atol=atol, equal_nan=True, err_msg=f)


def max_rel_err(got: Trace, want: Trace) -> float:
    worst = 0.0
    for f in Trace.FLOAT_FIELDS:
        x, y = getattr(got, f), getattr(want, f)
        m = np.isfinite(x) & np.isfinite(y)
        if m.any():
            worst = max(worst, float(np.max(np.abs(x[m] - y[m]) / np.maximum(np.abs(y[m]), 1e-300))))
    return worst
