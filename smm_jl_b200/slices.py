"""Slices and inference on top of the batched objective entry.

The reference's `doSlices` / `optSlices` (src/mopt/slices.jl:114-290) and `FD_gradient` / `getSigma` /
`get_stdErrors` (src/mopt/econometrics.jl:29-145) are nothing but batches of `evaluateObjective` calls: P x npoints grid
points, 1 + P (or 2P) perturbed vectors, `reps` un-seeded repetitions.  Here every such batch is ONE
`smm_bgp_eval_batch` call (`api.evaluateObjectiveBatch`), i.e. one launch of `objective_kernel` over all of them;
only the cheap bookkeeping around it (grid construction, argmin, finite differences, the sandwich formula) runs on
the host, as it does on the master process of the reference.  Names and argument meaning follow the reference.

No CPU path: the evaluator is the device one.  (Tests inject the oracle through the `evaluator` argument.)
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Callable, Dict, List, Optional

import numpy as np

# (api.py re-exports this module's functions, so `api` is imported where it is used, not here)


class Slice:
    """`Slice` (slices.jl:25-38): res[param][value] = {"moments": simMoments, "value": objective}"""

    def __init__(self, p: Dict[str, float], m: Dict[str, dict]):
        self.res: Dict[str, Dict[float, dict]] = {k: {} for k in p}
        self.p0 = OrderedDict(p)
        self.m0 = OrderedDict(m)

    def add(self, p: str, ev: "api.Eval") -> None:
        """add!(s, p, ev) (slices.jl:40-42)"""
        self.res[p][float(ev.params[p])] = {"moments": OrderedDict(ev.simMoments), "value": ev.value}

    def get(self, p: str, m: str = "value") -> Dict[str, np.ndarray]:
        """get(s, p, m) (slices.jl:44-59): x sorted, y = value or one simulated moment along the slice"""
        xs = np.array(list(self.res[p].keys()), dtype=float)
        if m == "value":
            ys = np.array([v["value"] for v in self.res[p].values()], dtype=float)
        else:  # a failed evaluation is stored without moments (mprob.jl:183-186): NaN along the slice
            ys = np.array([v["moments"].get(m, np.nan) for v in self.res[p].values()], dtype=float)
        ix = np.argsort(xs)
        return {"x": xs[ix], "y": ys[ix]}


def save(s: Slice, fname: str) -> None:
    """save(s::Slice, fname) (slices.jl:283-285); a pickle instead of JLD2"""
    import pickle
    with open(fname, "wb") as f:
        pickle.dump({"s": s}, f)


def load(fname: str) -> dict:
    """load(fname) (slices.jl:287-289): {"s": Slice}"""
    import pickle
    with open(fname, "rb") as f:
        return pickle.load(f)


Evaluator = Callable[["api.MProb", List[Dict[str, float]], bool, int], List["api.Eval"]]


def _evaluate(m, plist, noseed=False, rep0=0, evaluator: Optional[Evaluator] = None):
    if evaluator is None:
        from . import api
        evaluator = api.evaluateObjectiveBatch
    return evaluator(m, plist, noseed, rep0)


def doSlices(m: "api.MProb", npoints: int, parallel: bool = False, evaluator: Optional[Evaluator] = None) -> Slice:
    """doSlices(m, npoints, parallel) (slices.jl:242-281): for every sampled parameter, the objective along
    range(lb, ub, npoints) with the other parameters at their initial values.  All P x npoints evaluations travel in
    one batch (`parallel` is accepted for signature compatibility: the batch is the parallelism).  A failing objective
    does not stop the slice: as upstream, `evaluateObjective` turns the failure into an Eval with status -2, value -1.0
    and no moments (mprob.jl:181-186), and that record is stored like any other (the `<: Exception` branch of
    slices.jl:270 never fires, because the exception was already caught)."""
    res = Slice(m.initial_value, m.moments)
    plist, owner = [], []
    for pp, bb in m.params_to_sample.items():
        for pval in np.linspace(bb["lb"], bb["ub"], npoints):
            p = OrderedDict(m.initial_value)
            p[pp] = float(pval)
            plist.append(p)
            owner.append(pp)
    for pp, ev in zip(owner, _evaluate(m, plist, evaluator=evaluator)):
        res.add(pp, ev)
    return res


def optSlices(m: "api.MProb", npoints: int, parallel: bool = False, tol: float = 1e-5, update: Optional[float] = None,
              max_cycles: int = 1000, evaluator: Optional[Evaluator] = None) -> dict:
    """optSlices (slices.jl:114-240): naive cyclic coordinate descent.  Within a cycle the parameters are searched one
    after the other on a grid of `npoints`, each starting from the best point found so far; the search ranges shrink
    around the best point by the factor `update` after every cycle; stop when the cycle moved the point by less than
    `tol` (Euclidean norm).  One batched launch per (cycle, parameter).  Deliberate deviation: failed evaluations
    (status < 0, value -1.0) are skipped when the best grid point is chosen -- upstream compares their -1.0 like any
    value (slices.jl:188-200) and would walk towards the failures.  Returns {"best": {"p", "value"},
    "history": rows of {iter, param, val_idx, <params>, value}} (the reference also writes a JLD2 file)."""
    ranges = OrderedDict((k, dict(v)) for k, v in m.params_to_sample.items())
    bestp = OrderedDict(m.initial_value)
    dvec = OrderedDict((k, np.inf) for k in bestp)
    dout: dict = {"history": []}
    delta, it = np.inf, 0
    while delta > tol and it < max_cycles:
        it += 1
        for pp, bb in ranges.items():
            cur_param = OrderedDict(bestp)
            plist = []
            for pval in np.linspace(bb["lb"], bb["ub"], npoints):
                p = OrderedDict(cur_param)
                p[pp] = float(pval)
                plist.append(p)
            vv = _evaluate(m, plist, evaluator=evaluator)
            minv = np.inf
            bestp = OrderedDict(cur_param)
            for iv, ev in enumerate(vv, start=1):
                if ev.status < 0:
                    dout["history"].append({"iter": it, "param": pp, "val_idx": iv, "value": float("nan")})
                    continue
                row = {"iter": it, "param": pp, "val_idx": iv}
                row.update({k: float(v) for k, v in ev.params.items()})
                row["value"] = ev.value
                dout["history"].append(row)
                if np.isfinite(ev.value) and ev.value < minv:
                    minv = ev.value
                    bestp = OrderedDict(ev.params)
                    dout["best"] = {"p": OrderedDict(ev.params), "value": ev.value}
            dvec[pp] = cur_param[pp] - bestp[pp]
        if update is not None:  # maintain range boundaries (slices.jl:221-228)
            for k, v in bestp.items():
                if k not in ranges:
                    continue
                r = (ranges[k]["ub"] - ranges[k]["lb"]) / 2
                ranges[k]["lb"] = max(v - update * r, ranges[k]["lb"])
                ranges[k]["ub"] = min(v + update * r, ranges[k]["ub"])
        delta = float(np.linalg.norm([dvec[k] for k in ranges]))
    dout["iterations"] = it
    dout["ranges"] = ranges
    return dout


def range_length(m: "api.MProb") -> Dict[str, float]:
    """ub - lb of every sampled parameter (mprob.jl `range_length`)"""
    return OrderedDict((k, v["ub"] - v["lb"]) for k, v in m.params_to_sample.items())


def FD_gradient(m: "api.MProb", p: Dict[str, float], step_perc: float = 0.01, diff_method: str = "forward",
                use_range: bool = True, evaluator: Optional[Evaluator] = None) -> np.ndarray:
    """FD_gradient (econometrics.jl:29-88): finite-difference Jacobian of the simulated moments, a (k, n) matrix for k
    parameters (rows in the order of `p`) and n moments.  `step_perc` of the parameter range (`use_range`) or of the
    parameter value; "forward" or "central" differences.  1 + k (forward) or 1 + 2k (central) evaluations, one batch."""
    if diff_method not in ("forward", "central"):
        raise ValueError("only :central and :forward implemented")  # econometrics.jl:69
    mnames = list(m.moments.keys())
    rs = range_length(m)
    keys = list(p.keys())
    plist: List[Dict[str, float]] = [OrderedDict(p)]
    hs = []
    for k in keys:
        h = (rs[k] if use_range else p[k]) * step_perc
        hs.append(h)
        if diff_method == "forward":
            q = OrderedDict(p)
            q[k] = p[k] + h
            plist.append(q)
        else:
            for sgn in (+0.5, -0.5):
                q = OrderedDict(p)
                q[k] = p[k] + sgn * h
                plist.append(q)
    evs = _evaluate(m, plist, evaluator=evaluator)
    g = lambda ev: np.array([ev.simMoments[n] for n in mnames], dtype=float)
    gp = g(evs[0])
    D = np.zeros((len(keys), len(mnames)))
    for row, h in enumerate(hs):
        if diff_method == "forward":
            D[row] = (g(evs[1 + row]) - gp) / h
        else:
            D[row] = (g(evs[1 + 2 * row]) - g(evs[2 + 2 * row])) / h
    return D


def getSigma(m: "api.MProb", p: Dict[str, float], reps: int, rep0: int = 0,
             evaluator: Optional[Evaluator] = None) -> np.ndarray:
    """getSigma (econometrics.jl:125-145): var-cov matrix of the simulated moments over `reps` evaluations at `p` with
    UN-seeded shocks (`noseed`: repetition r draws its own stream, indexed by rep0 + r).  One batch."""
    mnames = list(m.moments.keys())
    evs = _evaluate(m, [OrderedDict(p) for _ in range(reps)], noseed=True, rep0=rep0, evaluator=evaluator)
    d = np.array([[ev.simMoments[n] for n in mnames] for ev in evs], dtype=float)
    return np.cov(d, rowvar=False)


def get_stdErrors(m: "api.MProb", p: Dict[str, float], reps: int = 300,
                  evaluator: Optional[Evaluator] = None) -> Dict[str, float]:
    """get_stdErrors (econometrics.jl:91-116): sandwich formula S = (J W J')^-1 (J W Sigma W J') (J W J')^-1 with
    Sigma from `getSigma`, J from `FD_gradient`, W = diag(weights)."""
    Sigma = np.atleast_2d(getSigma(m, p, reps, evaluator=evaluator))
    J = FD_gradient(m, p, evaluator=evaluator)
    W = np.diag([v["weight"] for v in m.moments.values()])
    A = np.linalg.pinv(J @ W @ J.T)
    SE = A @ (J @ W @ Sigma @ W @ J.T) @ A
    return OrderedDict(zip(p.keys(), np.sqrt(np.diag(SE))))
