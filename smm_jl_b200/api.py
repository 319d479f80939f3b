"""Host-side mirror of SMM.jl's user surface for the BGP hot path (Julia is not available in this
image, so the executable host layer above the C ABI is Python; julia/SMMB200.jl is the same thing
as a `ccall` shim).  Names and argument meaning follow the reference (Julia's trailing `!` dropped):

    MProb, addParam, addSampledParam, addMoment, addEvalFunc, addEvalFuncOpts   (mprob.jl:29-166)
    Eval + param/paramd/dataMoment/dataMomentW/setMoments/setValue               (Eval.jl:20-238)
    evaluateObjective(m, p; noseed)                                              (mprob.jl:175-205)
    MAlgoBGP(m, opts), run, restart, save, readMalgo, summary                    (AlgoBGP.jl:497-550,
                                                                                 AlgoAbstract.jl:27-102)
    BGPChain fields + history/best/mean/median/CI/allAccepted/params            (AlgoBGP.jl:42-206)

All computation happens in libsmm_b200.so on the GPU; there is no CPU path in this module.
"""
from __future__ import annotations

import pickle
import time as _time
from collections import OrderedDict
from typing import Any, Callable, Dict, Optional

import numpy as np

from . import _lib
from ._abi import (BGPConfig, SMM_OBJ_FAILS, SMM_OBJ_NORM, SMM_OBJ_NORM_MV, SMM_OBJ_NORM_SLOW, SMM_OBJ_PANEL, Trace)
from .configs import temperature_ladder


# ---------------------------------------------------------------------------------------------
# built-in objective functions: markers that select a device simulator
# ---------------------------------------------------------------------------------------------
class DeviceObjective:
    def __init__(self, name: str, objective_id: int, doc: str):
        self.__name__ = name
        self.objective_id = objective_id
        self.__doc__ = doc

    def __call__(self, ev: "Eval", **kw) -> "Eval":
        """`objfunc(ev)`: one evaluation through the device path (a batch of one)."""
        if ev._mprob is None:
            raise RuntimeError(f"{self.__name__}: this Eval is not attached to an MProb; use evaluateObjective(m, p)")
        m = ev._mprob
        saved = m.objfunc
        try:
            m.objfunc = self
            return evaluateObjective(m, ev, **kw)
        finally:
            m.objfunc = saved

    def __repr__(self):
        return f"<device objective {self.__name__}>"


objfunc_norm = DeviceObjective("objfunc_norm", SMM_OBJ_NORM, "ObjExamples.jl:59-116")
objfunc_norm_slow = DeviceObjective("objfunc_norm_slow", SMM_OBJ_NORM_SLOW, "ObjExamples.jl:124-184")
objfunc_norm_mv = DeviceObjective("objfunc_norm_mv", SMM_OBJ_NORM_MV, "means + variances (SURVEY.md 8d)")
objfunc_panel = DeviceObjective("objfunc_panel", SMM_OBJ_PANEL, "dynamic panel (SURVEY.md 8d)")
Testobj_fails = DeviceObjective("Testobj_fails", SMM_OBJ_FAILS, "ObjExamples.jl:27-32")


# ---------------------------------------------------------------------------------------------
# MProb (mprob.jl:29-166)
# ---------------------------------------------------------------------------------------------
class MProb:
    def __init__(self):
        self.initial_value: "OrderedDict[str, float]" = OrderedDict()
        self.params_to_sample: "OrderedDict[str, dict]" = OrderedDict()
        self.objfunc: Any = None
        self.objfunc_opts: Dict[str, Any] = {}
        self.moments: "OrderedDict[str, dict]" = OrderedDict()

    def __repr__(self):
        return (f"MProb Object:\n==============\n\nParameters to sample:\n{dict(self.params_to_sample)}\n"
                f"Moment Table:\n{dict(self.moments)}\n\nobjective function: {self.objfunc}\n")


def addParam(m: MProb, name, init=None):
    """addParam!(m, name, init) / addParam!(m, dict) (mprob.jl:60-75)"""
    if isinstance(name, dict):
        for k, v in name.items():
            m.initial_value[str(k)] = v
    else:
        m.initial_value[str(name)] = init
    return m


def addSampledParam(m: MProb, name, init=None, lb=None, ub=None):
    """addSampledParam!(m, name, init, lb, ub) / addSampledParam!(m, dict of (init, lb, ub)) (mprob.jl:81-98)"""
    if isinstance(name, dict):
        for k, v in name.items():
            addSampledParam(m, k, v[0], v[1], v[2])
        return m
    assert ub > lb, "ub > lb"  # mprob.jl:82
    m.initial_value[str(name)] = float(init)
    m.params_to_sample[str(name)] = {"lb": float(lb), "ub": float(ub)}
    return m


def addMoment(m: MProb, name, value=None, weight=1.0):
    """addMoment!(m, name, value[, weight]) / dict of {value, weight} / table with name,value,weight (mprob.jl:123-155)"""
    if hasattr(name, "columns") or (isinstance(name, dict) and "name" in name and "value" in name):  # DataFrame-like
        names, values = list(name["name"]), list(name["value"])
        weights = list(name["weight"]) if "weight" in name else [1.0] * len(names)
        for n, v, w in zip(names, values, weights):
            addMoment(m, n, v, w)
        return m
    if isinstance(name, dict):
        for k, v in name.items():
            addMoment(m, k, v["value"], v.get("weight", 1.0))
        return m
    m.moments[str(name)] = {"value": float(value), "weight": float(weight)}
    return m


def addEvalFunc(m: MProb, f):
    """addEvalFunc!(m, f) (mprob.jl:159-161).  Only the built-in device objectives can run on the GPU."""
    m.objfunc = f
    return m


def addEvalFuncOpts(m: MProb, d: dict):
    m.objfunc_opts = dict(d)
    return m


def ps_names(m: MProb):
    return list(m.initial_value.keys())


def ps2s_names(m: MProb):
    return list(m.params_to_sample.keys())


def ms_names(m: MProb):
    return list(m.moments.keys())


def mapto_01(p, lb, ub):
    """mprob.jl:246-253"""
    p = np.asarray(list(p.values()) if isinstance(p, dict) else p, dtype=float)
    return (p - np.asarray(lb, float)) / (np.asarray(ub, float) - np.asarray(lb, float))


def mapto_ab(p, lb, ub):
    """mprob.jl:270-272"""
    return np.asarray(p, float) * (np.asarray(ub, float) - np.asarray(lb, float)) + np.asarray(lb, float)


# ---------------------------------------------------------------------------------------------
# Eval (Eval.jl:20-238)
# ---------------------------------------------------------------------------------------------
class Eval:
    def __init__(self, mprob: Optional[MProb] = None, p: Optional[dict] = None):
        self.value = -1.0
        self.time = _time.time()
        self.status = -1
        self.params: "OrderedDict[str, float]" = OrderedDict()
        self.simMoments: "OrderedDict[str, float]" = OrderedDict()
        self.dataMoments: "OrderedDict[str, float]" = OrderedDict()
        self.dataMomentsW: "OrderedDict[str, float]" = OrderedDict()
        self.prob = 0.0
        self.accepted = False
        self.options: Dict[str, Any] = {}
        self._mprob = mprob
        if mprob is not None:
            for k, v in mprob.moments.items():
                self.dataMoments[k] = v["value"]
                self.dataMomentsW[k] = v["weight"]
            src = p if p is not None else mprob.initial_value
            for k, v in src.items():
                self.params[str(k)] = float(v)
        elif p is not None:
            for k, v in p.items():
                self.params[str(k)] = float(v)

    def __eq__(self, other):  # Eval.jl:157-166
        return (isinstance(other, Eval) and self.value == other.value and self.status == other.status
                and self.params == other.params and self.simMoments == other.simMoments
                and self.dataMoments == other.dataMoments and self.accepted == other.accepted)

    def __repr__(self):
        return (f"Eval Object:\n============\n\nObjective function value: {self.value}\nEvaluation Status: "
                f"{self.status}\nParameters:\n{list(self.params)}\nMoments:\n{list(self.dataMoments)}\n")


def start(ev: Eval):
    ev.time = _time.time()


def finish(ev: Eval):
    ev.time = _time.time() - ev.time


def param(ev: Eval, key=None):
    if key is None:
        return np.array(list(ev.params.values()), dtype=float)
    if isinstance(key, (list, tuple)):
        return np.array([ev.params[str(k)] for k in key], dtype=float)
    return ev.params[str(key)]


def paramd(ev: Eval):
    return ev.params


def dataMoment(ev: Eval, key=None):
    if key is None:
        return np.array(list(ev.dataMoments.values()), dtype=float)
    if isinstance(key, (list, tuple)):
        return np.array([ev.dataMoments[str(k)] for k in key], dtype=float)
    return ev.dataMoments[str(key)]


def dataMomentd(ev: Eval):
    return ev.dataMoments


def dataMomentW(ev: Eval, key=None):
    if key is None:
        return np.array(list(ev.dataMomentsW.values()), dtype=float)
    if isinstance(key, (list, tuple)):
        return np.array([ev.dataMomentsW[str(k)] for k in key], dtype=float)
    return ev.dataMomentsW[str(key)]


def dataMomentWd(ev: Eval):
    return ev.dataMomentsW


def setValue(ev: Eval, value: float):
    ev.value = float(value)


def setMoments(ev: Eval, k, value=None):
    if isinstance(k, dict):
        for kk, v in k.items():
            ev.simMoments[str(kk)] = float(v)
    elif isinstance(k, (list, tuple)):
        for kk, v in zip(k, value):
            ev.simMoments[str(kk)] = float(v)
    else:
        ev.simMoments[str(k)] = float(value)


# ---------------------------------------------------------------------------------------------
# MProb -> C config
# ---------------------------------------------------------------------------------------------
def _objective_id(m: MProb) -> int:
    f = m.objfunc
    if isinstance(f, DeviceObjective):
        return f.objective_id
    raise NotImplementedError(
        f"objective {f!r} is an arbitrary host function; only the built-in device objectives "
        "(objfunc_norm, objfunc_norm_slow, objfunc_norm_mv, objfunc_panel, Testobj_fails) run on the GPU "
        "and this package has no CPU fallback")


def _check_all_sampled(m: MProb):
    if list(m.initial_value.keys()) != list(m.params_to_sample.keys()):
        # upstream: proposal() broadcasts paramd(ev_old) against lb/ub of params_to_sample (AlgoBGP.jl:430-436)
        raise ValueError("every parameter must be a sampled parameter, in the same order (AlgoBGP.jl:430-436)")


def _base_config(m: MProb, n_chains: int, max_iter: int, opts: dict) -> BGPConfig:
    _check_all_sampled(m)
    names = list(m.params_to_sample.keys())
    N = n_chains
    if N > 1:
        temps = temperature_ladder(N, opts.get("maxtemp", 1.0))  # AlgoBGP.jl:508
    else:
        temps = np.ones(1)
    oo = m.objfunc_opts
    return BGPConfig(
        lb=[m.params_to_sample[k]["lb"] for k in names],
        ub=[m.params_to_sample[k]["ub"] for k in names],
        init=[m.initial_value[k] for k in names],
        data_mom=[v["value"] for v in m.moments.values()],
        data_w=[v["weight"] for v in m.moments.values()],
        n_chains=N, max_iter=max_iter,
        sigma0=opts.get("sigma", 0.05) * temps,
        acc_tuner=np.asarray(opts.get("acc_tuners", [2.0] * N), dtype=float)[:N],
        min_improve=np.asarray(opts.get("min_improve", [0.5] * N), dtype=float)[:N],
        objective_id=_objective_id(m),
        n_sim=int(oo.get("n_sim", 10000)), seed_sim=int(oo.get("seed", 1234)), noseed=int(bool(oo.get("noseed", False))),
        slow_seconds=float(oo.get("slow_seconds", 0.1)),
        panel_T=int(oo.get("panel_T", 0)), panel_N=int(oo.get("panel_N", 0)), panel_K=int(oo.get("panel_K", 0)),
        sigma_update_steps=int(opts.get("sigma_update_steps", 10)),
        sigma_adjust_by=float(opts.get("sigma_adjust_by", 0.01)),
        smpl_iters=int(opts.get("smpl_iters", 1000)),
        batch_size=int(opts.get("batch_size", len(names))),
        seed_algo=int(opts.get("seed", 20261017)),
        device=int(opts.get("device", 0)), world_size=int(opts.get("world_size", 1)), rank=int(opts.get("rank", 0)),
        nccl_id=opts.get("nccl_id", b""), exchange_mode=_exchange_mode(m, opts, len(names)),
        n_split=int(opts.get("n_split", 0)),
    )


def _exchange_mode(m: MProb, opts: dict, n_params: int) -> int:
    """opts["exchange_mode"], or the fastest mode the shape allows: on one GPU the barrier-free persistent kernel with a
    completion counter (2); on several the same kernel with the flag-in-data hand-over (3: value, sigma and proposal
    centre of every chain travel to every peer as {payload | iteration tag} words, no system fence and no counter round
    trip on the critical path; +4-7 % at 2 and 4 GPUs, 2 % slower than mode 2 on one); the multi-launch path (0) for
    the panel objective and for more than 32 parameters (include/smm_b200.h).  A persistent mode that does not fit the
    shape (SMM_E_UNSUPPORTED_SHAPE at create time) falls back to 0 in MAlgoBGP._handle."""
    if "exchange_mode" in opts:
        return int(opts["exchange_mode"])
    if _objective_id(m) == SMM_OBJ_PANEL or n_params > 32:
        return 0
    return 3 if int(opts.get("world_size", 1)) > 1 else 2


def evaluateObjective(m: MProb, p, noseed: bool = False, rep: int = 0) -> Eval:
    """evaluateObjective(m, p; noseed) / evaluateObjective(m, ev) (mprob.jl:175-205), on the device."""
    ev = p if isinstance(p, Eval) else Eval(m, p)
    ev._mprob = m
    if noseed:
        ev.options["noseed"] = True
    evs = evaluateObjectiveBatch(m, [ev.params], noseed=bool(ev.options.get("noseed", False)), rep0=rep)
    out = evs[0]
    ev.value, ev.status, ev.simMoments, ev.time = out.value, out.status, out.simMoments, out.time
    return ev


def evaluateObjectiveBatch(m: MProb, plist, noseed: bool = False, rep0: int = 0) -> "list[Eval]":
    """Many `evaluateObjective` calls in one launch (what doSlices / getSigma / FD_gradient loop over)."""
    names = list(m.params_to_sample.keys())
    cfg = _base_config(m, 1, 1, {"exchange_mode": 0})      # a one-chain handle that only carries the problem definition
    P = np.array([[float(p[k]) for k in names] for p in plist], dtype=float)
    t0 = _time.time()
    with _lib.BGPHandle(cfg) as h:
        value, mom, status = h.eval_batch(P, noseed=int(noseed), rep0=rep0)
    dt = (_time.time() - t0) / max(len(plist), 1)
    out = []
    for b, p in enumerate(plist):
        ev = Eval(m, OrderedDict((k, float(p[k])) for k in names))
        ev.value, ev.status, ev.time = float(value[b]), int(status[b]), dt
        if status[b] >= 0:
            for k, v in zip(m.moments.keys(), mom[b]):
                ev.simMoments[k] = float(v)
        if noseed:
            ev.options["noseed"] = True
        out.append(ev)
    return out


# ---------------------------------------------------------------------------------------------
# BGPChain: a view over the SoA trace with the reference's field names (AlgoBGP.jl:42-61)
# ---------------------------------------------------------------------------------------------
class _EvalList:
    """Lazy `evals::Array{Eval}`: records are materialised on access (SURVEY.md 7 'trace memory')."""

    def __init__(self, chain: "BGPChain"):
        self._c = chain

    def __len__(self):
        return self._c.n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if isinstance(i, (np.ndarray, list)):
            idx = np.asarray(i)
            if idx.dtype == bool:
                idx = np.nonzero(idx)[0]
            return [self[int(j)] for j in idx]
        c = self._c
        if i < 0:
            i += len(self)
        if not 0 <= i < c.iter:
            raise IndexError("Eval slot not computed yet (#undef upstream)")
        ev = Eval(c.m, OrderedDict(zip(c._pnames, c._params[i])))
        ev.value, ev.prob = float(c._value[i]), float(c._prob[i])
        ev.status, ev.accepted = int(c._status[i]), bool(c.accepted[i])
        if ev.status >= 0 or not np.isnan(c._mom[i]).all():
            for k, v in zip(c._mnames, c._mom[i]):
                ev.simMoments[k] = float(v)
        return ev


class BGPChain:
    def __init__(self, id: int, m: MProb, n: int, tr: Trace, col: int, it: int, sigma: float, accept_rate: float,
                 probs_acc: np.ndarray, cfg: BGPConfig):
        self.id, self.m, self.n, self.iter = id, m, n, it
        self._pnames, self._mnames = list(m.params_to_sample.keys()), list(m.moments.keys())
        pad = n - tr.n

        def col1(a, fill):
            return np.concatenate([a[:, col], np.full(pad, fill, dtype=a.dtype)]) if pad else a[:, col].copy()

        self._value, self._prob = col1(tr.value, np.nan), col1(tr.prob, np.nan)
        self._status = col1(tr.status, 0)
        self._params = tr.params[:, col, :].copy()      # copies: the trace buffer may be pinned memory that is
        self._mom = tr.sim_moments[:, col, :].copy()     # recycled when the algo is closed
        self.best_id = col1(tr.best_id, -1).astype(int)
        self.best_val = col1(tr.best_val, np.inf)
        self.curr_val = col1(tr.curr_val, np.inf)
        self.accepted = col1(tr.accepted, 0).astype(bool)
        self.exchanged = col1(tr.exchanged, 0).astype(int)
        self.probs_acc = probs_acc
        self.accept_rate, self.sigma = float(accept_rate), float(sigma)
        self.acc_tuner = float(np.asarray(cfg.acc_tuner).reshape(-1)[id - 1])
        self.min_improve = float(np.asarray(cfg.min_improve).reshape(-1)[id - 1])
        self.sigma_update_steps, self.sigma_adjust_by = cfg.sigma_update_steps, cfg.sigma_adjust_by
        self.smpl_iters = cfg.smpl_iters
        bs = cfg.batch_size or cfg.n_params
        self.batches = [range(lo + 1, lo + bs + 1) for lo in range(0, cfg.n_params, bs)]
        self.evals = _EvalList(self)


def allAccepted(c: BGPChain):
    return c.evals[c.accepted[: c.iter]]


def params(c: BGPChain, accepted_only: bool = True) -> "Dict[str, np.ndarray]":
    """AlgoBGP.jl:120-132"""
    sel = c.accepted[: c.iter] if accepted_only else np.ones(c.iter, dtype=bool)
    return {k: c._params[: c.iter][sel, j].copy() for j, k in enumerate(c._pnames)}


def history(c: BGPChain) -> "Dict[str, np.ndarray]":
    """history(c) (AlgoBGP.jl:138-160): columns iter,value,accepted,curr_val,best_val,prob,exchanged,params...
    Returned as an ordered dict of columns (a pandas DataFrame if pandas is importable)."""
    n = c.iter
    cols = OrderedDict()
    cols["iter"] = np.arange(1, n + 1)
    cols["value"] = c._value[:n]
    cols["accepted"] = c.accepted[:n]
    cols["curr_val"] = c.curr_val[:n]
    cols["best_val"] = c.best_val[:n]
    cols["prob"] = c._prob[:n]
    cols["exchanged"] = c.exchanged[:n]
    for j, k in enumerate(c._pnames):
        cols[k] = c._params[:n, j]
    try:
        import pandas as pd
        return pd.DataFrame(cols)
    except Exception:  # pragma: no cover
        return cols


def best(c: BGPChain):
    """best(c) -> (val, idx) (AlgoBGP.jl:167), idx 1-based"""
    v = c._value[: c.iter]
    i = int(np.argmin(v))
    return float(v[i]), i + 1


def _device_stats(algo: "MAlgoBGP", probs):
    """accepted-only mean / quantiles of every local chain, reduced on the device (smm_bgp_accepted_stats): the trace
    stays in HBM -- a 1024-chain x 1000-iteration run would otherwise ship 240 MB to the host to print a summary"""
    if algo.i == 0 or algo._h is None:
        raise RuntimeError("no device trace: run the algorithm first (and before close())")
    return algo._handle().accepted_stats(probs, 1, algo.i)


def mean(c):
    """mean(c::BGPChain) (AlgoBGP.jl:174-176) over the accepted draws; mean(algo): every local chain, on the device"""
    if isinstance(c, MAlgoBGP):
        names = list(c.m.params_to_sample.keys())
        _, mu, _ = _device_stats(c, ())
        return [dict(zip(names, map(float, row))) for row in mu]
    return {k: float(np.mean(v)) for k, v in params(c).items()}


def median(c):
    """median(c::BGPChain) (AlgoBGP.jl:178-180); median(algo): every local chain, on the device"""
    if isinstance(c, MAlgoBGP):
        names = list(c.m.params_to_sample.keys())
        _, _, q = _device_stats(c, (0.5,))
        return [dict(zip(names, map(float, row[:, 0]))) for row in q]
    return {k: float(np.median(v)) for k, v in params(c).items()}


def CI(c, level: float = 0.95):
    """CI(c::BGPChain; level) (AlgoBGP.jl:182-188): the (1-level)/2 and 1-(1-level)/2 quantiles of the accepted draws;
    CI(algo): every local chain, on the device"""
    q = [(1 - level) / 2, 1 - (1 - level) / 2]
    if isinstance(c, MAlgoBGP):
        names = list(c.m.params_to_sample.keys())
        _, _, qq = _device_stats(c, q)
        return [{k: row[j].copy() for j, k in enumerate(names)} for row in qq]
    return {k: np.quantile(v, q) for k, v in params(c).items()}


def _mode(x):
    vals, counts = np.unique(x, return_counts=True)
    return int(vals[np.argmax(counts)])


def summary(x):
    """summary(c::BGPChain) (AlgoBGP.jl:197-206) / summary(algo) (:541-550)"""
    if isinstance(x, MAlgoBGP):
        if x.i > 0 and x._h is not None:     # reduced on the device: no trace read-back, no chain objects
            h, cfg = x._handle(), x._cfg
            nx, mw, bv = h.chain_summary()
            _, acc = h.chain_state()
            rows = [OrderedDict(id=c * cfg.world_size + cfg.rank + 1, acc_rate=float(acc[c]), perc_exchanged=100.0 * int(nx[c]) / cfg.max_iter,
                                exchanged_most_with=int(mw[c]), best_val=float(bv[c])) for c in range(h.L)]
        else:
            rows = [summary(c) for c in x.chains]
        try:
            import pandas as pd
            return pd.DataFrame(rows)
        except Exception:  # pragma: no cover
            return rows
    c = x
    ex = c.exchanged[c.exchanged != 0]
    return OrderedDict(id=c.id, acc_rate=c.accept_rate, perc_exchanged=100.0 * np.sum(c.exchanged != 0) / len(c.exchanged),
                       exchanged_most_with=_mode(ex) if len(ex) else 0, best_val=float(c.best_val[-1]))


# ---------------------------------------------------------------------------------------------
# MAlgoBGP (AlgoBGP.jl:497-539) + run! (AlgoAbstract.jl:27-76)
# ---------------------------------------------------------------------------------------------
DEFAULT_OPTS = {"N": 3, "maxiter": 100, "maxtemp": 2, "sigma": 0.05, "sigma_update_steps": 10, "sigma_adjust_by": 0.01,
                "smpl_iters": 1000, "parallel": False, "min_improve": [0.0] * 3, "acc_tuners": [2.0] * 3}


class MAlgoBGP:
    """`MAlgoBGP(m, opts)`.  Extra opts understood here: "seed" (Zprop/Uacc/Pairs streams), "device",
    "world_size"/"rank"/"nccl_id" (one process per GPU), "n_split"."""

    def __init__(self, m: MProb, opts: Optional[dict] = None):
        self.m = m
        self.opts = dict(DEFAULT_OPTS if opts is None else opts)
        self.i = 0
        if self.opts.get("dist_fun", None) is not None:
            raise NotImplementedError('opts["dist_fun"]: only the default `-` (AlgoBGP.jl:537) runs on the device')
        self._cfg = _base_config(m, int(self.opts["N"]), int(self.opts["maxiter"]), self.opts)
        self._h: Optional[_lib.BGPHandle] = None
        self._trace: Optional[Trace] = None
        self._state = None
        self._chains = None
        self.device_ms = 0.0

    # algo["key"] (AlgoAbstract.jl:13-19)
    def __getitem__(self, key):
        return self.opts[key]

    def __setitem__(self, key, val):
        self.opts[key] = val

    def _handle(self) -> _lib.BGPHandle:
        if self._h is None:
            try:
                self._h = _lib.BGPHandle(self._cfg)
            except _lib.SMMError as e:
                # an automatically chosen persistent mode may not fit the shape (too many chains per SM): the
                # multi-launch kernels take any shape.  An explicitly requested mode is never overridden.
                if "exchange_mode" in self.opts or self._cfg.exchange_mode == 0 or "UNSUPPORTED_SHAPE" not in str(e):
                    raise
                self._cfg.exchange_mode = 0
                self._h = _lib.BGPHandle(self._cfg)
        return self._h

    def close(self):
        if self._h is not None:
            self._h.close()
            self._h = None
        st = getattr(self, "_streamed", None)
        if st is not None:            # chains already materialised hold copies, not views
            self._trace = None
            st.release()
            self._streamed = None

    @property
    def chains(self) -> "list[BGPChain]":
        """`algo.chains` -- this process's chains (all of them when world_size == 1; with several ranks the chains
        with ids rank + 1, rank + 1 + world, ...)."""
        if self._chains is None:
            self._materialise()
        return self._chains

    def _materialise(self):
        cfg, n = self._cfg, self._cfg.max_iter
        L = cfg.n_chains // cfg.world_size
        gid = np.arange(L) * cfg.world_size + cfg.rank      # local chain c is global chain c * world + rank
        if self.i == 0:
            tr = Trace(0, L, cfg.n_params, cfg.n_moments)
            sigma = np.asarray(cfg.sigma0, float)[gid]
            acc = np.zeros(L)
        else:
            if self._h is None:
                raise RuntimeError("this MAlgoBGP was closed: its device state and trace are gone (read algo.chains before close())")
            h = self._handle()
            st = getattr(self, "_streamed", None)
            tr = st if (st is not None and st.n == self.i) else h.read_trace(1, self.i)
            sigma, acc = h.chain_state()
        self._trace = tr
        self._chains = []
        for c in range(L):
            pa = _lib.acc_uniforms(cfg.seed_algo, int(gid[c]), 1, n)
            self._chains.append(BGPChain(int(gid[c]) + 1, self.m, n, tr, c, self.i, sigma[c], acc[c], pa, cfg))

    def __repr__(self):
        return (f"\nBGP Algorithm with {self.opts['N']} BGPChains\n============================\n\nAlgorithm\n---------\n"
                f"Current iteration: {self.i}\nNumber of params to estimate: {len(self.m.params_to_sample)}\n"
                f"Number of moments to match: {len(self.m.moments)}\n")


def computeNextIteration(algo: MAlgoBGP, n: int = 1):
    """computeNextIteration!(algo) (AlgoBGP.jl:589-640), n iterations at once on the device."""
    h = algo._handle()
    algo.device_ms += h.step(n)
    algo.i = h.iteration
    algo._chains = None
    algo._streamed = None


def run(algo: MAlgoBGP):
    """run!(algo) (AlgoAbstract.jl:27-76): iterations 1..maxiter, optional periodic save.  Without periodic saves
    the whole run is one `smm_bgp_run` call: the trace streams into page-locked host memory window by window
    while the device computes the next window, and `algo.chains` is served from that host copy."""
    t0 = _time.time()
    maxiter = int(algo["maxiter"])
    sf, fn = algo.opts.get("save_frequency"), algo.opts.get("filename")
    chunk = int(sf) if (sf and fn) else maxiter
    if not (sf and fn) and algo.i == 0 and maxiter > 0:
        h = algo._handle()
        tr = _lib.PinnedTrace.acquire(maxiter, h.L, algo._cfg.n_params, algo._cfg.n_moments)
        algo.device_ms += h.run(maxiter, into=tr)
        algo.i = h.iteration
        algo._chains = None
        algo._streamed = tr
    while algo.i < maxiter:
        computeNextIteration(algo, min(chunk, maxiter - algo.i))
        if sf and fn and algo.i % int(sf) == 0:
            save(algo, fn)
    algo.opts["time"] = round((_time.time() - t0) / 60, 1)
    if fn:
        save(algo, fn)
    return algo


def _rank_filename(filename: str, cfg: BGPConfig) -> str:
    """One checkpoint file per rank: a rank's blob holds only the chains it owns, so with several ranks
    the files are `<filename>.rank<r>of<world>`; with one rank the name is used as given (AlgoAbstract.jl:83)."""
    return filename if cfg.world_size == 1 else f"{filename}.rank{cfg.rank}of{cfg.world_size}"


def save(algo: MAlgoBGP, filename: str):
    """save(algo, filename) (AlgoAbstract.jl:83-88): problem + opts + device state checkpoint.  With world_size > 1
    every rank writes its own shard next to the others (`_rank_filename`); the blob itself records rank, world size,
    chain count and seeds, and `smm_bgp_import_state` refuses a blob that belongs to another rank or ensemble."""
    if algo.i > 0 and algo._h is None:
        raise RuntimeError("save: this MAlgoBGP was closed after running -- its device state is gone (save before close())")
    blob = algo._handle().export_state() if algo.i > 0 else b""
    if algo.i > 0 and algo._handle().iteration != algo.i:
        raise RuntimeError(f"save: the device is at iteration {algo._handle().iteration}, the algorithm object at {algo.i}")
    opts = {k: v for k, v in algo.opts.items() if k != "nccl_id"}     # a communicator id does not survive the process
    with open(_rank_filename(filename, algo._cfg), "wb") as f:
        pickle.dump({"m": algo.m, "opts": opts, "i": algo.i, "state": blob, "rank": algo._cfg.rank,
                     "world_size": algo._cfg.world_size}, f)


def readMalgo(filename: str, opts_override: Optional[dict] = None) -> MAlgoBGP:
    """readMalgo(filename) (AlgoAbstract.jl:95-102).  With several ranks each rank reads its own shard; pass the
    placement of the new job (`rank`, `world_size`, `device`, `nccl_id`) in `opts_override`."""
    o = dict(opts_override or {})
    probe = BGPConfig(lb=[0.0], ub=[1.0], init=[0.5], data_mom=[0.0], data_w=[1.0], n_chains=1, max_iter=1, sigma0=[1.0],
                      acc_tuner=[1.0], min_improve=[0.0], world_size=int(o.get("world_size", 1)), rank=int(o.get("rank", 0)))
    with open(_rank_filename(filename, probe), "rb") as f:
        d = pickle.load(f)
    if d.get("world_size", 1) != probe.world_size or d.get("rank", 0) != probe.rank:
        raise ValueError(f"{filename}: checkpoint of rank {d.get('rank')} of {d.get('world_size')}, "
                         f"asked for rank {probe.rank} of {probe.world_size}")
    opts = dict(d["opts"])
    opts.update(o)
    algo = MAlgoBGP(d["m"], opts)
    if d["i"] > 0:
        algo._handle().import_state(d["state"])
        if algo._handle().iteration != d["i"]:
            raise RuntimeError("readMalgo: the checkpoint's iteration count does not match its state blob")
        algo.i = d["i"]
    return algo


def restart(algo: MAlgoBGP, extraIter: int):
    """restart!(algo, extraIter) (AlgoBGP.jl:804-884): grow every chain by extraIter slots
    (extendBGPChain! :759-796) and continue.  Upstream re-runs iteration `algo.i` (its loop starts at
    initialIter, :835) with fresh random numbers; here all randomness is counter-indexed, so iteration
    `algo.i` would be recomputed bit-identically -- we therefore continue at algo.i + 1, and a restarted
    run equals a straight run of maxiter + extraIter (the property upstream's test meant to check)."""
    old_i = algo.i
    blob = algo._handle().export_state() if old_i > 0 else None
    algo.close()
    algo.opts["maxiter"] = old_i + int(extraIter)
    algo._cfg = _base_config(algo.m, int(algo.opts["N"]), int(algo.opts["maxiter"]), algo.opts)
    if blob is not None:
        algo._handle().import_state(blob)
    algo.i = old_i
    algo._chains = None
    return run(algo)


# Julia-name lookup for the drop-in table in INTEGRATION.md
JULIA_NAMES = {
    "addParam!": addParam, "addSampledParam!": addSampledParam, "addMoment!": addMoment, "addEvalFunc!": addEvalFunc,
    "addEvalFuncOpts!": addEvalFuncOpts, "setMoments!": setMoments, "setValue!": setValue, "run!": run,
    "restart!": restart, "computeNextIteration!": computeNextIteration, "evaluateObjective": evaluateObjective,
    "readMalgo": readMalgo, "save": save, "summary": summary, "history": history, "best": best, "mean": mean,
    "median": median, "CI": CI, "params": params, "param": param, "paramd": paramd, "dataMoment": dataMoment,
    "dataMomentd": dataMomentd, "dataMomentW": dataMomentW, "allAccepted": allAccepted, "start": start, "finish": finish,
}


# ---------------------------------------------------------------------------------------------
# slices and inference on the batched objective entry (slices.jl, econometrics.jl) -- see slices.py
# ---------------------------------------------------------------------------------------------
from .slices import (Slice, doSlices, optSlices, FD_gradient, getSigma, get_stdErrors, range_length)  # noqa: E402,F401
from .slices import save as saveSlice, load as loadSlice  # noqa: E402,F401  (save(s::Slice, f) / load(f), slices.jl:283-289)

JULIA_NAMES.update({"doSlices": doSlices, "optSlices": optSlices, "FD_gradient": FD_gradient, "getSigma": getSigma,
                    "get_stdErrors": get_stdErrors, "evaluateObjectiveBatch": evaluateObjectiveBatch})
