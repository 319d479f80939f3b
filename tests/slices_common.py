"""Shared by the CPU and GPU tests of smm_jl_b200/slices.py: the serialNormal problem and an evaluator backed by the
oracle (test infrastructure) with the signature of api.evaluateObjectiveBatch."""
from collections import OrderedDict

import numpy as np

from smm_jl_b200 import api


def serial_normal_problem():
    # Examples.jl:373-446 (snorm_impl, npar = 2)
    m = api.MProb()
    api.addSampledParam(m, {"p1": [0.2, -3, 3], "p2": [-0.2, -20, 20]})
    api.addMoment(m, {"name": ["mu1", "mu2"], "value": [-1.0, 10.0], "weight": [1.0, 1.0]})
    api.addEvalFunc(m, api.objfunc_norm)
    return m


def oracle_evaluator(oracle):
    def evaluate(m, plist, noseed=False, rep0=0):
        names = list(m.params_to_sample.keys())
        cfg = api._base_config(m, 1, 1, {})
        P = np.array([[float(p[k]) for k in names] for p in plist], dtype=float)
        value, mom, status = oracle.eval_batch(cfg, P, noseed=int(noseed), rep0=rep0, n_threads=4)
        out = []
        for b, p in enumerate(plist):
            ev = api.Eval(m, OrderedDict((k, float(p[k])) for k in names))
            ev.value, ev.status = float(value[b]), int(status[b])
            if status[b] >= 0:
                for k, v in zip(m.moments.keys(), mom[b]):
                    ev.simMoments[k] = float(v)
            out.append(ev)
        return out
    return evaluate
