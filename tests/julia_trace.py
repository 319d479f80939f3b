"""Reader / writer of the trace directory julia/parity_harness.jl produces (`<case>/julia_trace/`): raw little-endian
arrays that Julia wrote column-major with dimensions (N, I) / (P, N, I) / (M, N, I), i.e. C order [I][N] (+[P] / [M])."""
import os

import numpy as np

from smm_jl_b200._abi import Trace

FILES = {"value": ("value.f64", "<f8"), "prob": ("prob.f64", "<f8"), "curr_val": ("curr_val.f64", "<f8"),
         "best_val": ("best_val.f64", "<f8"), "params": ("params.f64", "<f8"), "sim_moments": ("sim_moments.f64", "<f8"),
         "accepted": ("accepted.u8", "u1"), "status": ("status.i32", "<i4"), "exchanged": ("exchanged.i32", "<i4"),
         "best_id": ("best_id.i32", "<i4")}


def read_meta(path: str) -> dict:
    out = {}
    with open(path) as f:
        for ln in f:
            if "=" in ln:
                k, v = ln.rstrip("\n").split("=", 1)
                out[k] = v
    return out


def load(trace_dir: str, N: int, I: int, P: int, M: int):
    """-> (Trace [I][N], sigma [N], accept_rate [N])"""
    tr = Trace(I, N, P, M)
    shapes = {"params": (I, N, P), "sim_moments": (I, N, M)}
    for f, (name, dt) in FILES.items():
        a = np.fromfile(os.path.join(trace_dir, name), dtype=dt).reshape(shapes.get(f, (I, N)))
        setattr(tr, f, a.astype(getattr(tr, f).dtype))
    sigma = np.fromfile(os.path.join(trace_dir, "sigma.f64"), dtype="<f8")
    acc = np.fromfile(os.path.join(trace_dir, "accept_rate.f64"), dtype="<f8")
    return tr, sigma, acc


def store(trace_dir: str, tr: Trace, sigma, accept_rate, **meta) -> None:
    """what the harness writes, from a Python trace (used by the self-test of the comparison path)"""
    os.makedirs(trace_dir, exist_ok=True)
    for f, (name, dt) in FILES.items():
        np.ascontiguousarray(getattr(tr, f)).astype(dt).tofile(os.path.join(trace_dir, name))
    np.asarray(sigma, dtype="<f8").tofile(os.path.join(trace_dir, "sigma.f64"))
    np.asarray(accept_rate, dtype="<f8").tofile(os.path.join(trace_dir, "accept_rate.f64"))
    with open(os.path.join(trace_dir, "meta.txt"), "w") as f:
        for k, v in meta.items():
            f.write(f"{k}={v}\n")
