"""ctypes mirror of include/smm_b200.h (structs, constants) -- interface only, no compute.

The product binding (smm_jl_b200/_lib.py -> libsmm_b200.so) fills `smm_bgp_config` from here; the test
suite's CPU checker fills the very same struct, which is what makes "same inputs" in the parity tests
literal.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

SMM_ABI_VERSION = 1

SMM_OK = 0
SMM_E_ARG = -1
SMM_E_CUDA = -2
SMM_E_NCCL = -3
SMM_E_UNSUPPORTED_SHAPE = -4
SMM_E_NEGATIVE_OBJECTIVE = -5
SMM_E_SAMPLER_EXHAUSTED = -6
SMM_E_STATE = -7

ERROR_NAMES = {
    SMM_E_ARG: "SMM_E_ARG",
    SMM_E_CUDA: "SMM_E_CUDA",
    SMM_E_NCCL: "SMM_E_NCCL",
    SMM_E_UNSUPPORTED_SHAPE: "SMM_E_UNSUPPORTED_SHAPE",
    SMM_E_NEGATIVE_OBJECTIVE: "SMM_E_NEGATIVE_OBJECTIVE",
    SMM_E_SAMPLER_EXHAUSTED: "SMM_E_SAMPLER_EXHAUSTED",
    SMM_E_STATE: "SMM_E_STATE",
}

SMM_OBJ_NORM = 0
SMM_OBJ_NORM_SLOW = 1
SMM_OBJ_NORM_MV = 2
SMM_OBJ_PANEL = 3
SMM_OBJ_FAILS = 4

SMM_NCCL_ID_BYTES = 128

_dp = C.POINTER(C.c_double)


class smm_bgp_config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("n_params", C.c_int32),
        ("n_moments", C.c_int32),
        ("lb", _dp),
        ("ub", _dp),
        ("init", _dp),
        ("data_mom", _dp),
        ("data_w", _dp),
        ("objective_id", C.c_int32),
        ("n_sim", C.c_int32),
        ("seed_sim", C.c_uint64),
        ("noseed", C.c_int32),
        ("slow_seconds", C.c_double),
        ("panel_T", C.c_int32),
        ("panel_N", C.c_int32),
        ("panel_K", C.c_int32),
        ("n_chains", C.c_int32),
        ("max_iter", C.c_int32),
        ("sigma0", _dp),
        ("acc_tuner", _dp),
        ("min_improve", _dp),
        ("sigma_update_steps", C.c_int32),
        ("sigma_adjust_by", C.c_double),
        ("smpl_iters", C.c_int32),
        ("batch_size", C.c_int32),
        ("seed_algo", C.c_uint64),
        ("device", C.c_int32),
        ("world_size", C.c_int32),
        ("rank", C.c_int32),
        ("nccl_id", C.c_uint8 * SMM_NCCL_ID_BYTES),
        ("exchange_mode", C.c_int32),
        ("n_split", C.c_int32),
    ]


class smm_trace_view(C.Structure):
    _fields_ = [
        ("value", _dp),
        ("prob", _dp),
        ("curr_val", _dp),
        ("best_val", _dp),
        ("params", _dp),
        ("sim_moments", _dp),
        ("accepted", C.POINTER(C.c_uint8)),
        ("status", C.POINTER(C.c_int32)),
        ("exchanged", C.POINTER(C.c_int32)),
        ("best_id", C.POINTER(C.c_int32)),
    ]


class smm_counters(C.Structure):
    _fields_ = [
        ("iterations", C.c_int64),
        ("evaluations", C.c_int64),
        ("kernel_launches", C.c_int64),
        ("collectives", C.c_int64),
        ("accepted", C.c_int64),
        ("swaps", C.c_int64),
        ("proposal_attempts", C.c_int64),
    ]


def _f64(a, n=None, name="array"):
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
    if n is not None and arr.size != n:
        raise ValueError(f"{name}: expected {n} values, got {arr.size}")
    return arr


@dataclass
class BGPConfig:
    """Plain-Python form of `smm_bgp_config`; `.c_struct()` pins numpy buffers and returns the C view."""

    lb: Sequence[float]
    ub: Sequence[float]
    init: Sequence[float]
    data_mom: Sequence[float]
    data_w: Sequence[float]
    n_chains: int
    max_iter: int
    sigma0: Sequence[float]
    acc_tuner: Sequence[float]
    min_improve: Sequence[float]
    objective_id: int = SMM_OBJ_NORM
    n_sim: int = 10000
    seed_sim: int = 1234
    noseed: int = 0
    slow_seconds: float = 0.1
    panel_T: int = 0
    panel_N: int = 0
    panel_K: int = 0
    sigma_update_steps: int = 10
    sigma_adjust_by: float = 0.01
    smpl_iters: int = 1000
    batch_size: Optional[int] = None
    seed_algo: int = 20261017
    device: int = 0
    world_size: int = 1
    rank: int = 0
    nccl_id: bytes = b""
    exchange_mode: int = 0
    n_split: int = 0
    _keep: list = field(default_factory=list, repr=False)

    @property
    def n_params(self) -> int:
        return len(np.asarray(self.lb).reshape(-1))

    @property
    def n_moments(self) -> int:
        return len(np.asarray(self.data_mom).reshape(-1))

    def c_struct(self) -> smm_bgp_config:
        P, M, N = self.n_params, self.n_moments, int(self.n_chains)
        bufs = dict(
            lb=_f64(self.lb, P, "lb"),
            ub=_f64(self.ub, P, "ub"),
            init=_f64(self.init, P, "init"),
            data_mom=_f64(self.data_mom, M, "data_mom"),
            data_w=_f64(self.data_w, M, "data_w"),
            sigma0=_f64(self.sigma0, N, "sigma0"),
            acc_tuner=_f64(self.acc_tuner, N, "acc_tuner"),
            min_improve=_f64(self.min_improve, N, "min_improve"),
        )
        self._keep = [bufs]
        s = smm_bgp_config()
        s.abi_version = SMM_ABI_VERSION
        s.n_params, s.n_moments = P, M
        for k, v in bufs.items():
            # a ctypes array view of the numpy buffer converts to the pointer field directly (several times cheaper than
            # ndarray.ctypes.data_as, and this constructor is inside the timed region of a short run)
            setattr(s, k, (C.c_double * max(v.size, 1)).from_buffer(v) if v.size else None)
        s.objective_id = int(self.objective_id)
        s.n_sim = int(self.n_sim)
        s.seed_sim = int(self.seed_sim)
        s.noseed = int(self.noseed)
        s.slow_seconds = float(self.slow_seconds)
        s.panel_T, s.panel_N, s.panel_K = int(self.panel_T), int(self.panel_N), int(self.panel_K)
        s.n_chains, s.max_iter = N, int(self.max_iter)
        s.sigma_update_steps = int(self.sigma_update_steps)
        s.sigma_adjust_by = float(self.sigma_adjust_by)
        s.smpl_iters = int(self.smpl_iters)
        s.batch_size = int(self.batch_size if self.batch_size is not None else P)
        s.seed_algo = int(self.seed_algo)
        s.device, s.world_size, s.rank = int(self.device), int(self.world_size), int(self.rank)
        idb = bytes(self.nccl_id)[:SMM_NCCL_ID_BYTES].ljust(SMM_NCCL_ID_BYTES, b"\0")
        C.memmove(s.nccl_id, idb, SMM_NCCL_ID_BYTES)
        s.exchange_mode = int(self.exchange_mode)
        s.n_split = int(self.n_split)
        return s


class Trace:
    """Host SoA buffers for `n` iterations of `L` chains, in the layout of `smm_trace_view`."""

    def __init__(self, n: int, L: int, P: int, M: int):
        self.n, self.L, self.P, self.M = n, L, P, M
        self.value = np.full((n, L), np.nan)
        self.prob = np.full((n, L), np.nan)
        self.curr_val = np.full((n, L), np.nan)
        self.best_val = np.full((n, L), np.nan)
        self.params = np.full((n, L, P), np.nan)
        self.sim_moments = np.full((n, L, M), np.nan)
        self.accepted = np.zeros((n, L), dtype=np.uint8)
        self.status = np.zeros((n, L), dtype=np.int32)
        self.exchanged = np.zeros((n, L), dtype=np.int32)
        self.best_id = np.zeros((n, L), dtype=np.int32)

    FLOAT_FIELDS = ("value", "prob", "curr_val", "best_val", "params", "sim_moments")
    INT_FIELDS = ("accepted", "status", "exchanged", "best_id")

    def view(self) -> smm_trace_view:
        v = smm_trace_view()
        for f in self.FLOAT_FIELDS:
            setattr(v, f, getattr(self, f).ctypes.data_as(_dp))
        v.accepted = self.accepted.ctypes.data_as(C.POINTER(C.c_uint8))
        for f in ("status", "exchanged", "best_id"):
            setattr(v, f, getattr(self, f).ctypes.data_as(C.POINTER(C.c_int32)))
        return v

    @staticmethod
    def interleave_ranks(parts: "list[Trace]") -> "Trace":
        """Join per-rank traces (each [n][L]) into the full [n][N] trace: rank r's column c is global chain
        c * world + r (chains are dealt to the ranks round robin)."""
        n, P, M, W = parts[0].n, parts[0].P, parts[0].M, len(parts)
        out = Trace(n, sum(p.L for p in parts), P, M)
        for f in Trace.FLOAT_FIELDS + Trace.INT_FIELDS:
            a = np.stack([getattr(p, f) for p in parts], axis=2)           # [n][L][W](...)
            setattr(out, f, np.ascontiguousarray(a.reshape((n, parts[0].L * W) + a.shape[3:])))
        return out
