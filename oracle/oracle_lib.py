"""ctypes binding of oracle/libsmm_oracle.so (the C++ restatement of the reference's algorithm).

TEST INFRASTRUCTURE: the product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from smm_jl_b200._abi import BGPConfig, Trace, smm_bgp_config, smm_trace_view, ERROR_NAMES  # interface only

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsmm_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (g++)."""
    if force and os.path.exists(_SO):
        os.remove(_SO)
    subprocess.run(["make", "-C", _HERE, "--no-print-directory"], check=True, capture_output=True)
    return _SO


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {msg}")
        self.code = code


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    L = C.CDLL(_SO)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    L.smm_oracle_last_error.restype = C.c_char_p
    L.smm_oracle_run.argtypes = [C.POINTER(smm_bgp_config), C.c_int, C.c_int, C.POINTER(smm_trace_view), dp, dp,
                                 C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.smm_oracle_eval_batch.argtypes = [C.POINTER(smm_bgp_config), dp, C.c_int, C.c_int, C.c_uint32, C.c_int, dp, dp, ip]
    L.smm_oracle_pairs.argtypes = [C.c_uint64, C.c_int, C.c_int, ip]
    L.smm_oracle_philox.argtypes = [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]
    L.smm_oracle_philox.restype = None
    L.smm_oracle_normals.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, dp]
    L.smm_oracle_normals.restype = None
    L.smm_oracle_normal_from_words.argtypes = [C.c_uint32] * 4 + [dp]
    L.smm_oracle_normal_from_words.restype = None
    L.smm_oracle_zig_normals.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, dp]
    L.smm_oracle_zig_normals.restype = None
    L.smm_oracle_zig_from_words.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_int)]
    L.smm_oracle_zig_from_words.restype = C.c_double
    L.smm_oracle_zig_select.argtypes = [C.c_uint32, C.c_int]
    L.smm_oracle_zig_select.restype = C.c_uint32
    L.smm_oracle_exp_neg.argtypes = [C.c_double]
    L.smm_oracle_exp_neg.restype = C.c_double
    L.smm_oracle_proxy_rate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    L.smm_oracle_proxy_rate.restype = C.c_double
    L.smm_oracle_neglog01.argtypes = [C.c_double]
    L.smm_oracle_neglog01.restype = C.c_double
    L.smm_oracle_acc_uniform.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
    L.smm_oracle_acc_uniform.restype = C.c_double
    L.smm_oracle_pair_unrank.argtypes = [C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.smm_oracle_pair_unrank.restype = None
    L.smm_oracle_accept_reject.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, dp,
                                           C.POINTER(C.c_int), C.POINTER(C.c_int)]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise OracleError(rc, lib().smm_oracle_last_error().decode())


class OracleResult:
    def __init__(self, trace, sigma, accept_rate, swaps, attempts):
        self.trace, self.sigma, self.accept_rate, self.swaps, self.attempts = trace, sigma, accept_rate, swaps, attempts


def run(cfg: BGPConfig, n_iters: int, n_threads: int = 1) -> OracleResult:
    """run!(MAlgoBGP(...)) for n_iters iterations; returns the full [n_iters][N] trace."""
    cs = cfg.c_struct()
    tr = Trace(n_iters, cfg.n_chains, cfg.n_params, cfg.n_moments)
    v = tr.view()
    sigma = np.zeros(cfg.n_chains)
    acc = np.zeros(cfg.n_chains)
    swaps, att = C.c_int64(0), C.c_int64(0)
    dp = C.POINTER(C.c_double)
    _check(lib().smm_oracle_run(C.byref(cs), n_iters, n_threads, C.byref(v), sigma.ctypes.data_as(dp),
                                acc.ctypes.data_as(dp), C.byref(swaps), C.byref(att)))
    return OracleResult(tr, sigma, acc, swaps.value, att.value)


def eval_batch(cfg: BGPConfig, params, noseed: int = 0, rep0: int = 0, n_threads: int = 1):
    params = np.ascontiguousarray(np.asarray(params, dtype=np.float64).reshape(-1, cfg.n_params))
    B = params.shape[0]
    cs = cfg.c_struct()
    value = np.zeros(B)
    mom = np.zeros((B, cfg.n_moments))
    status = np.zeros(B, dtype=np.int32)
    dp = C.POINTER(C.c_double)
    _check(lib().smm_oracle_eval_batch(C.byref(cs), params.ctypes.data_as(dp), B, noseed, rep0, n_threads,
                                       value.ctypes.data_as(dp), mom.ctypes.data_as(dp),
                                       status.ctypes.data_as(C.POINTER(C.c_int32))))
    return value, mom, status


def pairs(seed_algo: int, it: int, N: int) -> np.ndarray:
    buf = np.zeros((max(N, 1), 2), dtype=np.int32)
    n = lib().smm_oracle_pairs(seed_algo, it, N, buf.ctypes.data_as(C.POINTER(C.c_int32)))
    return buf[:n].copy()


def philox(c, k) -> np.ndarray:
    out = (C.c_uint32 * 4)()
    lib().smm_oracle_philox(c[0], c[1], c[2], c[3], k[0], k[1], out)
    return np.array(list(out), dtype=np.uint32)


def normals(seed: int, k: int, c2: int, c3: int, n_pairs: int) -> np.ndarray:
    out = np.zeros(2 * n_pairs)
    lib().smm_oracle_normals(seed, k, c2, c3, n_pairs, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def normal_from_words(x, y, z, w):
    out = np.zeros(2)
    lib().smm_oracle_normal_from_words(x, y, z, w, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def zig_normals(seed: int, k: int, c2: int, c3: int, n_blocks: int) -> np.ndarray:
    """the 3 * n_blocks ziggurat normals of Philox blocks (j, k, c2, c3), j < n_blocks (the MvNormal simulator stream)"""
    out = np.zeros(3 * n_blocks)
    lib().smm_oracle_zig_normals(seed, k, c2, c3, n_blocks, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def zig_from_words(u: int, sel: int):
    """one ziggurat normal from its 32-bit uniform and 10-bit select field -> (z, slow_path_taken)"""
    slow = C.c_int(0)
    z = lib().smm_oracle_zig_from_words(u, sel, C.byref(slow))
    return z, bool(slow.value)


def zig_select(w: int, t: int) -> int:
    return int(lib().smm_oracle_zig_select(w, t))


def exp_neg(t: float) -> float:
    return lib().smm_oracle_exp_neg(t)


def proxy_rate(n_dims: int, n_sim: int, n_evals_per_thread: int, n_threads: int) -> float:
    """evaluations/s of the "reference-speed proxy": the reference's data flow (draw matrix materialised, then reduced)
    with a xoshiro256++ / ziggurat generator of the kind Julia's randn is; not stream compatible, timing only"""
    return lib().smm_oracle_proxy_rate(n_dims, n_sim, n_evals_per_thread, n_threads)


def neglog01(u: float) -> float:
    return lib().smm_oracle_neglog01(u)


def acc_uniform(seed, chain, it):
    return lib().smm_oracle_acc_uniform(seed, chain, it)


def pair_unrank(q):
    i, j = C.c_uint32(), C.c_uint32()
    lib().smm_oracle_pair_unrank(q, C.byref(i), C.byref(j))
    return i.value, j.value


def accept_reject(it, old_value, new_value, new_status, acc_tuner, u):
    prob, acc, st = C.c_double(), C.c_int(), C.c_int()
    _check(lib().smm_oracle_accept_reject(it, old_value, new_value, new_status, acc_tuner, u, C.byref(prob),
                                          C.byref(acc), C.byref(st)))
    return prob.value, bool(acc.value), st.value
