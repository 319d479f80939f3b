"""ctypes binding of libsmm_b200.so -- the C ABI declared in include/smm_b200.h.

There is no CPU fallback: if the shared library is missing it is built with nvcc
(smm_jl_b200/build.py); if that fails, or no CUDA device is present, compute calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build
from ._abi import (BGPConfig, ERROR_NAMES, SMM_NCCL_ID_BYTES, Trace, smm_bgp_config, smm_counters, smm_trace_view)

_lib = None

EXPORTS = [
    "smm_abi_version", "smm_last_error", "smm_device_count", "smm_nccl_unique_id", "smm_bgp_create",
    "smm_bgp_destroy", "smm_bgp_step", "smm_bgp_iteration", "smm_bgp_local_chains", "smm_bgp_stream",
    "smm_bgp_read_trace", "smm_bgp_read_chain_state", "smm_bgp_get_counters", "smm_bgp_eval_batch",
    "smm_bgp_state_bytes", "smm_bgp_export_state", "smm_bgp_import_state", "smm_debug_normals", "smm_debug_zig_normals",
    "smm_debug_pairs", "smm_debug_rng_throughput", "smm_stream_acc_uniforms", "smm_bgp_set_profiling",
    "smm_bgp_kernel_times", "smm_debug_phase_ts", "smm_debug_sim_throughput", "smm_debug_barrier_bench",
    "smm_bgp_run", "smm_host_alloc", "smm_host_free", "smm_shutdown", "smm_bgp_accepted_stats", "smm_bgp_chain_summary",
]


class SMMError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {msg}")
        self.code = code


def lib_path() -> str:
    return _build.SO


def lib():
    """Load (building first if needed) libsmm_b200.so."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_build.SO):
        _build.build()
    L = C.CDLL(_build.SO, mode=C.RTLD_GLOBAL)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    vp = C.c_void_p
    L.smm_abi_version.restype = C.c_int
    L.smm_last_error.restype = C.c_char_p
    L.smm_device_count.restype = C.c_int
    L.smm_nccl_unique_id.argtypes = [C.POINTER(C.c_uint8)]
    L.smm_bgp_create.argtypes = [C.POINTER(smm_bgp_config), C.POINTER(vp)]
    L.smm_bgp_destroy.argtypes = [vp]
    L.smm_bgp_destroy.restype = None
    L.smm_bgp_step.argtypes = [vp, C.c_int32, C.POINTER(C.c_float)]
    L.smm_bgp_iteration.argtypes = [vp]
    L.smm_bgp_run.argtypes = [vp, C.c_int32, C.c_int32, C.POINTER(smm_trace_view), C.POINTER(C.c_float)]
    L.smm_host_alloc.argtypes = [C.c_int64, C.POINTER(vp)]
    L.smm_host_free.argtypes = [vp]
    L.smm_host_free.restype = None
    L.smm_shutdown.restype = None
    L.smm_bgp_accepted_stats.argtypes = [vp, C.c_int32, C.c_int32, dp, C.c_int32, C.POINTER(C.c_int64), dp, dp]
    L.smm_bgp_chain_summary.argtypes = [vp, C.POINTER(C.c_int64), ip, dp]
    L.smm_bgp_local_chains.argtypes = [vp]
    L.smm_bgp_stream.argtypes = [vp]
    L.smm_bgp_stream.restype = vp
    L.smm_bgp_read_trace.argtypes = [vp, C.c_int32, C.c_int32, C.POINTER(smm_trace_view)]
    L.smm_bgp_read_chain_state.argtypes = [vp, dp, dp]
    L.smm_bgp_get_counters.argtypes = [vp, C.POINTER(smm_counters)]
    L.smm_bgp_eval_batch.argtypes = [vp, dp, C.c_int32, C.c_int32, C.c_uint32, dp, dp, ip]
    L.smm_bgp_state_bytes.argtypes = [vp]
    L.smm_bgp_state_bytes.restype = C.c_int64
    L.smm_bgp_export_state.argtypes = [vp, vp, C.c_int64]
    L.smm_bgp_import_state.argtypes = [vp, vp, C.c_int64]
    L.smm_debug_normals.argtypes = [C.c_int32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, dp]
    L.smm_debug_zig_normals.argtypes = [C.c_int32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, dp]
    L.smm_debug_pairs.argtypes = [vp, C.c_int32, ip, ip, ip]
    L.smm_debug_rng_throughput.argtypes = [C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_float), dp]
    L.smm_stream_acc_uniforms.argtypes = [C.c_uint64, C.c_uint32, C.c_int32, C.c_int32, dp]
    L.smm_bgp_set_profiling.argtypes = [vp, C.c_int32]
    L.smm_bgp_kernel_times.argtypes = [vp, dp, C.POINTER(C.c_int64)]
    L.smm_debug_phase_ts.argtypes = [vp, C.POINTER(C.c_uint64), C.c_int64]
    L.smm_debug_barrier_bench.argtypes = [vp, C.c_int32, C.c_int32, C.POINTER(C.c_float)]
    L.smm_debug_sim_throughput.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float)]
    _lib = L
    import atexit
    atexit.register(L.smm_shutdown)   # cached communicators / exchange arenas / streams end with the process
    return L


def check(rc: int):
    if rc != 0:
        raise SMMError(rc, lib().smm_last_error().decode(errors="replace"))


def device_count() -> int:
    return lib().smm_device_count()


def nccl_unique_id() -> bytes:
    buf = (C.c_uint8 * SMM_NCCL_ID_BYTES)()
    check(lib().smm_nccl_unique_id(buf))
    return bytes(buf)


class PinnedTrace(Trace):
    """A `Trace` whose columns live in page-locked host memory (smm_host_alloc), so that smm_bgp_run /
    smm_bgp_read_trace copy into it asynchronously."""

    def __init__(self, n: int, L: int, P: int, M: int):
        self.n, self.L, self.P, self.M = n, L, P, M
        shapes = {"value": (n, L), "prob": (n, L), "curr_val": (n, L), "best_val": (n, L), "params": (n, L, P),
                  "sim_moments": (n, L, M), "accepted": (n, L), "status": (n, L), "exchanged": (n, L), "best_id": (n, L)}
        dtypes = {"accepted": np.uint8, "status": np.int32, "exchanged": np.int32, "best_id": np.int32}
        sizes = {f: int(np.prod(sh)) * np.dtype(dtypes.get(f, np.float64)).itemsize for f, sh in shapes.items()}
        total = sum((v + 63) // 64 * 64 for v in sizes.values())
        self._base = C.c_void_p()
        check(lib().smm_host_alloc(total, C.byref(self._base)))
        off = 0
        for f, sh in shapes.items():
            dt = np.dtype(dtypes.get(f, np.float64))
            buf = (C.c_char * sizes[f]).from_address(self._base.value + off)
            setattr(self, f, np.frombuffer(buf, dtype=dt).reshape(sh))
            off += (sizes[f] + 63) // 64 * 64

    def view(self):
        """the C view of the columns, built once: the page-locked buffers never move (a run's setup is inside the timed
        region of a short run; ten ndarray.ctypes.data_as calls cost ~20 us)"""
        v = getattr(self, "_view", None)
        if v is None:
            v = self._view = Trace.view(self)
        return v

    _pool: dict = {}

    @classmethod
    def acquire(cls, n: int, L: int, P: int, M: int) -> "PinnedTrace":
        """a buffer from the free list (page-locking 60 MB costs milliseconds; estimations are run repeatedly)"""
        free = cls._pool.get((n, L, P, M))
        return free.pop() if free else cls(n, L, P, M)

    def release(self):
        """back to the free list; the caller must not keep views into the columns"""
        if getattr(self, "_base", None) is not None and self._base.value:
            lst = PinnedTrace._pool.setdefault((self.n, self.L, self.P, self.M), [])
            if len(lst) < 2:
                lst.append(self)
            else:
                self.free()

    def free(self):
        if getattr(self, "_base", None) is not None and self._base.value:
            self._view = None
            for f in Trace.FLOAT_FIELDS + Trace.INT_FIELDS:   # detach the views first
                setattr(self, f, np.array(getattr(self, f)))
            lib().smm_host_free(self._base)
            self._base = C.c_void_p()

    def __del__(self):
        try:
            if getattr(self, "_base", None) is not None and self._base.value:
                lib().smm_host_free(self._base)
                self._base = C.c_void_p()
        except Exception:  # pragma: no cover
            pass


class BGPHandle:
    """Thin owner of an `smm_bgp*`: create / step / read_trace / destroy."""

    def __init__(self, cfg: BGPConfig):
        self.cfg = cfg
        self._cs = cfg.c_struct()
        self._h = C.c_void_p()
        check(lib().smm_bgp_create(C.byref(self._cs), C.byref(self._h)))
        self.L = lib().smm_bgp_local_chains(self._h)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().smm_bgp_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def iteration(self) -> int:
        return lib().smm_bgp_iteration(self._h)

    @property
    def stream(self) -> int:
        return lib().smm_bgp_stream(self._h) or 0

    def step(self, n_iters: int) -> float:
        """Run n_iters iterations; returns the CUDA-event time of the region in ms."""
        ms = C.c_float(0.0)
        check(lib().smm_bgp_step(self._h, n_iters, C.byref(ms)))
        return ms.value

    def run(self, n_iters: int, into: Trace | None = None, window: int = 0) -> float:
        """Run n_iters iterations, streaming their trace rows into `into` (row 0 = the first new iteration)
        while the device keeps computing; returns the CUDA-event time of the compute region in ms."""
        ms = C.c_float(0.0)
        if into is not None:
            if into.n < n_iters or into.L != self.L:
                raise ValueError("trace buffer too small for this run")
            v = into.view()
            check(lib().smm_bgp_run(self._h, n_iters, window, C.byref(v), C.byref(ms)))
        else:
            check(lib().smm_bgp_run(self._h, n_iters, window, None, C.byref(ms)))
        return ms.value

    def read_trace(self, iter_lo: int = 1, iter_hi: int | None = None, into: Trace | None = None) -> Trace:
        hi = self.iteration if iter_hi is None else iter_hi
        n = hi - iter_lo + 1
        tr = into if into is not None else Trace(n, self.L, self.cfg.n_params, self.cfg.n_moments)
        v = tr.view()
        check(lib().smm_bgp_read_trace(self._h, iter_lo, hi, C.byref(v)))
        return tr

    def chain_state(self):
        sigma, acc = np.zeros(self.L), np.zeros(self.L)
        dp = C.POINTER(C.c_double)
        check(lib().smm_bgp_read_chain_state(self._h, sigma.ctypes.data_as(dp), acc.ctypes.data_as(dp)))
        return sigma, acc

    def counters(self) -> dict:
        c = smm_counters()
        check(lib().smm_bgp_get_counters(self._h, C.byref(c)))
        return {f: getattr(c, f) for f, _ in smm_counters._fields_}

    def eval_batch(self, params, noseed: int = 0, rep0: int = 0):
        P, M = self.cfg.n_params, self.cfg.n_moments
        params = np.ascontiguousarray(np.asarray(params, dtype=np.float64).reshape(-1, P))
        B = params.shape[0]
        value, mom, status = np.zeros(B), np.zeros((B, M)), np.zeros(B, dtype=np.int32)
        dp = C.POINTER(C.c_double)
        check(lib().smm_bgp_eval_batch(self._h, params.ctypes.data_as(dp), B, noseed, rep0, value.ctypes.data_as(dp),
                                       mom.ctypes.data_as(dp), status.ctypes.data_as(C.POINTER(C.c_int32))))
        return value, mom, status

    def accepted_stats(self, probs=(0.5,), iter_lo: int = 1, iter_hi: int | None = None):
        """(count [L], mean [L][P], quantiles [L][P][len(probs)]) of the accepted parameter draws, reduced on the device"""
        hi = self.iteration if iter_hi is None else iter_hi
        P = self.cfg.n_params
        pr = np.ascontiguousarray(np.asarray(probs, dtype=np.float64).reshape(-1))
        cnt = np.zeros(self.L, dtype=np.int64)
        mean, q = np.zeros((self.L, P)), np.zeros((self.L, P, max(pr.size, 1)))
        dp = C.POINTER(C.c_double)
        check(lib().smm_bgp_accepted_stats(self._h, iter_lo, hi, pr.ctypes.data_as(dp), pr.size,
                                           cnt.ctypes.data_as(C.POINTER(C.c_int64)), mean.ctypes.data_as(dp), q.ctypes.data_as(dp)))
        return cnt, mean, q[:, :, :pr.size]

    def chain_summary(self):
        """(iterations with an exchange [L], partner exchanged with most often [L] (1-based, 0 = none), best_val [L])"""
        nx, mw, bv = np.zeros(self.L, dtype=np.int64), np.zeros(self.L, dtype=np.int32), np.zeros(self.L)
        check(lib().smm_bgp_chain_summary(self._h, nx.ctypes.data_as(C.POINTER(C.c_int64)), mw.ctypes.data_as(C.POINTER(C.c_int32)),
                                          bv.ctypes.data_as(C.POINTER(C.c_double))))
        return nx, mw, bv

    def set_profiling(self, on: bool):
        check(lib().smm_bgp_set_profiling(self._h, int(on)))

    def kernel_times(self) -> dict:
        """{kind: (ms_sum, launches)} accumulated while profiling was on"""
        ms = (C.c_double * 4)()
        n = (C.c_int64 * 4)()
        iters = lib().smm_bgp_kernel_times(self._h, ms, n)
        if iters < 0:
            check(iters)
        kinds = ("eval", "exchange", "pairs", "allgather")
        out = {k: (ms[i], n[i]) for i, k in enumerate(kinds)}
        out["eval_iterations"] = iters
        return out

    def phase_ts(self) -> np.ndarray:
        """[L][n_split][4] globaltimer stamps (ns) of the last iteration (needs SMM_PHASE_TS=1 at create)"""
        out = np.zeros(1 << 18, dtype=np.uint64)
        nblk = lib().smm_debug_phase_ts(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64)), out.size)
        if nblk <= 0:
            check(nblk)
        return out[: nblk * 4].reshape(nblk, 4)

    def barrier_bench(self, variant: int, n: int) -> float:
        """microseconds per grid barrier"""
        ms = C.c_float(0.0)
        check(lib().smm_debug_barrier_bench(self._h, variant, n, C.byref(ms)))
        return ms.value * 1e3 / n

    def sim_throughput(self, n_blocks_per_thread: int, blocks: int, threads: int, dynamic: bool):
        """the simulate inner loop alone (three normals per Philox block): returns (ms, normals/s)"""
        ms = C.c_float(0.0)
        check(lib().smm_debug_sim_throughput(self._h, n_blocks_per_thread, blocks, threads, int(dynamic), C.byref(ms)))
        return ms.value, 3.0 * n_blocks_per_thread * blocks * threads / (ms.value * 1e-3)

    def export_state(self) -> bytes:
        n = lib().smm_bgp_state_bytes(self._h)
        buf = C.create_string_buffer(n)
        check(lib().smm_bgp_export_state(self._h, C.cast(buf, C.c_void_p), n))
        return buf.raw

    def import_state(self, blob: bytes):
        buf = C.create_string_buffer(blob, len(blob))
        check(lib().smm_bgp_import_state(self._h, C.cast(buf, C.c_void_p), len(blob)))

    def debug_pairs(self, it: int):
        n_s = self.cfg.n_chains if self.cfg.n_chains >= 3 else self.cfg.n_chains - 1
        ij = np.zeros((n_s, 2), dtype=np.int32)
        off = np.zeros(n_s + 1, dtype=np.int32)
        nlev = C.c_int32(0)
        ip = C.POINTER(C.c_int32)
        check(lib().smm_debug_pairs(self._h, it, ij.ctypes.data_as(ip), off.ctypes.data_as(ip), C.byref(nlev)))
        return ij, off[: nlev.value + 1], nlev.value


def acc_uniforms(seed_algo: int, chain: int, iter_lo: int, iter_hi: int) -> np.ndarray:
    out = np.zeros(iter_hi - iter_lo + 1)
    check(lib().smm_stream_acc_uniforms(seed_algo, chain, iter_lo, iter_hi, out.ctypes.data_as(C.POINTER(C.c_double))))
    return out


def debug_normals(seed: int, k: int, c2: int, c3: int, n_pairs: int, device: int = 0) -> np.ndarray:
    out = np.zeros(2 * n_pairs)
    check(lib().smm_debug_normals(device, seed, k, c2, c3, n_pairs, out.ctypes.data_as(C.POINTER(C.c_double))))
    return out


def debug_zig_normals(seed: int, k: int, c2: int, c3: int, n_blocks: int, device: int = 0) -> np.ndarray:
    """device evaluation of smm_zig_triple on Philox blocks (j, k, c2, c3), j < n_blocks: 3 * n_blocks normals"""
    out = np.zeros(3 * n_blocks)
    check(lib().smm_debug_zig_normals(device, seed, k, c2, c3, n_blocks, out.ctypes.data_as(C.POINTER(C.c_double))))
    return out


def rng_throughput(n_pairs_per_thread: int, blocks: int, device: int = 0):
    """RNG-only micro-kernel (Philox + Box-Muller + 2 accumulations): returns (ms, normals/s)."""
    ms, cs = C.c_float(0.0), C.c_double(0.0)
    check(lib().smm_debug_rng_throughput(device, n_pairs_per_thread, blocks, 128, C.byref(ms), C.byref(cs)))
    normals = 2.0 * n_pairs_per_thread * blocks * 128
    return ms.value, normals / (ms.value * 1e-3)
