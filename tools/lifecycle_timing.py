"""Where the wall time of one complete run!(MAlgoBGP) goes outside the kernels: create / run / destroy."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smm_jl_b200 import _lib, configs

for n_iter in (10, 1000, 1000, 1000):
    cfg = configs.mvnormal(256, n_iter, exchange_mode=1)
    t0 = time.perf_counter()
    h = _lib.BGPHandle(cfg)
    t1 = time.perf_counter()
    buf = _lib.PinnedTrace.acquire(n_iter, 256, 8, 16)
    t2 = time.perf_counter()
    h.run(n_iter, into=buf)
    t3 = time.perf_counter()
    h.close()
    t4 = time.perf_counter()
    buf.release()
    print(f"max_iter {n_iter}: create {1e3*(t1-t0):.2f} ms, pinned {1e3*(t2-t1):.2f} ms, run {1e3*(t3-t2):.2f} ms, destroy {1e3*(t4-t3):.2f} ms")
