"""Tiny workload for ncu: C2 (256 chains), a few launches of the dominant kernel + the RNG-only ceiling kernel."""
import sys
sys.path.insert(0, ".")
from smm_jl_b200 import configs, _lib
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = configs.mvnormal(256, 200, exchange_mode=mode)
with _lib.BGPHandle(cfg) as h:
    h.step(20)      # warm-up launch(es)
    h.step(20)      # profiled: one persistent launch of 20 iterations (mode 1) / 20 evaluation launches (mode 0)
    h.step(20)
    h.sim_throughput(500, 148, 1024, True)
