"""Tiny workload for ncu: C2 (256 chains) for a few iterations + one RNG micro-kernel launch."""
import sys
sys.path.insert(0, ".")
from smm_jl_b200 import configs, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = configs.mvnormal(256, n)
with _lib.BGPHandle(cfg) as h:
    h.step(n)
_lib.rng_throughput(500, 148 * 8)
