import sys
sys.path.insert(0, ".")
from smm_jl_b200 import configs, _lib
cfg = configs.mvnormal(256, 10)
with _lib.BGPHandle(cfg) as h:
    for (blocks, threads) in [(148 * 8, 128), (148 * 2, 512), (148, 1024), (148, 768), (148, 512), (148, 640), (148, 896)]:
        for dyn in (0, 1, 2, 3):
            ms, rate = h.sim_throughput(2000 * 1024 // (blocks * threads // 148), blocks, threads, dyn)
            print(f"blocks={blocks:5d} threads={threads:5d} queue={dyn & 1} tight_launch_bound={dyn >> 1}: {rate/1e9:7.1f} G normals/s  ({ms:.3f} ms)")
ms, rate = _lib.rng_throughput(2000, 148 * 8)
print(f"Box-Muller rng kernel (first half of round 1): {rate/1e9:.1f} G normals/s")
