import sys
sys.path.insert(0, ".")
from smm_jl_b200 import configs, _lib
cfg = configs.mvnormal(256, 10)
with _lib.BGPHandle(cfg) as h:
    for (blocks, threads) in [(148 * 8, 128), (148 * 4, 256), (148 * 2, 512), (148, 1024), (148 * 6, 128), (148 * 4, 128)]:
        for dyn in (0, 1):
            ms, rate = h.sim_throughput(2000 * 1024 // (blocks * threads // 148), blocks, threads, dyn)
            print(f"blocks={blocks:5d} threads={threads:5d} dynamic={dyn}: {rate/1e9:7.1f} G normals/s  ({ms:.3f} ms)")
ms, rate = _lib.rng_throughput(2000, 148 * 8)
print(f"pure rng kernel: {rate/1e9:.1f} G normals/s")
