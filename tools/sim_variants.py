"""Ablations of the simulate inner loop (sim_throughput_kernel variants): where its cycles go.
variant 0 = the product's loop, 1 = deferred queue off, 2 = also no table lookup, 3 = Philox alone."""
import sys
sys.path.insert(0, ".")
from smm_jl_b200 import configs, _lib
cfg = configs.mvnormal(256, 10)
names = {0: "full", 1: "no deferred queue", 2: "no queue, no table", 3: "Philox only"}
with _lib.BGPHandle(cfg) as h:
    for threads in (1024, 768):
        for variant in (0, 1, 2, 3):
            ms, rate = h.sim_throughput(2000 * 1024 // threads, 148, threads, variant << 2)
            steps = 2000 * 1024 // threads * (threads // 32) / 4          # warp steps per SMSP
            print(f"threads={threads:5d} {names[variant]:22s}: {rate/1e9:7.1f} G normals/s  ({ms:.3f} ms, "
                  f"{ms * 1e-3 * 1.965e9 / steps:6.1f} cycles per warp step per SMSP at 1965 MHz)")
