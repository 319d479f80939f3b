"""Quick device-side timing sweep (not the contract bench): C2 at several n_split values + RNG micro-kernel."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from smm_jl_b200 import configs, _lib

def run(n_chains, n_split, iters=200, warm=20, **kw):
    cfg = configs.mvnormal(n_chains, iters + warm, n_split=n_split, **kw)
    with _lib.BGPHandle(cfg) as h:
        h.step(warm)
        ms = h.step(iters)
    return n_chains * iters / (ms * 1e-3), ms / iters * 1e3

if __name__ == "__main__":
    chains = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    for blocks in (148 * 8,):
        ms, rate = _lib.rng_throughput(2000, blocks)
        print(f"rng-only: {rate/1e9:.2f} G normals/s ({ms:.3f} ms)")
    for ns in [0, 1, 2, 3, 4, 5, 6, 8, 12, 16, 24, 37]:
        try:
            r, us = run(chains, ns)
            print(f"chains={chains} n_split={ns:2d}: {r/1e6:.3f} M evals/s, {us:.1f} us/iter, {r*80000/1e9:.1f} G normals/s")
        except Exception as e:
            print(ns, "failed", e)
