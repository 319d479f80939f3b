"""Quick device-side timing sweep (not the contract bench)."""
import sys
sys.path.insert(0, ".")
from smm_jl_b200 import configs, _lib

def run(n_chains, n_split, mode, iters=300, warm=20):
    cfg = configs.mvnormal(n_chains, iters + warm, n_split=n_split, exchange_mode=mode)
    with _lib.BGPHandle(cfg) as h:
        h.step(warm)
        ms = h.step(iters)
    return n_chains * iters / (ms * 1e-3), ms / iters * 1e3

if __name__ == "__main__":
    chains = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    ms, rate = _lib.rng_throughput(2000, 148 * 8)
    print(f"rng-only: {rate/1e9:.2f} G normals/s ({ms:.3f} ms)")
    for mode, splits in ((0, [0, 4, 8]), (1, [0, 4, 5, 6, 7, 8])):
        for ns in splits:
            try:
                r, us = run(chains, ns, mode)
                print(f"mode={mode} chains={chains} n_split={ns:2d}: {r/1e6:.3f} M evals/s, {us:.1f} us/iter, {r*80000/1e9:.1f} G normals/s")
            except Exception as e:
                print(mode, ns, "failed", e)
