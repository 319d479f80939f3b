"""A/B of the persistent kernel's two hand-over schemes on C2 (device-side timing, 950 iterations after 50)."""
import sys
sys.path.insert(0, ".")
from smm_jl_b200 import configs, _lib
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for mode in (1, 2, 1, 2):
    cfg = configs.mvnormal(chains, 1000, exchange_mode=mode)
    with _lib.BGPHandle(cfg) as h:
        h.step(50)
        ms = h.step(950)
        print(f"exchange_mode={mode} chains={chains}: {chains * 950 / (ms * 1e-3) / 1e6:.3f} M evals/s, {ms / 950 * 1e3:.2f} us/iter", flush=True)
