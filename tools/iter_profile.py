"""Evidence for two choices of the C2 workload (smm_jl_b200/configs.py):
  (a) smpl_iters = 100000 instead of the reference's default 1000: with 1000 rejection attempts the reference ALGORITHM
      aborts at this scale (`error("no draw in support ...")`, AlgoBGP.jl:409) -- shown on the CPU oracle;
  (b) a run is 1000 iterations: the sigma adaptation (AlgoBGP.jl:381-390) grows the hot chains' proposal s.d. without
      bound, attempts per proposal explode and even 100000 attempts run out after a few thousand iterations -- shown on
      the device (needs a GPU) block by block, until SMM_E_SAMPLER_EXHAUSTED.
    python tools/iter_profile.py [exchange_mode] > profiles/iter_profile_r2.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smm_jl_b200 import configs, _lib
from oracle import oracle_lib

print("(a) CPU oracle, C2 with the reference's default smpl_iters = 1000:")
for n in (200, 400, 700, 1000, 1500, 2000):
    try:
        r = oracle_lib.run(configs.mvnormal(256, n, smpl_iters=1000), n, n_threads=os.cpu_count() or 1)
        print(f"    {n:4d} iterations: ok, {r.attempts / (256 * (n - 1)):.1f} attempts per proposal")
    except Exception as e:
        print(f"    {n:4d} iterations: {e}")
        break
if _lib.device_count() < 1:
    print("(b) skipped: no CUDA device")
    sys.exit(0)
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n_total = 4000
print(f"(b) device, C2 with smpl_iters = 100000, exchange_mode {mode}, blocks of 100 iterations:")
cfg = configs.mvnormal(256, n_total, exchange_mode=mode)
with _lib.BGPHandle(cfg) as h:
    prev_att = 0
    for blk in range(n_total // 100):
        try:
            ms = h.step(100)
        except _lib.SMMError as e:
            print(f"    iterations {blk*100+1:5d}-{blk*100+100:5d}: {e}")
            break
        c = h.counters()
        att = c["proposal_attempts"] - prev_att
        prev_att = c["proposal_attempts"]
        sigma, acc = h.chain_state()
        print(f"    iterations {blk*100+1:5d}-{blk*100+100:5d}: {ms*10:7.1f} us/iter  attempts/proposal {att/(256*100):9.1f}  sigma max {sigma.max():.3f}")
