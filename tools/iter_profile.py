import sys
sys.path.insert(0, ".")
import numpy as np
from smm_jl_b200 import configs, _lib
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = configs.mvnormal(256, 1600, exchange_mode=mode)
with _lib.BGPHandle(cfg) as h:
    prev_att = 0
    for blk in range(16):
        ms = h.step(100)
        c = h.counters()
        att = c["proposal_attempts"] - prev_att
        prev_att = c["proposal_attempts"]
        sigma, acc = h.chain_state()
        print(f"iters {blk*100+1:5d}-{blk*100+100:5d}: {ms*10:7.1f} us/iter  attempts/proposal {att/(256*100):8.1f}  sigma max {sigma.max():.3f}")
