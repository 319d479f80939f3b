"""Per-CTA phase stamps of rank 0 in an N-GPU run of the bench workload (256 chains per GPU, exchange_mode 2): where an
iteration's time goes when the records travel over NVLink.  Launch under torch.distributed.run with SMM_PHASE_TS=1."""
import os, sys
os.environ["SMM_PHASE_TS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from smm_jl_b200 import configs, _lib

world, rank, lr = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0 and world > 1:
    idt = torch.tensor(list(_lib.nccl_unique_id()), dtype=torch.uint8, device="cuda")
if world > 1:
    dist.broadcast(idt, 0)
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = configs.mvnormal(chains * world, 200, exchange_mode=int(os.environ.get("SMM_PHASE_MODE", "2")))
cfg.device, cfg.world_size, cfg.rank, cfg.nccl_id = lr, world, rank, bytes(idt.cpu().tolist())
pct = lambda a: tuple(np.percentile(a, [0, 50, 90, 100]))
with _lib.BGPHandle(cfg) as h:
    h.step(50)
    ms = h.step(100)
    raw = h.phase_ts().astype(np.int64)
    L = chains
    G = (raw.shape[0] - L) // 6                                      # rows: 4 per CTA | one per chain | 2 per CTA
    per_chain = raw[4 * G:4 * G + L]
    ts = raw[:4 * G].reshape(-1, 2, 2, 4)
    pub = raw[4 * G + L:].reshape(G, 2, 4)                           # [CTA][parity]{before fence, after fence, after adds}
    last_par = h.iteration & 1
    cur, prev = ts[:, last_par], ts[:, 1 - last_par]
    t0 = prev[:, 0, 0].min()
    f = lambda a: (a - t0) / 1e3
    if rank == 0:
        print(f"world {world}, {chains} chains/GPU: us/iter {ms * 10:.2f}")
        print("prev: A start            us: min %.1f med %.1f p90 %.1f max %.1f" % pct(f(prev[:, 0, 0])))
        print("prev: warp0 out of units us: min %.1f med %.1f p90 %.1f max %.1f" % pct(f(prev[:, 0, 1])))
        print("prev: CTA all warps done us: min %.1f med %.1f p90 %.1f max %.1f" % pct(f(prev[:, 0, 2])))
        print("last: wait done (B2 exit)us: min %.1f med %.1f p90 %.1f max %.1f" % pct(f(cur[:, 0, 3])))
        own = cur[:, 1, 1] > 0
        print("last: exchange done      us: min %.1f med %.1f p90 %.1f max %.1f" % pct(f(cur[own, 1, 0])))
        print("last: proposals done     us: min %.1f med %.1f p90 %.1f max %.1f" % pct(f(cur[own, 1, 1])))
        print("last: A start            us: min %.1f med %.1f p90 %.1f max %.1f" % pct(f(cur[:, 0, 0])))
        pp = pub[:, 1 - last_par]                                    # the previous iteration's publish, complete for every CTA
        okp = pp[:, 0] > 0
        if okp.any():
            print("prev: publish: fence start  us: min %.1f med %.1f p90 %.1f max %.1f" % pct(f(pp[okp, 0])))
            print("prev: publish: fence took   us: min %.1f med %.1f p90 %.1f max %.1f" % pct((pp[okp, 1] - pp[okp, 0]) / 1e3))
            print("prev: publish: adds issued  us: min %.1f med %.1f p90 %.1f max %.1f" % pct((pp[okp, 2] - pp[okp, 1]) / 1e3))
            print("prev: publish: all done     us: min %.1f med %.1f p90 %.1f max %.1f" % pct(f(pp[okp, 2])))
        tags = (per_chain[:, 1] - np.median(cur[:, 0, 0])) / 1e3
        print("local chains' tags (relative to the last A start): med %.1f p90 %.1f max %.1f" % tuple(np.percentile(tags, [50, 90, 100])))
if world > 1:
    dist.destroy_process_group()
