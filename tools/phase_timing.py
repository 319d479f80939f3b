"""Per-CTA phase stamps (debug aid).  mode 0: one evaluation-kernel launch; mode 1: the last two iterations
of a persistent launch."""
import os, sys
os.environ["SMM_PHASE_TS"] = "1"
sys.path.insert(0, ".")
import numpy as np
from smm_jl_b200 import configs, _lib
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 256
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = configs.mvnormal(chains, 200, exchange_mode=mode)
pct = lambda a: tuple(np.percentile(a, [0, 50, 90, 100]))
with _lib.BGPHandle(cfg) as h:
    h.step(50)
    if mode == 0:
        h.step(1)
        ts = h.phase_ts().astype(np.int64)
        t0 = ts[:, 0].min()
        start, prop, sim, end = [(ts[:, i] - t0) / 1e3 for i in range(4)]
        print("proposal dur us: min %.1f med %.1f p90 %.1f max %.1f" % pct(prop - start))
        print("simulate dur us: min %.1f med %.1f p90 %.1f max %.1f" % pct(sim - prop))
        print("sim end      us: min %.1f med %.1f p90 %.1f max %.1f" % pct(sim))
    else:
        ms = h.step(100)
        print("us/iter", ms * 10)
        raw = h.phase_ts().astype(np.int64)
        nchain = chains
        G = (raw.shape[0] - nchain) // 6                             # rows: 4 per CTA | one per chain | 2 per CTA
        per_chain = raw[4 * G:4 * G + nchain]                        # {publish start, tag time, finishing CTA, iteration}
        ts = raw[:4 * G].reshape(-1, 2, 2, 4)                        # [block][parity][half][stamp]
        last_par = h.iteration & 1
        cur, prev = ts[:, last_par], ts[:, 1 - last_par]
        t0 = prev[:, 0, 0].min()
        f = lambda a: (a - t0) / 1e3
        own = cur[:, 1, 1] > 0
        print("prev: A start            us: min %.1f med %.1f p90 %.1f max %.1f" % pct(f(prev[:, 0, 0])))
        print("prev: warp0 out of units us: min %.1f med %.1f p90 %.1f max %.1f" % pct(f(prev[:, 0, 1])))
        print("prev: CTA all warps done us: min %.1f med %.1f p90 %.1f max %.1f" % pct(f(prev[:, 0, 2])))
        print("prev: B2 exit            us: min %.1f med %.1f p90 %.1f max %.1f" % pct(f(prev[:, 0, 3])))
        print("last: exchange done      us: min %.1f med %.1f p90 %.1f max %.1f" % pct(f(cur[own, 1, 0])))
        print("last: proposals done     us: min %.1f med %.1f p90 %.1f max %.1f" % pct(f(cur[own, 1, 1])))
        print("last: A start (B1 exit)  us: min %.1f med %.1f p90 %.1f max %.1f" % pct(f(cur[:, 0, 0])))
        print("last: warp0 out of units us: min %.1f med %.1f p90 %.1f max %.1f" % pct(f(cur[:, 0, 1])))
        print("last: CTA all warps done us: min %.1f med %.1f p90 %.1f max %.1f" % pct(f(cur[:, 0, 2])))
        print("last: B2 exit            us: min %.1f med %.1f p90 %.1f max %.1f" % pct(f(cur[:, 0, 3])))
        fin = prev[:, 1, 3] > 0
        for nm, arr in (("prev", prev), ("last", cur)):
            done = arr[:, 0, 2].astype(float)
            order = np.argsort(-done)[:6]
            print(nm, "slowest CTAs (all warps done):", [(int(b), round(float(f(done[b])) - (0 if nm == "prev" else float(f(np.median(cur[:, 0, 0])))), 1)) for b in order])
        print("prev: last finish: publish start us: min %.1f med %.1f p90 %.1f max %.1f" % pct(f(prev[fin, 1, 2])))
        print("prev: last finish: tag published us: min %.1f med %.1f p90 %.1f max %.1f" % pct(f(prev[fin, 1, 3])))
        print("prev: last finish: duration      us: min %.1f med %.1f p90 %.1f max %.1f" % pct((prev[fin, 1, 3] - prev[fin, 1, 2]) / 1e3))

        # the last iteration, chain by chain: who finished it, when, relative to that CTA's own timeline
        tA = np.median(cur[:, 0, 0])
        fb = per_chain[:, 2]
        rel = lambda a: (a - tA) / 1e3
        order = np.argsort(-per_chain[:, 1])[:12]
        print("last iteration: latest chains (chain, finishing CTA, publish start, tag, that CTA: warp0 out / all warps done):")
        for c in order:
            b = int(fb[c])
            print("  chain %3d cta %3d  pub %.1f tag %.1f | cta warp0-out %.1f all-done %.1f  (iter %d)" % (
                c, b, rel(per_chain[c, 0]), rel(per_chain[c, 1]), rel(cur[b, 0, 1]), rel(cur[b, 0, 2]), per_chain[c, 3]))
        print("tags: med %.1f p90 %.1f max %.1f" % tuple(np.percentile(rel(per_chain[:, 1]), [50, 90, 100])))
        cnt = np.bincount(fb, minlength=cur.shape[0])
        print("chains finished per CTA: ", np.bincount(cnt))
