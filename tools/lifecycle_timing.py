"""Where the wall time of one complete run!(MAlgoBGP) goes outside the kernels: Python constructor, smm_bgp_create
(SMM_TIMING=1 prints its stamps), run, read, destroy -- at the driver's bench length (25 iterations) and at 1000."""
import os, sys, time
os.environ.setdefault("SMM_TIMING", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from smm_jl_b200 import api, configs

cfg = configs.mvnormal(256, 25, 8)
m = api.MProb()
for k in range(8):
    api.addSampledParam(m, f"p{k + 1}", cfg.init[k], cfg.lb[k], cfg.ub[k])
for k in range(16):
    api.addMoment(m, f"m{k + 1}", cfg.data_mom[k], cfg.data_w[k])
api.addEvalFunc(m, api.objfunc_norm_mv)
for n_iter in (25, 25, 25, 25, 1000, 1000, 25):
    opts = {"N": 256, "maxtemp": 5.0, "sigma": 0.05, "acc_tuners": list(np.asarray(cfg.acc_tuner)), "min_improve": [0.0] * 256,
            "smpl_iters": cfg.smpl_iters, "seed": cfg.seed_algo, "maxiter": n_iter}
    t0 = time.perf_counter()
    algo = api.MAlgoBGP(m, opts)
    t1 = time.perf_counter()
    algo._handle()
    t2 = time.perf_counter()
    api.run(algo)
    t3 = time.perf_counter()
    best = float(algo._streamed.best_val[n_iter - 1].min())
    t4 = time.perf_counter()
    algo.close()
    t5 = time.perf_counter()
    print(f"maxiter {n_iter}: MAlgoBGP() {1e6*(t1-t0):.0f} us, smm_bgp_create {1e6*(t2-t1):.0f} us, run! {1e6*(t3-t2):.0f} us "
          f"(device {1e3*algo.device_ms:.0f} us), read {1e6*(t4-t3):.0f} us, close {1e6*(t5-t4):.0f} us, total {1e6*(t5-t0):.0f} us "
          f"-> {256*n_iter/(t5-t0)/1e6:.2f} M evals/s", flush=True)
