#!/usr/bin/env python3
"""Dump the four injected random streams (include/smm_stream.h) of a parity case for julia/parity_harness.jl, which
feeds them into the REAL SMM.jl `computeNextIteration!` and writes back a trace the pytest suite compares with the CUDA
path and the oracle (tests/test_julia_parity.py).  This is the route from "parity unpinned" to a reference pin.

    python tools/dump_streams.py                      # both cases -> tests/golden/julia/<case>/
    julia --project=<SMM.jl checkout> julia/parity_harness.jl tests/golden/julia/c1_serial_normal
    julia --project=<SMM.jl checkout> julia/parity_harness.jl tests/golden/julia/mvnormal_8chains
    python -m pytest tests/test_julia_parity.py       # oracle vs Julia here, -m gpu: CUDA vs Julia

Directory layout (raw little-endian arrays, C order; Julia reads them with the dimensions reversed):

    meta.txt       key=value: case, objective, n_chains N, n_params P, n_moments M, n_sim S, n_iter I, n_attempts A,
                   n_pairs n_s, maxtemp, sigma, sigma_update_steps, sigma_adjust_by, smpl_iters, batch_size,
                   seed_algo, seed_sim
    lb ub init     f64 [P]        data_mom data_w  f64 [M]        acc_tuner min_improve  f64 [N]
    zsim.f64       [P][S]         Zsim[row k, draw s]   (ziggurat stream of the MvNormal objectives)
    zprop.f64      [N][I][A][P]   Zprop[chain, iter, attempt, param]  (iter index 0 = iteration 1, unused)
    uacc.f64       [N][I]         Uacc[chain, iter] = BGPChain.probs_acc
    pairs.i32      [I][n_s][2]    Pairs[iter][t] = (i, j), 1-based, i < j  (rows of iteration 1 are zero: no exchange)

The streams are evaluated by the C++ restatement (oracle/) -- test infrastructure, like this script."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STREAM_PROP = 2
N_ATTEMPTS = 16          # rejection attempts dumped per (chain, iteration); the harness stops if a case needs more


def cases():
    from smm_jl_b200 import configs
    return {
        # SMM.serialNormal(2, 200) (Examples.jl:118-153, 373-446) and a small ensemble with C2's objective
        "c1_serial_normal": (configs.c1_serial_normal(200, smpl_iters=N_ATTEMPTS), "norm"),
        "mvnormal_8chains": (configs.mvnormal(8, 60, 4, smpl_iters=N_ATTEMPTS, n_sim=2000), "norm_mv"),
    }


def dump(case: str, cfg, objective: str, out_dir: str) -> None:
    from oracle import oracle_lib as ol
    os.makedirs(out_dir, exist_ok=True)
    N, P, M, S, I = cfg.n_chains, cfg.n_params, cfg.n_moments, cfg.n_sim, cfg.max_iter
    A = N_ATTEMPTS
    n_s = N - 1 if N < 3 else N
    bs = cfg.batch_size or P

    def w(name, arr, dt):
        np.ascontiguousarray(arr, dtype=dt).tofile(os.path.join(out_dir, name))

    for f in ("lb", "ub", "init", "data_mom", "data_w", "acc_tuner", "min_improve", "sigma0"):
        w(f + ".f64", np.asarray(getattr(cfg, f), float), "<f8")
    zsim = np.stack([ol.zig_normals(cfg.seed_sim, k, 0, 1 << 28, (S + 2) // 3)[:S] for k in range(P)])
    w("zsim.f64", zsim, "<f8")
    zprop = np.zeros((N, I, A, P))
    for c in range(N):
        for it in range(2, I + 1):
            for kp in range((P + 1) // 2):
                z = ol.normals(cfg.seed_algo, kp, c, (STREAM_PROP << 28) | it, A).reshape(A, 2)
                zprop[c, it - 1, :, 2 * kp] = z[:, 0]
                if 2 * kp + 1 < P:
                    zprop[c, it - 1, :, 2 * kp + 1] = z[:, 1]
    w("zprop.f64", zprop, "<f8")
    w("uacc.f64", np.array([[ol.acc_uniform(cfg.seed_algo, c, it) for it in range(1, I + 1)] for c in range(N)]), "<f8")
    pairs = np.zeros((I, n_s, 2), dtype=np.int32)
    for it in range(2, I + 1):
        pairs[it - 1] = np.asarray(ol.pairs(cfg.seed_algo, it, N), dtype=np.int32).reshape(n_s, 2) + 1
    w("pairs.i32", pairs, "<i4")
    meta = dict(case=case, objective=objective, n_chains=N, n_params=P, n_moments=M, n_sim=S, n_iter=I, n_attempts=A,
                n_pairs=n_s, maxtemp=5.0, sigma=float(np.asarray(cfg.sigma0)[0]), sigma_update_steps=cfg.sigma_update_steps,
                sigma_adjust_by=cfg.sigma_adjust_by, smpl_iters=min(cfg.smpl_iters, A), batch_size=bs,
                seed_algo=cfg.seed_algo, seed_sim=cfg.seed_sim)
    with open(os.path.join(out_dir, "meta.txt"), "w") as f:
        for k, v in meta.items():
            f.write(f"{k}={v!r}\n" if isinstance(v, float) else f"{k}={v}\n")


def main():
    base = os.path.join(ROOT, "tests", "golden", "julia")
    only = sys.argv[1:] or None
    for case, (cfg, objective) in cases().items():
        if only and case not in only:
            continue
        dump(case, cfg, objective, os.path.join(base, case))
        print("wrote", os.path.join(base, case))


if __name__ == "__main__":
    main()
