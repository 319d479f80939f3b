#!/usr/bin/env python3
"""Secondary BASELINE.json configurations (bench.py itself measures C2, the headline):

    python tools/bench_configs.py --config c4 [--chains-per-gpu 64] [--iters 60] [--warmup 5]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_configs.py --config c3

  c3  MvNormal SMM, 1024 chains over the GPUs of the box (128 per GPU at 8), fused peer-store all-gather
  c4  dynamic-panel SMM (K = 8, T = 50, 5000 individuals), 64 chains per GPU, ncclAllGather per iteration
  c5  slow objective (0.1 s per evaluation), 64 chains per GPU: iteration-time overhead over 0.1 s

Timing: CUDA events on the library's stream around exactly `iters` iterations after `warmup`, max over ranks.
Prints one JSON line per run (rank 0)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=["c3", "c4", "c5"])
    ap.add_argument("--chains-per-gpu", type=int, default=0)
    ap.add_argument("--iters", type=int, default=0)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--exchange-mode", type=int, default=-1)
    ap.add_argument("--slow-seconds", type=float, default=0.1)
    ap.add_argument("--cpu-sample", action="store_true", help="also time the oracle on a bounded sample (rank 0)")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from smm_jl_b200 import _lib, configs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def fresh_id() -> bytes:
        if world == 1:
            return b""
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(_lib.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        return bytes(idt.cpu().tolist())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W = max(args.warmup, 3)
    if args.config == "c3":
        cpg = args.chains_per_gpu or (1024 // world if world > 1 else 128)
        K = args.iters or 500
        mode = 1 if args.exchange_mode < 0 else args.exchange_mode
        cfg = configs.mvnormal(cpg * world, K + W, exchange_mode=mode)
        name = f"C3: MvNormal SMM, {cpg * world} chains ({cpg}/GPU), 8 params, 16 moments, 10k draws/eval"
        bytes_per_eval = 2 * 8 * 8 * 10000 + 8 * (8 + 16 + 4) + 13
    elif args.config == "c4":
        cpg = args.chains_per_gpu or 64
        K = args.iters or 60
        dm = configs.panel_data_moments_gpu(8, 50, 5000, device=local_rank)   # every rank computes the same bits
        cfg = configs.dynamic_panel(cpg * world, K + W, data_mom=dm)
        name = f"C4: dynamic-panel SMM, {cpg * world} chains ({cpg}/GPU), 20 params, 40 moments, T=50 x N=5000"
        bytes_per_eval = 2 * 8 * 9 * 5000 * 50 + 8 * (20 + 40 + 4) + 13
    else:
        cpg = args.chains_per_gpu or 64
        K = args.iters or 20
        mode = 1 if args.exchange_mode < 0 else args.exchange_mode
        cfg = configs.slow_normal(cpg * world, K + W, slow_seconds=args.slow_seconds, exchange_mode=mode)
        name = f"C5: slow objective ({args.slow_seconds} s/eval), {cpg * world} chains ({cpg}/GPU)"
        bytes_per_eval = 2 * 8 * 2 * 10000 + 8 * (2 + 2 + 4) + 13
    cfg.device, cfg.world_size, cfg.rank, cfg.nccl_id = local_rank, world, rank, fresh_id()
    h = _lib.BGPHandle(cfg)
    h.step(W)
    l0 = h.counters()["kernel_launches"]
    barrier()
    t0 = time.perf_counter()
    ms = h.step(K)
    barrier()
    wall = time.perf_counter() - t0
    ctr = h.counters()
    h.close()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    n_chains = cpg * world
    value = n_chains * K / (ms * 1e-3)
    line = {"config": args.config, "workload": name, "metric": "objective-evals/sec (all chains)", "value": value,
            "unit": "evals/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "wall_seconds": wall, "gpu_launches": int(ctr["kernel_launches"] - l0), "exchange_mode": cfg.exchange_mode,
            "algorithmic_GBps_per_gpu": value / world * bytes_per_eval / 1e9, "swaps": ctr["swaps"],
            "accept_rate_mean": ctr["accepted"] / max(ctr["evaluations"], 1)}
    if args.config == "c5":
        line["iteration_overhead_ms_over_sleep"] = ms / K - 1e3 * args.slow_seconds
        line["efficiency_vs_sleep_floor"] = 1e3 * args.slow_seconds / (ms / K)
    if args.cpu_sample and rank == 0:
        from oracle import oracle_lib
        import copy
        c = copy.copy(cfg)
        c.world_size, c.rank, c.n_chains = 1, 0, cpg
        for f in ("sigma0", "acc_tuner", "min_improve"):
            setattr(c, f, np.asarray(getattr(cfg, f))[:cpg])
        it = 2 if args.config != "c3" else 20
        t0 = time.perf_counter()
        oracle_lib.run(c, it, n_threads=os.cpu_count() or 1)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": cpg * it / dt, "unit": "evals/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"{cpg} chains x {it} iterations, {dt:.1f} s"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
