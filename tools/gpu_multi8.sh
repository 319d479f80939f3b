#!/bin/bash
# 8-GPU (or N-GPU) visit, kept short (charged N x): the multi-GPU parity tests, the headline bench at N GPUs at the
# driver's length and at the default length (each line carries the parity leg and C3 / C4 / C5), phase stamps of rank 0
N=$1; TAG=$2
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu ) 2>&1 | tail -6 > gpurun_out/pytest_multi_${TAG}_n$N.txt; cat gpurun_out/pytest_multi_${TAG}_n$N.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_n${N}_driver.json 2> gpurun_out/bench_${TAG}_n$N.err; tail -3 gpurun_out/bench_${TAG}_n$N.err | cut -c1-300
timeout 300 $TR bench.py --gpus $N > gpurun_out/bench_${TAG}_n${N}.json 2> gpurun_out/bench_${TAG}_n$N.err; tail -3 gpurun_out/bench_${TAG}_n$N.err | cut -c1-300
SMM_PHASE_TS=1 timeout 120 $TR tools/phase_timing_multi.py > gpurun_out/phase_${TAG}_n$N.txt 2>&1; tail -30 gpurun_out/phase_${TAG}_n$N.txt
python - <<PY
import json
for f in ("gpurun_out/bench_${TAG}_n${N}_driver.json", "gpurun_out/bench_${TAG}_n${N}.json"):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches", "steps")}, "e2e", d["e2e"]["value"], d["e2e"]["seconds"], "parity ok:", d["parity"]["ok"], d["parity"]["world"])
    for k, v in (d.get("secondary") or {}).items():
        print("   ", k, {a: v.get(a) for a in ("value", "ms_per_step", "n_chains", "error", "efficiency_vs_sleep_floor")})
PY
