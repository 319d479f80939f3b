#!/bin/bash
# N-GPU bench only (driver length + default length + phase stamps), no tests: the cheap way to look at scaling
N=$1; TAG=$2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_n${N}_driver.json 2> gpurun_out/bench_${TAG}_n$N.err; tail -3 gpurun_out/bench_${TAG}_n$N.err | cut -c1-300
timeout 300 $TR bench.py --gpus $N --no-secondary > gpurun_out/bench_${TAG}_n${N}.json 2> gpurun_out/bench_${TAG}_n$N.err; tail -3 gpurun_out/bench_${TAG}_n$N.err | cut -c1-300
SMM_PHASE_TS=1 timeout 120 $TR tools/phase_timing_multi.py > gpurun_out/phase_${TAG}_n$N.txt 2>&1; tail -12 gpurun_out/phase_${TAG}_n$N.txt
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
