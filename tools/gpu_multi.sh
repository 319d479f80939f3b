#!/bin/bash
# one multi-GPU box visit: N-GPU parity tests, the headline bench at N GPUs (driver length and default length; each
# line carries the parity leg and the secondary block C3 / C4 / C5), phase stamps of one rank
N=$1; TAG=$2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -8
( time timeout 900 python -m pytest tests -q -m gpu ) 2>&1 | tail -12 > gpurun_out/pytest_gpu_${TAG}_n$N.txt; cat gpurun_out/pytest_gpu_${TAG}_n$N.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_n${N}_driver.json 2> gpurun_out/bench_${TAG}_n$N.err; tail -3 gpurun_out/bench_${TAG}_n$N.err | cut -c1-300
timeout 300 $TR bench.py --gpus $N > gpurun_out/bench_${TAG}_n${N}.json 2> gpurun_out/bench_${TAG}_n$N.err; tail -3 gpurun_out/bench_${TAG}_n$N.err | cut -c1-300
python - <<PY
import json
for f in ("gpurun_out/bench_${TAG}_n${N}_driver.json", "gpurun_out/bench_${TAG}_n${N}.json"):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches", "steps")}, "e2e", d["e2e"]["value"], d["e2e"]["seconds"], d["parity"])
    print("  secondary:", json.dumps(d.get("secondary"))[:1800])
PY
