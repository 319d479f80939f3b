#!/bin/bash
# 2-GPU visit: multi-GPU parity tests (modes 0, 1, 2) and the headline bench at 2 GPUs for both persistent modes
N=$1; TAG=$2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -8
(timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -8) | tee gpurun_out/pytest_multi_$TAG.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for MODE in 2 1; do
  timeout 300 $TR bench.py --gpus $N --steps 450 --warmup 50 --exchange-mode $MODE > gpurun_out/bench_${TAG}_n${N}_m$MODE.json 2> gpurun_out/bench_${TAG}_n$N.err; tail -3 gpurun_out/bench_${TAG}_n$N.err | cut -c1-300; cut -c1-400 gpurun_out/bench_${TAG}_n${N}_m$MODE.json
done
timeout 300 $TR tools/bench_configs.py --config c3 --exchange-mode 2 > gpurun_out/c3_${TAG}_n$N.json 2> gpurun_out/c3.err; tail -2 gpurun_out/c3.err; cat gpurun_out/c3_${TAG}_n$N.json
