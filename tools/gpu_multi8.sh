#!/bin/bash
# 8-GPU smoke: the headline bench (weak scaling, 256 chains/GPU) and the C3 / C4 shapes of BASELINE.json
N=$1; TAG=$2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 240 $TR bench.py --gpus $N --steps 450 --warmup 50 > gpurun_out/bench_${TAG}_n${N}.json 2> gpurun_out/bench_${TAG}_n$N.err; tail -3 gpurun_out/bench_${TAG}_n$N.err | cut -c1-300; cut -c1-300 gpurun_out/bench_${TAG}_n${N}.json
timeout 120 $TR tools/bench_configs.py --config c3 --exchange-mode 1 --iters 300 > gpurun_out/c3_${TAG}_n$N.json 2> gpurun_out/c3.err; tail -2 gpurun_out/c3.err | cut -c1-300; cat gpurun_out/c3_${TAG}_n$N.json
timeout 120 $TR tools/bench_configs.py --config c4 --iters 40 > gpurun_out/c4_${TAG}_n$N.json 2> gpurun_out/c4.err; tail -2 gpurun_out/c4.err | cut -c1-300; cat gpurun_out/c4_${TAG}_n$N.json
