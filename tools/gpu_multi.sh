#!/bin/bash
# one multi-GPU box visit: N-GPU parity tests, the headline bench at N GPUs, the secondary configs (C3/C4/C5)
N=$1; TAG=$2
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -8
python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_multi_$TAG.txt; cat gpurun_out/pytest_multi_$TAG.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 450 --warmup 50 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; tail -3 gpurun_out/bench_${TAG}_n$N.err; cat gpurun_out/bench_${TAG}_n$N.json
$TR tools/bench_configs.py --config c3 > gpurun_out/c3_${TAG}_n$N.json 2> gpurun_out/c3.err; tail -2 gpurun_out/c3.err; cat gpurun_out/c3_${TAG}_n$N.json
$TR tools/bench_configs.py --config c3 --exchange-mode 0 --iters 200 > gpurun_out/c3_nccl_${TAG}_n$N.json 2> gpurun_out/c3n.err; tail -2 gpurun_out/c3n.err; cat gpurun_out/c3_nccl_${TAG}_n$N.json
$TR tools/bench_configs.py --config c4 --iters 40 > gpurun_out/c4_${TAG}_n$N.json 2> gpurun_out/c4.err; tail -2 gpurun_out/c4.err; cat gpurun_out/c4_${TAG}_n$N.json
$TR tools/bench_configs.py --config c5 --iters 10 > gpurun_out/c5_${TAG}_n$N.json 2> gpurun_out/c5.err; tail -2 gpurun_out/c5.err; cat gpurun_out/c5_${TAG}_n$N.json
