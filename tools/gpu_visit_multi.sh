#!/bin/bash
# one multi-GPU box visit: GPU tests (incl. world-N oracle parity), bench.py under torchrun with the parity leg
TAG=$1; N=$2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader | head -8; nproc
( time timeout 900 python -m pytest tests -q -m gpu -rs ) 2>&1 | tail -30 > gpurun_out/pytest_gpu_${TAG}_n$N.txt; cat gpurun_out/pytest_gpu_${TAG}_n$N.txt
for M in ${MODES:-1}; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps ${STEPS:-400} --warmup 20 --exchange-mode $M > gpurun_out/bench_${TAG}_n${N}_m$M.json 2> gpurun_out/bench_${TAG}_n${N}_m$M.err
tail -3 gpurun_out/bench_${TAG}_n${N}_m$M.err; cat gpurun_out/bench_${TAG}_n${N}_m$M.json
done
