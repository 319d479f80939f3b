#!/bin/bash
# one single-GPU box visit: GPU tests, headline bench, reference arm, phase stamps, secondary configs, ncu captures
TAG=$1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader | head -2; nproc
( time timeout 900 python -m pytest tests -q -m gpu ) 2>&1 | tail -25 > gpurun_out/pytest_gpu_$TAG.txt; cat gpurun_out/pytest_gpu_$TAG.txt
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -2 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; cat gpurun_out/bench_ref_$TAG.json
timeout 120 python tools/phase_timing.py 256 2 > gpurun_out/phase_$TAG.txt 2>&1; cat gpurun_out/phase_$TAG.txt
timeout 120 python tools/sim_ceiling.py > gpurun_out/ceiling_$TAG.txt 2>&1; tail -5 gpurun_out/ceiling_$TAG.txt
timeout 300 python tools/bench_configs.py --config c4 --iters 40 --cpu-sample > gpurun_out/c4_$TAG.json 2> gpurun_out/c4.err; tail -2 gpurun_out/c4.err; cat gpurun_out/c4_$TAG.json
timeout 300 python tools/bench_configs.py --config c5 --iters 10 > gpurun_out/c5_$TAG.json 2> gpurun_out/c5.err; tail -2 gpurun_out/c5.err; cat gpurun_out/c5_$TAG.json
if [ "$2" = "ncusim" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bgp_persistent_kernel|sim_throughput_kernel" -s 1 -c 3 -f -o gpurun_out/prof_${TAG}_persistent \
  python tools/profile_target.py 2 > gpurun_out/prof_$TAG.log 2>&1; tail -3 gpurun_out/prof_$TAG.log
elif [ "$2" != "noncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/launches_$TAG.log 2>&1; tail -2 gpurun_out/launches_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bgp_persistent_kernel|sim_throughput_kernel" -s 1 -c 3 -f -o gpurun_out/prof_${TAG}_persistent \
  python tools/profile_target.py 2 > gpurun_out/prof_$TAG.log 2>&1; tail -3 gpurun_out/prof_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:panel_sim_kernel -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_panel \
  python tools/bench_configs.py --config c4 --iters 3 --warmup 3 > gpurun_out/prof_${TAG}_panel.log 2>&1; tail -3 gpurun_out/prof_${TAG}_panel.log
fi
ls -la gpurun_out | tail -20
