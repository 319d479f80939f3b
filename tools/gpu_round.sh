#!/bin/bash
# one GPU-box visit: GPU tests, the headline bench, ncu captures (summaries are copied to profiles/ afterwards)
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
python bench.py > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; tail -2 gpurun_out/bench_$1.err; cat gpurun_out/bench_$1.json
if [ "$2" = "ncu_panel" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:panel_sim_kernel -s 3 -c 1 -f -o gpurun_out/prof_$1_panel \
    python tools/bench_configs.py --config c4 --iters 3 --warmup 3 > gpurun_out/prof_$1_panel.log 2>&1
  tail -3 gpurun_out/prof_$1_panel.log
fi
