#!/bin/bash
# one single-GPU box visit: GPU tests, headline bench (driver length + default), reference arm, phase stamps, inner-loop
# ceiling, iteration profile, then the ncu launch list of bench.py and `ncu --set full` captures (persistent kernel +
# inner loop, panel kernel) into gpurun_out/
TAG=$1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader | head -2; nproc
( time timeout 900 python -m pytest tests -q -m gpu ) 2>&1 | tail -25 > gpurun_out/pytest_gpu_$TAG.txt; cat gpurun_out/pytest_gpu_$TAG.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_driver.json 2> gpurun_out/bench_$TAG.err; tail -2 gpurun_out/bench_$TAG.err
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -2 gpurun_out/bench_$TAG.err
python - <<PY
import json
for f in ("gpurun_out/bench_${TAG}_driver.json", "gpurun_out/bench_$TAG.json"):
    d = json.load(open(f))
    print(f, {k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "steps")}, "e2e", d["e2e"]["value"], d["parity"]["ok"], d["roofline"]["frac"], (d.get("cpu_baseline") or {}).get("value"))
PY
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; cut -c1-200 gpurun_out/bench_ref_$TAG.json
timeout 120 python tools/phase_timing.py 256 2 > gpurun_out/phase_$TAG.txt 2>&1; head -12 gpurun_out/phase_$TAG.txt
timeout 120 python tools/sim_ceiling.py > gpurun_out/ceiling_$TAG.txt 2>&1; tail -5 gpurun_out/ceiling_$TAG.txt
timeout 120 python tools/lifecycle_timing.py > gpurun_out/lifecycle_$TAG.txt 2>&1; grep maxiter gpurun_out/lifecycle_$TAG.txt | tail -3
timeout 600 python tools/iter_profile.py 2 > gpurun_out/iter_profile_$TAG.txt 2>&1; tail -8 gpurun_out/iter_profile_$TAG.txt
if [ "$2" != "noncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-secondary > gpurun_out/launches_$TAG.log 2>&1; tail -2 gpurun_out/launches_$TAG.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bgp_persistent_kernel|sim_throughput_kernel" -s 1 -c 3 -f -o gpurun_out/prof_${TAG}_persistent \
  python tools/profile_target.py 2 > gpurun_out/prof_$TAG.log 2>&1; tail -3 gpurun_out/prof_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:panel_sim_kernel -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_panel \
  python tools/bench_configs.py --config c4 --iters 3 --warmup 3 > gpurun_out/prof_${TAG}_panel.log 2>&1; tail -3 gpurun_out/prof_${TAG}_panel.log
fi
