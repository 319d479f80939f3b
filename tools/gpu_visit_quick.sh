TAG=$1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu -x ) 2>&1 | tail -25 > gpurun_out/pytest_gpu_$TAG.txt; cat gpurun_out/pytest_gpu_$TAG.txt
timeout 120 python tools/sim_variants.py > gpurun_out/variants_$TAG.txt 2>&1; cat gpurun_out/variants_$TAG.txt | tail -30
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -2 gpurun_out/bench_$TAG.err; python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$TAG.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['parity']['ok'], d['roofline']['frac'])"
timeout 120 python tools/phase_timing.py 256 2 > gpurun_out/phase_$TAG.txt 2>&1; cat gpurun_out/phase_$TAG.txt
