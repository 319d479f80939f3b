TAG=$1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu -x ) 2>&1 | tail -25 > gpurun_out/pytest_gpu_$TAG.txt; cat gpurun_out/pytest_gpu_$TAG.txt
timeout 120 python tools/lifecycle_timing.py > gpurun_out/lifecycle_$TAG.txt 2>&1; tail -40 gpurun_out/lifecycle_$TAG.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_driver.json 2> gpurun_out/bench_$TAG.err; tail -2 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -2 gpurun_out/bench_$TAG.err
python - <<PY
import json
for f in ("gpurun_out/bench_${TAG}_driver.json", "gpurun_out/bench_$TAG.json"):
    d = json.load(open(f))
    print(f, {k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "steps")}, "e2e", d["e2e"]["value"], d["e2e"]["seconds"], d["parity"]["ok"], d["roofline"]["frac"])
    print("  secondary:", json.dumps(d.get("secondary"))[:1500])
PY
