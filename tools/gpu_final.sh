# HEAD-state check the way the driver does it: smoke(), pytest -m gpu, bench at the driver's length and the reference arm
TAG=$1
mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -4 | tee gpurun_out/smoke_$TAG.txt
( time timeout 900 python -m pytest tests -q -m gpu ) 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_$TAG.txt
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_driver.json 2> gpurun_out/bench_$TAG.err; tail -2 gpurun_out/bench_$TAG.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}_driver.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "steps")}, "e2e", d["e2e"]["value"], d["parity"]["ok"], d["roofline"]["frac"], (d.get("cpu_baseline") or {}).get("value"))
PY
