N=$1
mkdir -p gpurun_out
if [ "$2" = "tests" ]; then ( timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x ) 2>&1 | tail -4; fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for M in 3 2; do
  timeout 200 $TR bench.py --gpus $N --no-secondary --exchange-mode $M > gpurun_out/bench_m${M}_n$N.json 2> gpurun_out/b.err; tail -2 gpurun_out/b.err | cut -c1-300
done
SMM_PHASE_MODE=3 SMM_PHASE_TS=1 timeout 120 $TR tools/phase_timing_multi.py > gpurun_out/phase_m3_n$N.txt 2>&1; tail -14 gpurun_out/phase_m3_n$N.txt
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_m*_n$N.json")):
    try:
        d = json.load(open(f)); print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", d["parity"]["ok"], d["parity"]["mode"])
    except Exception as e: print(f, "unreadable", e)
PY
