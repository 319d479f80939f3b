mkdir -p gpurun_out
for V in 2 6 7 8; do
  echo "== variant $V"
  SMM_PANEL_VARIANT=$V timeout 300 python -m pytest tests/test_gpu_panel.py -q -m gpu -x 2>&1 | tail -2
  SMM_PANEL_VARIANT=$V timeout 300 python tools/bench_configs.py --config c4 --iters 40 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
done
