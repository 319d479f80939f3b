mkdir -p gpurun_out
V=${1:-7}
SMM_PANEL_VARIANT=$V timeout 600 ncu --set full --clock-control none --import-source on -k regex:"panel_lanes_kernel|panel_sim_kernel" -s 3 -c 1 -f -o gpurun_out/prof_panel_v$V \
    python tools/bench_configs.py --config c4 --iters 3 --warmup 3 > gpurun_out/prof_panel_v$V.log 2>&1; tail -3 gpurun_out/prof_panel_v$V.log
