N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
mkdir -p gpurun_out
SMM_PHASE_TS=1 timeout 120 $TR tools/phase_timing_multi.py > gpurun_out/phase_r2m_n$N.txt 2>&1; tail -16 gpurun_out/phase_r2m_n$N.txt
