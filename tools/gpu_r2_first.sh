#!/bin/bash
# round 2, first call: in-situ decode timeline (large-v3, 64 clips), log-mel timing, ncu --set full of logmel_kernel and gemm_tc_kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_first_smi.txt
TWB200_TRACE=100 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_trace_bench.json 2> gpurun_out/r2_trace.err
echo "trace bench exit $?"; grep -c "twb200 trace" gpurun_out/r2_trace.err
timeout 300 python tools/ncu_logmel.py > gpurun_out/r2_logmel_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:logmel -s 4 -c 2 -o gpurun_out/r2_logmel python tools/ncu_logmel.py > gpurun_out/r2_logmel_ncu.log 2>&1
echo "logmel ncu exit $?"; cat gpurun_out/r2_logmel_plain.log
timeout 300 python tools/ncu_encoder_small.py 16 > gpurun_out/r2_enc_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 10 -c 3 -o gpurun_out/r2_gemm_tc python tools/ncu_encoder_small.py 16 > gpurun_out/r2_gemm_ncu.log 2>&1
echo "gemm ncu exit $?"; tail -1 gpurun_out/r2_enc_plain.log
