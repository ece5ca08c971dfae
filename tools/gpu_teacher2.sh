#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "teacher or general_attention or encoder_attention_kernels or encoder_bf16" > gpurun_out/teacher_tests.log 2>&1
echo "tests exit $?"; tail -6 gpurun_out/teacher_tests.log
timeout 300 python tools/bench_teacher.py > gpurun_out/bench_teacher.log 2>&1; tail -2 gpurun_out/bench_teacher.log
TWB200_ATTN=simt timeout 300 python tools/bench_teacher.py 2>&1 | tail -1 | tee -a gpurun_out/bench_teacher.log
timeout 120 python tools/microbench.py encoder_attention 2>&1 | grep "tcgen05"
