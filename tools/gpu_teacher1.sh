#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "teacher or encoder_attention_kernels or encoder_fp32 or encoder_bf16" > gpurun_out/teacher_tests.log 2>&1
echo "tests exit $?"; tail -25 gpurun_out/teacher_tests.log
