#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "split or skinny_lite or decode_attention_kernel" > gpurun_out/split_tests.log 2>&1
echo "tests exit $?"; tail -5 gpurun_out/split_tests.log
timeout 400 python tools/probe_split.py 1 2 3 4 > gpurun_out/probe_split.log 2>&1
echo "probe exit $?"; tail -8 gpurun_out/probe_split.log
