#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "split or decode_attention_kernel or teacher or bit_identical or golden or transcribe_host" > gpurun_out/split_tests.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/split_tests.log
timeout 600 python tools/probe_split.py "TWB200_SPLIT=1" > gpurun_out/probe_split8.log 2>&1
echo "probe exit $?"; grep setting gpurun_out/probe_split8.log
