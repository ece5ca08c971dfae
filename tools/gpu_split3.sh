#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "split or decode_attention_kernel" > gpurun_out/split_tests.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/split_tests.log
timeout 600 python tools/probe_split.py "TWB200_SPLIT=2,TWB200_LITE=1,TWB200_TRACE=100" "TWB200_SPLIT=1,TWB200_LITE=1,TWB200_TRACE=-1" "TWB200_SPLIT=2,TWB200_LITE=1,TWB200_TRACE=-1" "TWB200_SPLIT=3,TWB200_LITE=1,TWB200_TRACE=-1"  > gpurun_out/probe_split3.log 2>&1
echo "probe exit $?"; grep setting gpurun_out/probe_split3.log
