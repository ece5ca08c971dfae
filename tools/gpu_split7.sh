#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "split or decode_attention_kernel" > gpurun_out/split_tests.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/split_tests.log
(PROBE_BK=32 timeout 120 python tools/probe_coresident.py;  PROBE_BK=64 PROBE_LITE=0 timeout 120 python tools/probe_coresident.py) 2>&1 | grep -v Warn > gpurun_out/probe_cores3.log; cat gpurun_out/probe_cores3.log
timeout 600 python tools/probe_split.py "TWB200_SPLIT=1,TWB200_DA_STAGES=4" "TWB200_SPLIT=1,TWB200_DA_STAGES=3" "TWB200_SPLIT=2,TWB200_SPLIT_PDL=0" "TWB200_SPLIT=3,TWB200_SPLIT_PDL=0" "TWB200_SPLIT=2,TWB200_SPLIT_PDL=0,TWB200_TRACE=100" > gpurun_out/probe_split7.log 2>&1
echo "probe exit $?"; grep setting gpurun_out/probe_split7.log
