#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "split or decode_attention_kernel or first_pass" > gpurun_out/split_tests.log 2>&1
echo "tests exit $?"; tail -4 gpurun_out/split_tests.log
(PROBE_BK=32 timeout 120 python tools/probe_coresident.py; PROBE_BK=64 PROBE_LITE=0 timeout 120 python tools/probe_coresident.py) 2>&1 | grep -v Warn | grep "stream kernel:\|gemm lite (impl\|self-att" > gpurun_out/probe_cores4.log; cat gpurun_out/probe_cores4.log
timeout 900 python tools/probe_split.py "TWB200_SPLIT=1" "TWB200_SPLIT=2,TWB200_SPLIT_PDL=0" "TWB200_SPLIT=3,TWB200_SPLIT_PDL=0" "TWB200_SPLIT=4,TWB200_SPLIT_PDL=0" "TWB200_SPLIT=2,TWB200_SPLIT_PDL=1" "TWB200_SPLIT=3,TWB200_SPLIT_PDL=1" > gpurun_out/probe_split9.log 2>&1
echo "probe exit $?"; grep setting gpurun_out/probe_split9.log
