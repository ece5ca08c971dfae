#!/bin/bash
mkdir -p gpurun_out
(echo "== old (7bef087)"; timeout 300 python tools/probes/ab_old/tools/probe_split.py 1 2>&1 | grep setting
 echo "== new"; timeout 300 python tools/probe_split.py 1 2>&1 | grep setting
 echo "== new, split 3"; timeout 300 python tools/probe_split.py "TWB200_SPLIT=3" 2>&1 | grep setting) > gpurun_out/ab2.log 2>&1
cat gpurun_out/ab2.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "split or decode_attention_kernel or teacher_forced or bit_identical" 2>&1 | tail -2
