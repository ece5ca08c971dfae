#!/bin/bash
mkdir -p gpurun_out
(echo "== combine 64-thread CTAs"; timeout 300 python tools/probe_split.py "TWB200_SPLIT=1,TWB200_COMBINE_HPC=1" 2>&1 | grep setting
 echo "== combine 256-thread CTAs"; timeout 300 python tools/probe_split.py "TWB200_SPLIT=1,TWB200_COMBINE_HPC=4" 2>&1 | grep setting
 echo "== combine 64-thread CTAs again"; timeout 300 python tools/probe_split.py "TWB200_SPLIT=1,TWB200_COMBINE_HPC=1" 2>&1 | grep setting) > gpurun_out/ab6.log 2>&1
cat gpurun_out/ab6.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "paged_self_attention or decode_attention_kernel" 2>&1 | tail -2
