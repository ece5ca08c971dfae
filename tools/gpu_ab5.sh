#!/bin/bash
mkdir -p gpurun_out
(echo "== previous build"; timeout 300 python tools/probes/ab_old/tools/probe_split.py 1 2>&1 | grep setting
 echo "== paged cache, table row in smem"; timeout 300 python tools/probe_split.py 1 2>&1 | grep setting) > gpurun_out/ab5.log 2>&1
cat gpurun_out/ab5.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bit_identical or teacher_forced or golden or split or self_attention" 2>&1 | tail -2
