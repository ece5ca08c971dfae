#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "attention or teacher_logits or encoder_bf16" > gpurun_out/r2_fa3_tests.log 2>&1
rc=$?; echo "attention tests exit $rc"; tail -4 gpurun_out/r2_fa3_tests.log
if [ $rc -ne 0 ]; then grep -n "Error\|assert \|^E " gpurun_out/r2_fa3_tests.log | head -20; fi
for v in 3 4; do TWB200_FA_VARIANT=$v timeout 120 python tools/microbench.py encoder_attention 2>&1 | grep tcgen05 | sed "s/^/FA$v /"; done
for v in 3 4; do TWB200_FA_VARIANT=$v timeout 300 python tools/probe_encoder.py 2>&1 | tail -1; done
