#!/bin/bash
# attention v2 (P and O in tensor memory): unit tests vs torch under a timeout, kernel throughput, encoder stage A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "attention or teacher_logits or encoder_bf16" > gpurun_out/r2_fa_tests.log 2>&1
rc=$?; echo "attention tests exit $rc"; tail -5 gpurun_out/r2_fa_tests.log
if [ $rc -ne 0 ]; then grep -n "Error\|assert\|err" gpurun_out/r2_fa_tests.log | head -20; fi
for v in 1 2; do TWB200_FA_VARIANT=$v timeout 120 python tools/microbench.py encoder_attention 2>&1 | grep tcgen05; done
for v in 1 2; do TWB200_FA_VARIANT=$v timeout 300 python tools/probe_encoder.py 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_gpu_widths.py -x -q -m gpu -k "row_budgets or benched_width_bf16" > gpurun_out/r2_fa_tests2.log 2>&1
echo "width tests exit $?"; tail -3 gpurun_out/r2_fa_tests2.log
