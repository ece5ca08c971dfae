#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/fa_variants.log
for v in 0 1 2 3 4 6 7; do
  echo "== variant $v" >> gpurun_out/fa_variants.log
  TWB200_FA_VARIANT=$v timeout 120 python tools/microbench.py encoder_attention 2>&1 | grep "encoder_attention(tc" >> gpurun_out/fa_variants.log
done
for v in 2 3 7; do
  echo "== tests variant $v" >> gpurun_out/fa_variants.log
  TWB200_FA_VARIANT=$v timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "encoder_attention_kernels or encoder_bf16" 2>&1 | tail -2 >> gpurun_out/fa_variants.log
done
cat gpurun_out/fa_variants.log
