#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/fa_variants3.log
for v in 1 8; do
  echo "== variant $v" >> gpurun_out/fa_variants3.log
  TWB200_FA_VARIANT=$v timeout 120 python tools/microbench.py encoder_attention 2>&1 | grep "encoder_attention(tc" >> gpurun_out/fa_variants3.log
  TWB200_FA_VARIANT=$v timeout 300 python tools/probe_encoder.py 2>&1 | grep encoder >> gpurun_out/fa_variants3.log
done
echo "== tests variant 8" >> gpurun_out/fa_variants3.log
TWB200_FA_VARIANT=8 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "encoder_attention_kernels or encoder_bf16 or general_attention or teacher_logits_bf16" 2>&1 | tail -3 >> gpurun_out/fa_variants3.log
cat gpurun_out/fa_variants3.log
