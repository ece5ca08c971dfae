#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/fa_variants4.log
echo "== tests variant 2" >> gpurun_out/fa_variants4.log
TWB200_FA_VARIANT=2 timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "encoder_attention_kernels or encoder_bf16 or general_attention or teacher_logits_bf16" 2>&1 | tail -4 >> gpurun_out/fa_variants4.log
for v in 1 2; do
  echo "== variant $v" >> gpurun_out/fa_variants4.log
  TWB200_FA_VARIANT=$v timeout 120 python tools/microbench.py encoder_attention 2>&1 | grep "encoder_attention(tc" >> gpurun_out/fa_variants4.log
  TWB200_FA_VARIANT=$v timeout 200 python tools/probe_encoder.py 2>&1 | grep encoder >> gpurun_out/fa_variants4.log
done
cat gpurun_out/fa_variants4.log
