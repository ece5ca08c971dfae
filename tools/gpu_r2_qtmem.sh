#!/bin/bash
# flash attention with the query tile in tensor memory (Q K^T as a TMEM-A instruction): parity tests, kernel TFLOP/s and encoder time A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_widths.py -q -x -k "attention or encoder or teacher" 2>&1 | tail -3
for q in 1 0; do
  echo "== TWB200_FA_QTMEM=$q"
  TWB200_FA_QTMEM=$q timeout 300 python tools/microbench.py encoder_attention 2>&1 | tail -3
  TWB200_FA_QTMEM=$q timeout 600 python tools/probe_encoder.py 2>&1 | tail -1
done
