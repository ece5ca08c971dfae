#!/bin/bash
# pipeline timeline of the absorbed cross-attention kernel (CTA 0, clock64), bench shape
mkdir -p gpurun_out
for tk in 64 32; do
TWB200_AB_TK=$tk TWB200_AB_TRACE=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -s -k "absorbed_attention_kernel_vs_torch and 1500-64-20 and 0-" > gpurun_out/ab5_trace_tk$tk.log 2>&1
grep -c "ab trace" gpurun_out/ab5_trace_tk$tk.log
done
