#!/bin/bash
# log-mel frames-per-tile A/B, log-mel tests, ncu --set full of the attention v2 kernel (source-level stall reasons)
mkdir -p gpurun_out
for fr in 32 16; do TWB200_LM_FR=$fr timeout 120 python tools/ncu_logmel.py 2>&1 | tail -1 | sed "s/^/FR=$fr /"; done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "logmel or transcribe_host or pipeline or encoder_fp32" > gpurun_out/r2_lm_tests.log 2>&1
echo "logmel tests exit $?"; tail -3 gpurun_out/r2_lm_tests.log
timeout 300 python tools/ncu_encoder_small.py 16 > gpurun_out/r2_enc_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:encoder_attention_tc2 -s 1 -c 1 -o gpurun_out/r2_fa2 python tools/ncu_encoder_small.py 16 > gpurun_out/r2_fa2_ncu.log 2>&1
echo "fa2 ncu exit $?"
