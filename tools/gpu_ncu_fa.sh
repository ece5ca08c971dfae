#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ncu_encoder_small.py 16 > gpurun_out/fa_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/fa_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encoder_attention_tc -s 1 -c 1 -f -o gpurun_out/prof_fa python tools/ncu_encoder_small.py 16 > gpurun_out/fa_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/fa_ncu.log; ls -la gpurun_out/prof_fa.ncu-rep
