#!/bin/bash
# round-2 ncu evidence for the decode side: --set full of the dominant kernel, and the launch list of decode steps (2 layers)
mkdir -p gpurun_out
timeout 200 python tools/ncu_decode_attention.py > gpurun_out/r2_da_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:decode_attention_stream -s 1 -c 2 -o gpurun_out/r2_decode_attention_stream python tools/ncu_decode_attention.py > gpurun_out/r2_da_ncu.log 2>&1
echo "stream ncu exit $?"
timeout 300 python tools/ncu_decode_small.py 20 > gpurun_out/r2_ds_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_decode_step_launches.csv python tools/ncu_decode_small.py 20 > gpurun_out/r2_ds_ncu.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/r2_decode_step_launches.csv
