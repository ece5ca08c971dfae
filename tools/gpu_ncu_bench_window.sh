#!/bin/bash
# ncu launch list (gpu__time_duration) of a window of the bench command's timed step.
# TWB200_GRAPH=0 so that every kernel launch of the decode loop is a separate launch record.
mkdir -p gpurun_out
export TWB200_GRAPH=0
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 250 $CMD > gpurun_out/window_plain.log 2>&1 || { echo "plain run failed"; tail -3 gpurun_out/window_plain.log; exit 1; }
# 3 warm-up steps x ~99.5k launches each precede the timed step; land 20k launches into its decode loop
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k regex:"gemm_tc|decode_attention|self_attention|layernorm|select_tokens|embed_kernel|advance|encoder_attention|im2col|logmel" \
  -s 320000 -c 2000 --csv --log-file gpurun_out/bench_window_launches.csv $CMD > gpurun_out/window_ncu.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/bench_window_launches.csv; tail -2 gpurun_out/window_ncu.log
