#!/bin/bash
# launch list (per-launch device time) of a short run of the bench command
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --max-length 16 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_small.csv $CMD > gpurun_out/ncu.log 2>&1
echo "exit $?"; tail -2 gpurun_out/plain.log | cut -c1-600; wc -l gpurun_out/launches_small.csv
