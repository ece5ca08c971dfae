#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
echo "bench exit $?"; tail -3 gpurun_out/bench1.err; cat gpurun_out/bench1.json
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "ref exit $?"; tail -3 gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
