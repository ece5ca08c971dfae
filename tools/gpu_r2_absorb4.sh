#!/bin/bash
# absorbed cross-attention with independent accumulators: kernel tests + in-situ stream time from bench.py per tile size
mkdir -p gpurun_out
for tk in 64 32; do
  TWB200_AB_TK=$tk timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "absorbed_attention" 2>&1 | tail -2
  TWB200_AB_TK=$tk timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-parity --no-hf-cuda --no-ragged --no-e2e > gpurun_out/ab4_bench_tk$tk.json 2> gpurun_out/ab4_bench_tk$tk.err
  echo "bench tk=$tk exit $?"; tail -2 gpurun_out/ab4_bench_tk$tk.err
  python - <<PY
import json
d=json.load(open('gpurun_out/ab4_bench_tk$tk.json'))
print('tk $tk value',d['value'],'roof',d['roofline']['frac'], d['roofline']['avg_launch_us'], d['config']['stage_ms_last_step'])
PY
done
