#!/bin/bash
# absorbed cross-attention: key-tile size A/B (ring look-ahead): kernel tests + in-situ stream time from bench.py
mkdir -p gpurun_out
for tk in 32 48 64; do
  TWB200_AB_TK=$tk timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "absorbed_attention" 2>&1 | tail -2
  TWB200_AB_TK=$tk timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-parity --no-hf-cuda --no-ragged --no-e2e > gpurun_out/ab3_bench_tk$tk.json 2> gpurun_out/ab3_bench_tk$tk.err
  echo "bench tk=$tk exit $?"; tail -2 gpurun_out/ab3_bench_tk$tk.err
  python - <<PY
import json
d=json.load(open('gpurun_out/ab3_bench_tk$tk.json'))
print('tk $tk value',d['value'],'roof',d['roofline']['frac'], d['roofline']['avg_launch_us'])
print('stages', {k:(round(v['ms'],1),round(v['frac'],3)) for k,v in d['stages'].items()})
PY
done
TWB200_AB_TK=32 TWB200_AB_REV=0 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-parity --no-hf-cuda --no-ragged --no-e2e > gpurun_out/ab3_bench_tk32_norev.json 2> gpurun_out/ab3_bench_tk32_norev.err
python - <<PY
import json
d=json.load(open('gpurun_out/ab3_bench_tk32_norev.json'))
print('tk 32 norev value',d['value'],'roof',d['roofline']['frac'], d['roofline']['avg_launch_us'])
PY
