#!/bin/bash
# round 2: -m gpu suite after the active-row / row-budget change + bench with the ragged variant
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/r2_pytest_gpu2.log 2>&1
echo "pytest exit $?"; tail -14 gpurun_out/r2_pytest_gpu2.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-parity --no-hf-cuda > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
echo "bench exit $?"; tail -3 gpurun_out/r2_bench_b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_b.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'], d['roofline']['avg_launch_us'])
print('ragged',d.get('ragged')); print('stages', {k:(round(v['ms'],1),round(v['frac'],3)) for k,v in d['stages'].items()})
PY
