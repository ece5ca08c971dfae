#!/bin/bash
# round 2: full -m gpu suite (incl. the benched-width / EOS / call-site tests) + one bench line with the parity block
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --durations=15 > gpurun_out/r2_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -25 gpurun_out/r2_pytest_gpu.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
echo "bench exit $?"; tail -3 gpurun_out/r2_bench_a.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_a.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'])
print('parity',d.get('parity')); print('extra',d.get('extra')); print('cpu',d.get('cpu_baseline'))
PY
