#!/bin/bash
# round-2 evidence run: full -m gpu suite, smoke(), bench lines (clips N=1 with parity / comparator / ragged, long-form, reference arm)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu --durations=10 > gpurun_out/r2_pytest_full.log 2>&1
echo "pytest exit $?"; tail -16 gpurun_out/r2_pytest_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_full_n1.json 2> gpurun_out/r2_bench_full_n1.err
echo "bench exit $?"; tail -2 gpurun_out/r2_bench_full_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_full_n1.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'], d['roofline']['avg_launch_us'])
print('stages', {k:(round(v['ms'],1),round(v['frac'],3)) for k,v in d['stages'].items()})
print('parity',d.get('parity')); print('extra',d.get('extra')); print('ragged',d.get('ragged')); print('cpu',d.get('cpu_baseline')); print('clocks', d.get('clocks'))
PY
timeout 900 python bench.py --workload longform --steps 2 --warmup 1 > gpurun_out/r2_bench_longform.json 2> gpurun_out/r2_bench_longform.err
echo "longform exit $?"; tail -2 gpurun_out/r2_bench_longform.err; cut -c1-900 gpurun_out/r2_bench_longform.json
( time timeout 1500 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
echo "reference exit $?"; tail -4 gpurun_out/r2_bench_ref.err; cut -c1-700 gpurun_out/r2_bench_ref.json
