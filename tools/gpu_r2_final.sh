#!/bin/bash
# round-2 final evidence: full -m gpu suite, smoke(), the default bench line (N=1), long-form line, reference arm
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/final_pytest.log 2>&1
echo "pytest exit $?"; tail -14 gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 1500 python bench.py --steps 3 --warmup 3 ) > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
echo "bench exit $?"; tail -5 gpurun_out/final_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/final_bench_n1.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'], d['roofline']['avg_launch_us'],'launches',d['gpu_launches'])
print('sequential',{k:(v if k!='e2e' else v['value']) for k,v in d['sequential'].items() if k!='clocks'})
print('stages', {k:(round(v['ms'],1),round(v['frac'],3)) for k,v in d['stages'].items()})
print('parity',d.get('parity')); print('extra',d.get('extra')); print('ragged',d.get('ragged')); print('cpu',d.get('cpu_baseline')); print('clocks', d.get('clocks'))
PY
timeout 900 python bench.py --workload longform --steps 2 --warmup 1 > gpurun_out/final_bench_longform.json 2> gpurun_out/final_bench_longform.err
echo "longform exit $?"; tail -2 gpurun_out/final_bench_longform.err; cut -c1-600 gpurun_out/final_bench_longform.json
