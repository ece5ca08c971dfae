#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_final.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu_final.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_final3_n1.json 2> gpurun_out/bench_final3_n1.err
echo "bench exit $?"; cut -c1-400 gpurun_out/bench_final3_n1.json
python -c "
import json; d=json.load(open('gpurun_out/bench_final3_n1.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'],d['roofline']['avg_launch_us'],'stages',{k:(round(v['ms'],1),round(v['frac'],3)) for k,v in d['stages'].items()}, 'cpu',d.get('cpu_baseline',{}).get('value'))"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
