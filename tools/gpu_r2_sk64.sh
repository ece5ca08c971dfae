#!/bin/bash
# skinny GEMM as M = 64 instructions (+128-wide tiles for the vocab head): kernel tests, full suite, bench A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "gemm" 2>&1 | tail -3
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/sk64_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/sk64_pytest.log
for ab in 0 1; do
  TWB200_ABSORB=$ab timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-hf-cuda --no-ragged > gpurun_out/sk64_bench_ab$ab.json 2> gpurun_out/sk64_bench_ab$ab.err
  echo "bench ab=$ab exit $?"; tail -2 gpurun_out/sk64_bench_ab$ab.err
  python - <<PY
import json
d=json.load(open('gpurun_out/sk64_bench_ab$ab.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'], d['roofline']['avg_launch_us'])
print('stages', {k:(round(v['ms'],1),round(v['frac'],3)) for k,v in d['stages'].items()})
PY
done
TWB200_TRACE=100 TWB200_GRAPH=0 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-hf-cuda --no-ragged --no-e2e > /dev/null 2> gpurun_out/sk64_trace.err
grep "twb200 trace" gpurun_out/sk64_trace.err | head -13
