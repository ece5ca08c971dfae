#!/bin/bash
# absorbed cross-attention in the model: full -m gpu suite, then the bench A/B (TWB200_ABSORB=1 / 0) and the step timeline
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --durations=5 > gpurun_out/ab2_pytest.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/ab2_pytest.log
for ab in 1 0; do
  TWB200_ABSORB=$ab timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-hf-cuda --no-ragged > gpurun_out/ab2_bench_ab$ab.json 2> gpurun_out/ab2_bench_ab$ab.err
  echo "bench ab=$ab exit $?"; tail -2 gpurun_out/ab2_bench_ab$ab.err
  python - <<PY
import json
d=json.load(open('gpurun_out/ab2_bench_ab$ab.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'], d['roofline']['avg_launch_us'])
print('stages', {k:(round(v['ms'],1),round(v['frac'],3)) for k,v in d['stages'].items()})
PY
done
TWB200_TRACE=100 TWB200_GRAPH=0 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-hf-cuda --no-ragged --no-e2e > /dev/null 2> gpurun_out/ab2_trace.err
grep "twb200 trace" gpurun_out/ab2_trace.err | head -40
