#!/bin/bash
# elect-style (warp-uniform) MMA / TMA issue: correctness of every tcgen05 kernel, then kernel and bench timings
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "gemm or attention or teacher or encoder or tokens" > gpurun_out/r2_elect_tests.log 2>&1
rc=$?; echo "kernel tests exit $rc"; tail -4 gpurun_out/r2_elect_tests.log
if [ $rc -ne 0 ]; then grep -n "Error\|assert " gpurun_out/r2_elect_tests.log | head -20; exit 1; fi
for v in 1 2; do TWB200_FA_VARIANT=$v timeout 120 python tools/microbench.py encoder_attention 2>&1 | grep tcgen05 | sed "s/^/FA$v /"; done
timeout 200 python tools/microbench.py gemm 2>&1 | tail -12
for v in 1 2; do TWB200_FA_VARIANT=$v timeout 300 python tools/probe_encoder.py 2>&1 | tail -1; done
timeout 900 python bench.py --steps 2 --warmup 3 --no-parity --no-hf-cuda --no-cpu-baseline --no-ragged > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err
echo "bench exit $?"; tail -2 gpurun_out/r2_bench_c.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_c.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'], d['roofline']['avg_launch_us'])
print('stages', {k:(round(v['ms'],1),round(v['frac'],3)) for k,v in d['stages'].items()})
PY
