#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ncu_encoder_small.py 16 > gpurun_out/r2_enc_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:encoder_attention_tc3 -s 1 -c 1 -o gpurun_out/r2_fa3 python tools/ncu_encoder_small.py 16 > gpurun_out/r2_fa3_ncu.log 2>&1
echo "fa3 ncu exit $?"
timeout 900 python bench.py --steps 2 --warmup 3 --no-parity --no-hf-cuda --no-cpu-baseline --no-ragged > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err
echo "bench exit $?"; tail -2 gpurun_out/r2_bench_d.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_d.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'], d['roofline']['avg_launch_us'])
print('stages', {k:(round(v['ms'],1),round(v['frac'],3)) for k,v in d['stages'].items()})
PY
