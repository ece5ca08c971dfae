#!/bin/bash
# encoder chain with programmatic dependent launch + 4-chain row max in the attention kernel: tests, then A/B of the encoder stage
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_widths.py -x -q -m gpu -k "attention or teacher or encoder or benched_width or transcribe_host" > gpurun_out/r2_encpdl_tests.log 2>&1
rc=$?; echo "tests exit $rc"; tail -3 gpurun_out/r2_encpdl_tests.log
if [ $rc -ne 0 ]; then grep -n "Error\|assert \|^E " gpurun_out/r2_encpdl_tests.log | head; exit 1; fi
timeout 120 python tools/microbench.py encoder_attention 2>&1 | grep tcgen05
for p in 0 1; do TWB200_ENC_PDL=$p timeout 300 python tools/probe_encoder.py 2>&1 | tail -1 | sed "s/^/ENC_PDL=$p /"; done
for p in 0 1; do TWB200_ENC_PDL=$p timeout 600 python bench.py --steps 3 --warmup 3 --no-parity --no-hf-cuda --no-cpu-baseline --no-ragged 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ENC_PDL=$p value',round(d['value'],1),'stages',{k:(round(v['ms'],1),round(v['frac'],3)) for k,v in d['stages'].items()})"; done
