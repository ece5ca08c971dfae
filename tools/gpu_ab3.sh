#!/bin/bash
mkdir -p gpurun_out
(echo "== old"; timeout 300 python tools/probes/ab_old/tools/probe_encoder.py 2>&1 | grep encoder
 echo "== new (prefetch)"; timeout 300 python tools/probe_encoder.py 2>&1 | grep encoder
 echo "== new, FA variant 0"; TWB200_FA_VARIANT=0 timeout 300 python tools/probe_encoder.py 2>&1 | grep encoder) > gpurun_out/ab3.log 2>&1
cat gpurun_out/ab3.log
