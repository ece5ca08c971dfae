#!/bin/bash
mkdir -p gpurun_out
(timeout 120 python tools/probe_coresident.py; PROBE_PDL=1 timeout 120 python tools/probe_coresident.py; PROBE_BK=32 timeout 120 python tools/probe_coresident.py;  PROBE_BK=64 PROBE_LITE=0 timeout 120 python tools/probe_coresident.py) > gpurun_out/probe_cores2.log 2>&1
echo "cores exit $?"; cat gpurun_out/probe_cores2.log | grep -v Warn
timeout 600 python tools/probe_split.py "TWB200_SPLIT=2,TWB200_LITE=1,TWB200_TRACE=100,TWB200_PDL=0" "TWB200_SPLIT=2,TWB200_LITE=1,TWB200_TRACE=-1,TWB200_PDL=0" "TWB200_SPLIT=1,TWB200_LITE=1,TWB200_TRACE=-1,TWB200_PDL=0" > gpurun_out/probe_split5.log 2>&1
echo "probe exit $?"; grep setting gpurun_out/probe_split5.log
