#!/bin/bash
mkdir -p gpurun_out
(echo "== old (7bef087)"; timeout 300 python tools/probes/ab_old/tools/probe_split.py 1 2>&1 | grep setting
 echo "== new"; timeout 300 python tools/probe_split.py 1 2>&1 | grep setting
 echo "== old again"; timeout 300 python tools/probes/ab_old/tools/probe_split.py 1 2>&1 | grep setting
 echo "== new, DA_STAGES=3"; timeout 300 python tools/probe_split.py "TWB200_SPLIT=1,TWB200_DA_STAGES=3" 2>&1 | grep setting) > gpurun_out/ab.log 2>&1
cat gpurun_out/ab.log
