#!/bin/bash
mkdir -p gpurun_out
(echo "== combine 640-thread CTAs"; timeout 300 python tools/probe_split.py "TWB200_SPLIT=1,TWB200_COMBINE_HPC=10" 2>&1 | grep setting
 echo "== combine 256-thread CTAs (default)"; timeout 300 python tools/probe_split.py "TWB200_SPLIT=1" 2>&1 | grep setting) > gpurun_out/ab7.log 2>&1
cat gpurun_out/ab7.log
