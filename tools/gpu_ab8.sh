#!/bin/bash
mkdir -p gpurun_out
(echo "== fc1 with 64-wide tiles (80 CTAs, default)"; timeout 300 python tools/probe_split.py "TWB200_SPLIT=1,TWB200_SK_WAVES=1" 2>&1 | grep setting
 echo "== fc1 with 32-wide tiles (160 tiles on 148 CTAs)"; timeout 300 python tools/probe_split.py "TWB200_SPLIT=1,TWB200_SK_WAVES=2" 2>&1 | grep setting) > gpurun_out/ab8.log 2>&1
cat gpurun_out/ab8.log
