#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/probe_split.py "TWB200_SPLIT=1,TWB200_TRACE=100" "TWB200_SPLIT=2,TWB200_LITE=1,TWB200_TRACE=100" "TWB200_SPLIT=2,TWB200_LITE=0,TWB200_TRACE=100" "TWB200_SPLIT=2,TWB200_LITE=0,TWB200_TRACE=-1" "TWB200_SPLIT=1,TWB200_LITE=1,TWB200_TRACE=-1" > gpurun_out/probe_split2.log 2>&1
echo "probe exit $?"; grep setting gpurun_out/probe_split2.log
