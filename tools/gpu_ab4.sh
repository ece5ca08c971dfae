#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bit_identical or teacher_forced or golden or transcribe_host or split or self_attention or first_pass" > gpurun_out/paged_tests.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/paged_tests.log
(echo "== previous build"; timeout 300 python tools/probes/ab_old/tools/probe_split.py 1 2>&1 | grep setting
 echo "== paged cache"; timeout 300 python tools/probe_split.py 1 2>&1 | grep setting) > gpurun_out/ab4.log 2>&1
cat gpurun_out/ab4.log
