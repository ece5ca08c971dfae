#!/bin/bash
# builds a probe copy of the library with phase tracing in the tcgen05 GEMM (not the product build)
set -e
cd "$(dirname "$0")/../.."
mkdir -p tools/probes/trace_build
for f in logmel gemm_simt gemm_tc gemm_skinny elementwise attention attention_tc select model; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
       --expt-relaxed-constexpr -DTW_GEMM_TRACE -c taiwan-whisper_b200/csrc/$f.cu -o tools/probes/trace_build/$f.o &
done
wait
nvcc -shared -o tools/probes/trace_build/libtwb200.so tools/probes/trace_build/*.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC
