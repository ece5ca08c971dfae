#!/bin/bash
# bring-up of the absorbed cross-attention kernel: grouped GEMM + kernel tests, both MN-major descriptor field orders
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/ab_smi.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "gemm_grouped" > gpurun_out/ab_grouped.log 2>&1
echo "grouped rc=$?" >> gpurun_out/ab_grouped.log
for desc in 0 1; do
  TWB200_AB_DESC=$desc timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "absorbed_attention" > gpurun_out/ab_kernel_desc$desc.log 2>&1
  echo "desc=$desc rc=$?" >> gpurun_out/ab_kernel_desc$desc.log
done
tail -n 5 gpurun_out/ab_grouped.log gpurun_out/ab_kernel_desc0.log gpurun_out/ab_kernel_desc1.log
