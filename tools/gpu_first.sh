#!/bin/bash
# first bring-up on the GPU box: environment facts + parity tests (no -x: collect every failure)
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
nproc; free -g | head -2
python -c "import torch, transformers; print(torch.__version__, transformers.__version__, torch.cuda.is_available())"
} > gpurun_out/env.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -rA --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -60 gpurun_out/pytest_gpu.log
