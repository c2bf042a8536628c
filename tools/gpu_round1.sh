#!/bin/bash
# First GPU bring-up: everything except tcgen05 first, then the tensor-core path in separate processes.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
export PYTHONUNBUFFERED=1
echo "=== kernels (no tc)"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "not tc" 2>&1 | tail -25 | tee gpurun_out/k_simt.log
echo "=== forward simt"; EGTR_B200_GEMM=simt timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q -s 2>&1 | tail -40 | tee gpurun_out/f_simt.log
echo "=== gemm tc plain"; timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tc and plain" 2>&1 | tail -40 | tee gpurun_out/k_tc_plain.log
echo "=== gemm tc conv"; timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tc and (conv or stem)" 2>&1 | tail -30 | tee gpurun_out/k_tc_conv.log
echo "=== forward tc"; timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q -s 2>&1 | tail -40 | tee gpurun_out/f_tc.log
