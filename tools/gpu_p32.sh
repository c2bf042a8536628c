#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== p32 tests"; timeout 600 python -m pytest tests/test_gpu_p32.py -m gpu -q -x 2>&1 | tail -30 | tee gpurun_out/p32_tests.log
echo "=== gemm bench fp32 rows"; timeout 300 python tools/gemm_bench.py --iters 10 2>&1 | tee gpurun_out/gemm_bench_f32.txt
echo "=== gemm bench p32 rows"; timeout 300 python tools/gemm_bench.py --iters 10 --p32 2>&1 | tee gpurun_out/gemm_bench_p32.txt
echo "=== gemm bench p32 rows, p32 out"; timeout 300 python tools/gemm_bench.py --iters 10 --p32 --p32out 2>&1 | tee gpurun_out/gemm_bench_p32out.txt
