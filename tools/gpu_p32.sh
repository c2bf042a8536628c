#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== p32 tests (OCC2)"; EGTR_GEMM_OCC2=1 timeout 600 python -m pytest tests/test_gpu_p32.py -m gpu -q -x 2>&1 | tail -${TAIL:-8} | tee gpurun_out/p32_tests_occ2.log
bash tools/gpu_ab_env.sh "EGTR_GEMM_OCC2=0" "EGTR_GEMM_OCC2=1" "EGTR_GEMM_OCC2=1 EGTR_B200_PDL=1" "EGTR_GEMM_OCC2=1 EGTR_PIPE_CONCURRENCY=1" "EGTR_GEMM_OCC2=0 EGTR_PIPE_CONCURRENCY=1"
