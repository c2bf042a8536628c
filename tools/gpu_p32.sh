#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== p32 tests (default)"; timeout 600 python -m pytest tests/test_gpu_p32.py -m gpu -q -x 2>&1 | tail -${TAIL:-15} | tee gpurun_out/p32_tests.log
echo "=== p32 tests (1 CTA)"; EGTR_GEMM_CTAS=1 timeout 600 python -m pytest tests/test_gpu_p32.py -m gpu -q -x 2>&1 | tail -${TAIL:-15} | tee gpurun_out/p32_tests_1cta.log
bash tools/gpu_ab_env.sh "EGTR_PIPE_CONCURRENCY=1" "EGTR_PIPE_CONCURRENCY=3"
