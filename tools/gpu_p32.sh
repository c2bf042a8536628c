#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== p32 tests"; timeout 600 python -m pytest tests/test_gpu_p32.py -m gpu -q -x 2>&1 | tail -${TAIL:-12} | tee gpurun_out/p32_tests.log
echo "=== p32 tests 1cta"; EGTR_GEMM_CTAS=1 timeout 600 python -m pytest tests/test_gpu_p32.py -m gpu -q -x -k layernorm 2>&1 | tail -${TAIL:-6}
echo "=== forward"; timeout 600 python -m pytest tests/test_gpu_forward.py -m gpu -q -x 2>&1 | tail -6
bash tools/gpu_ab_env.sh "X=1" "EGTR_PIPE_CONCURRENCY=1"
