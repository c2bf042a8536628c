#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== p32 tests"; timeout 600 python -m pytest tests/test_gpu_p32.py -m gpu -q 2>&1 | tail -${TAIL:-30} | tee gpurun_out/p32_tests.log
