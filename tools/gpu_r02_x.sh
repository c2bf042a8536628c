#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_stem.py tests/test_gpu_forward.py tests/test_gpu_serving.py -m gpu -q -s 2>&1 | grep -E "fp64|passed|failed|Error|error" | cut -c1-300 | head -30
timeout 300 python tools/stem_bench.py 2>&1 | grep -E "gather|tma" | tee gpurun_out/r02x_stem_bench.txt
