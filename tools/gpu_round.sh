#!/bin/bash
# quick GPU pass: new full-size parity tests + MSDA kernel tests, then the MSDA A/B builds
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -m gpu -q -x -s -k "full_size or msda" 2>&1 | tail -40 | tee gpurun_out/fullsize_tests.log
bash tools/gpu_msda_ab.sh 2>&1 | tee gpurun_out/msda_ab.txt
