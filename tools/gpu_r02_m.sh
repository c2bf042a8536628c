#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
bash tools/gpu_decoder_tests.sh r02m_decoder_tests > /dev/null 2>&1; grep -E "passed|failed|fault|Error" gpurun_out/r02m_decoder_tests.log | head -20
timeout 300 python tools/decoder_profile.py > gpurun_out/r02m_decoder_profile.txt 2>&1; tail -16 gpurun_out/r02m_decoder_profile.txt
