#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
bash tools/gpu_decoder_tests.sh r02m_decoder_tests > /dev/null 2>&1; grep -E "^===|passed|failed|fault|Error" gpurun_out/r02m_decoder_tests.log | cut -c1-200 | head -40
for cl in 8 16; do
EGTR_DECODER=fused EGTR_DECODER_CLUSTER=$cl timeout 300 python tools/decoder_profile.py > gpurun_out/r02m_decoder_profile_c$cl.txt 2>&1; tail -14 gpurun_out/r02m_decoder_profile_c$cl.txt
done
