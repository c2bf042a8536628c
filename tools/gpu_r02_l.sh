#!/bin/bash
# round 2, call L: fused decoder — throughput A/B (stage_cost), launch list of a lone forward, full GPU test suite
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/stage_cost.txt
EGTR_DECODER=layers timeout 300 python tools/stage_cost.py dec_small > gpurun_out/r02l_stage_layers.log 2>&1; tail -3 gpurun_out/r02l_stage_layers.log
EGTR_DECODER=fused timeout 300 python tools/stage_cost.py dec_small msda_enc > gpurun_out/r02l_stage_fused.log 2>&1; tail -4 gpurun_out/r02l_stage_fused.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02l_traffic.csv python tools/profile_forward.py > gpurun_out/r02l_traffic.log 2>&1; tail -1 gpurun_out/r02l_traffic.log
python tools/launch_summary.py gpurun_out/r02l_traffic.csv 12 > gpurun_out/r02l_launch_summary.txt 2>&1; head -32 gpurun_out/r02l_launch_summary.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02l_gpu_tests.log; tail -8 gpurun_out/r02l_gpu_tests.log
