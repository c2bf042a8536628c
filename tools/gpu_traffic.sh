#!/bin/bash
# DRAM traffic per launch of one forward (workload B, batch 1): feeds the `traffic` field of bench.py's roofline objects.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/traffic.csv python tools/profile_forward.py > gpurun_out/traffic.log 2>&1; tail -1 gpurun_out/traffic.log
python tools/traffic_summary.py gpurun_out/traffic.csv gpurun_out/r01_traffic.json
