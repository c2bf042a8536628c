#!/bin/bash
# round 2, call G: single-GPU bench lines of BASELINE.json configs C / D / E at their per-GPU batch; reference arm; traffic pass
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for wl in C D E; do
  EGTR_BENCH_KERNELS=1 timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 --cpu-sample 1 --reference-gpu 2 > gpurun_out/r02g_bench_$wl.json 2> gpurun_out/r02g_bench_$wl.err
  head -c 330 gpurun_out/r02g_bench_$wl.json; echo; grep -v Warn gpurun_out/r02g_bench_$wl.err | head -8
done
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02g_traffic.csv python tools/profile_forward.py > gpurun_out/r02g_traffic.log 2>&1; tail -1 gpurun_out/r02g_traffic.log
python tools/traffic_summary.py gpurun_out/r02g_traffic.csv gpurun_out/r02_traffic.json | head -12
python tools/launch_summary.py gpurun_out/r02g_traffic.csv 12 > gpurun_out/r02g_launch_summary.txt 2>&1; head -30 gpurun_out/r02g_launch_summary.txt
