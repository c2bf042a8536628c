#!/bin/bash
# round 2, call C: full GPU suite with the new tests, the rewritten bench (CUPTI rooflines, u8 -> triplets e2e, reference_gpu),
# launch list and ncu --set full of the fused relation-head kernel
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r02c_gpu_tests.log; tail -6 gpurun_out/r02c_gpu_tests.log
EGTR_BENCH_KERNELS=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; tail -c 2500 gpurun_out/r02c_bench.json; tail -45 gpurun_out/r02c_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02c_launches.csv python tools/profile_forward.py > gpurun_out/r02c_launches.log 2>&1; tail -1 gpurun_out/r02c_launches.log
python tools/launch_summary.py gpurun_out/r02c_launches.csv 30 > gpurun_out/r02c_launch_summary.txt; head -32 gpurun_out/r02c_launch_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -k regex:relhead_kernel -c 1 -o gpurun_out/r02_relhead python tools/profile_forward.py > gpurun_out/r02c_ncu_relhead.log 2>&1; tail -1 gpurun_out/r02c_ncu_relhead.log
