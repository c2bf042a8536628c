#!/bin/bash
# Round-end refresh (short form): tests, smoke, bench (both arms), launch list, ncu full of the MSDA encoder launch.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
bash tools/gpu_check.sh
echo "=== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tee gpurun_out/bench_ref.json | cut -c1-300
N="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
timeout 300 $N -k regex:msda_kernel -c 1 -o gpurun_out/r01_msda_enc python tools/profile_forward.py > gpurun_out/ncu_b.log 2>&1; tail -1 gpurun_out/ncu_b.log
