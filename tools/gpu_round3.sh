#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== gpu tests"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/tests.log
echo "=== bench"; timeout 900 python bench.py 2>gpurun_out/bench.err | tee gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "=== bench reference"; timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tee gpurun_out/bench_ref.json
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/profile_forward.py > gpurun_out/prof1.log 2>&1; tail -2 gpurun_out/prof1.log
