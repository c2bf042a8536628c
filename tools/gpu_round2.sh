#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== gpu tests"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/tests.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
echo "=== bench"; timeout 900 python bench.py 2>gpurun_out/bench.err | tee gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "=== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tee gpurun_out/bench_ref.json
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/profile_forward.py > gpurun_out/prof1.log 2>&1; tail -2 gpurun_out/prof1.log
echo "=== ncu full msda"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:msda_kernel -c 2 -o gpurun_out/msda_r1 -f python tools/profile_forward.py > gpurun_out/prof2.log 2>&1; tail -2 gpurun_out/prof2.log
echo "=== ncu full gemm (encoder fc1-like, skip the first 60 gemm launches)"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_sbf16 -s 60 -c 3 -o gpurun_out/gemm_r1 -f python tools/profile_forward.py > gpurun_out/prof3.log 2>&1; tail -2 gpurun_out/prof3.log
ls -la gpurun_out
