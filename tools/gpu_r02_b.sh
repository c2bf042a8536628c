#!/bin/bash
# round 2, call B: first hardware run of the fused relation-head kernel (relhead.cu)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_relhead.py -m gpu -q -x -s 2>&1 | tail -40 > gpurun_out/r02b_relhead_tests.log; tail -25 gpurun_out/r02b_relhead_tests.log
timeout 600 python -m pytest tests/test_gpu_forward.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02b_forward_tests.log; tail -8 gpurun_out/r02b_forward_tests.log
timeout 300 python bench.py --steps 40 --warmup 5 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; tail -c 1500 gpurun_out/r02b_bench.json; tail -3 gpurun_out/r02b_bench.err
EGTR_RELHEAD=unfused timeout 300 python bench.py --steps 40 --warmup 5 > gpurun_out/r02b_bench_unfused.json 2> gpurun_out/r02b_bench_unfused.err; head -c 300 gpurun_out/r02b_bench_unfused.json
