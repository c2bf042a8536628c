#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python tools/gemm_bench.py 2>&1 | tee gpurun_out/gemm_bench.txt
EGTR_GEMM_BLOCK_N=128 timeout 600 python tools/gemm_bench.py 2>&1 | tee gpurun_out/gemm_bench_bn128.txt
EGTR_GEMM_BLOCK_N=64 timeout 600 python tools/gemm_bench.py --only enc 2>&1 | tee gpurun_out/gemm_bench_bn64.txt
for i in 0 1 7; do
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_sbf16 -o gpurun_out/gemm_shape$i -f python tools/gemm_bench.py --profile $i > gpurun_out/gprof$i.log 2>&1; tail -1 gpurun_out/gprof$i.log
done
