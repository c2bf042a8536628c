#!/bin/bash
mkdir -p gpurun_out
for i in 0 1 3; do
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_p32 -o gpurun_out/p32_shape$i -f python tools/gemm_bench.py --p32 --p32out --profile $i > gpurun_out/gprof$i.log 2>&1; tail -1 gpurun_out/gprof$i.log
done
