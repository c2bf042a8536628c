#!/bin/bash
# round 2, call O: bench B with the fused decoder for forwards in flight; smoke
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
EGTR_BENCH_KERNELS=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err; head -c 600 gpurun_out/r02o_bench.json; echo; grep -v Warn gpurun_out/r02o_bench.err | head -40
