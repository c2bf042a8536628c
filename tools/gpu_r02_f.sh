#!/bin/bash
# round 2, call F: full GPU suite after the H16 / BIG-relhead / tie-aware comparison changes; bench B; smoke
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q -s --tb=short 2>&1 | grep -v "^$" | cut -c1-330 > gpurun_out/r02f_gpu_tests.log; grep -n "^E \|FAILED\|passed\|failed\|vs oracle\|batched\|note:\|P=200\|P=65\|P=128\|P=256" gpurun_out/r02f_gpu_tests.log | head -70
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
EGTR_BENCH_KERNELS=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; head -c 400 gpurun_out/r02f_bench.json; grep -v Warn gpurun_out/r02f_bench.err | head -12
