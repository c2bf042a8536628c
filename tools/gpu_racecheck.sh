#!/bin/bash
# Next-round starting point for the partial-grid race: locate the kernel, then let compute-sanitizer
# look at its shared-memory and barrier traffic.  Each step writes under gpurun_out/.
#   1. per-kernel repeatability on half of the SMs (which encoder GEMM stops being bit-identical to its full-grid launch?)
#   2. racecheck / synccheck of that test restricted to the P32 GEMM kernel (slow: the sanitizer serialises; 2 launches are enough)
#   3. the whole-forward failure-rate matrix (tools/diag_race.py) for before / after comparisons
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_z_gpu_stress.py -m gpu -q -k per_kernel -W always 2>&1 | tail -30 | tee gpurun_out/race_locator.txt
for tool in racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --kernel-regex kns=gemm_p32 --print-limit 40 \
    python -m pytest tests/test_z_gpu_stress.py -m gpu -q -x -k "per_kernel and fc2" 2>&1 | tail -60 > gpurun_out/race_$tool.txt
  tail -5 gpurun_out/race_$tool.txt
done
N_EAGER=40 N_GRAPH=40 N_MULTI=8 timeout 300 python tools/diag_race.py 2>&1 | tail -25
