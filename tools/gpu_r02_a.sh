#!/bin/bash
# round 2, call A: does the round-1 race fix hold? GPU suite, failure-rate matrix, per-kernel locator, bench full vs half grids.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests -m gpu -q -W always 2>&1 | tail -40 > gpurun_out/r02a_gpu_tests.log; tail -5 gpurun_out/r02a_gpu_tests.log
N_EAGER=60 N_GRAPH=60 N_MULTI=12 timeout 500 python tools/diag_race.py 2>&1 | tail -25; cp gpurun_out/race.txt gpurun_out/r02a_race_matrix.txt
timeout 200 python bench.py --steps 40 --warmup 5 > gpurun_out/r02a_bench_div1.json 2> gpurun_out/r02a_bench_div1.err; tail -c 600 gpurun_out/r02a_bench_div1.json
EGTR_THROUGHPUT_GRID_DIV=2 EGTR_PIPE_CONCURRENCY=8 timeout 200 python bench.py --steps 40 --warmup 5 > gpurun_out/r02a_bench_div2.json 2> gpurun_out/r02a_bench_div2.err; tail -c 600 gpurun_out/r02a_bench_div2.json
for tool in racecheck synccheck; do
  timeout 400 compute-sanitizer --tool $tool --kernel-regex kns=gemm_p32 --print-limit 40 \
    python -m pytest tests/test_z_gpu_stress.py -m gpu -q -x -k "per_kernel and fc2" 2>&1 | tail -60 > gpurun_out/r02a_race_$tool.txt
  tail -5 gpurun_out/r02a_race_$tool.txt
done
