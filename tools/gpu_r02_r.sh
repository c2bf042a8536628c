#!/bin/bash
# round 2, call R: grid divisor x forwards in flight (throughput configuration), workload B
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
: > gpurun_out/r02r_grid_div.txt
for cfg in "2 8" "3 8" "4 8" "3 12" "4 12" "4 16" "2 12"; do
  set -- $cfg
  echo "== grid_div $1, $2 in flight" >> gpurun_out/r02r_grid_div.txt
  EGTR_THROUGHPUT_GRID_DIV=$1 STAGE_COST_INFLIGHT=$2 STAGE_COST_STEPS=192 timeout 200 python tools/stage_cost.py p32_to_rows 2>&1 | grep "nothing" >> gpurun_out/r02r_grid_div.txt
done
cat gpurun_out/r02r_grid_div.txt
