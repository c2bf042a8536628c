#!/bin/bash
# grid divisor x forwards in flight (throughput configuration), workload B
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
out=gpurun_out/${1:-r02r_grid_div}.txt
: > $out
for cfg in ${CFGS:-"2:8 3:8 4:8 3:12 4:12 4:16 2:12"}; do
  d=${cfg%%:*}; n=${cfg##*:}
  echo "== grid_div $d, $n in flight" >> $out
  EGTR_THROUGHPUT_GRID_DIV=$d STAGE_COST_INFLIGHT=$n STAGE_COST_STEPS=192 timeout 300 python tools/stage_cost.py p32_to_rows 2>&1 | grep "nothing" >> $out
done
cat $out
