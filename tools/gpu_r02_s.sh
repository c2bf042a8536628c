#!/bin/bash
# round 2, call S: quarter grids + 16 images in flight: bench B / C / E (no CPU / reference legs)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for wl in B C E; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --cpu-sample 0 --reference-gpu 0 > gpurun_out/r02s_bench_$wl.json 2> gpurun_out/r02s_bench_$wl.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r02s_bench_$wl.json'))
print('$wl value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms',round(d['ms_per_step'],3),'check',d['output_check']['deviating'],d['output_check']['worst_rel_err'],'gemm frac',round(d['roofline']['frac'],3))
PY
  grep -v Warn gpurun_out/r02s_bench_$wl.err | grep -i "error\|Traceback" | head -3
done
