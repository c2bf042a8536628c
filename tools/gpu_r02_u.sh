#!/bin/bash
# round 2, call U: final single-GPU bench lines of BASELINE.json configs C / D / E (all legs)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for wl in C D E; do
  EGTR_BENCH_KERNELS=1 timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 --cpu-sample 1 --reference-gpu 2 > gpurun_out/r02u_bench_$wl.json 2> gpurun_out/r02u_bench_$wl.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r02u_bench_$wl.json'))
print('$wl value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms',round(d['ms_per_step'],3),'lat',round(d['config']['single_forward_latency_ms'],2),'check',d['output_check']['deviating'],'frac',round(d['roofline']['frac'],3),'refgpu',round(d['reference_gpu']['as_shipped']['images_per_s'],1),'cpu',round(d['cpu_baseline']['value'],2))
PY
done
