#!/bin/bash
# final full GPU suite, smoke, final bench lines B (all legs) and C / D / E on one GPU
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_final_gpu_tests.log; tail -3 gpurun_out/r02_final_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
EGTR_BENCH_KERNELS=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench_B.json 2> gpurun_out/r02_final_bench_B.err
for wl in C D E; do
  timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 --cpu-sample 1 --reference-gpu 2 > gpurun_out/r02_final_bench_$wl.json 2> gpurun_out/r02_final_bench_$wl.err
done
python - <<'PY'
import json
for wl in 'BCDE':
    try:
        d=json.load(open(f'gpurun_out/r02_final_bench_{wl}.json'))
        print(wl,'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms',round(d['ms_per_step'],3),'lat',round(d['config']['single_forward_latency_ms'],2),'check',d['output_check']['deviating'],'frac',round(d['roofline']['frac'],3),'launches',d['gpu_launches'], 'dec', (d.get('roofline_decoder') or {}).get('avg_launch_us'))
    except Exception as e: print(wl,'failed',e)
PY
