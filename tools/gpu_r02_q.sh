#!/bin/bash
# round 2, call Q: full GPU suite with the fused decoder everywhere; smoke; bench B
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02q_gpu_tests.log; tail -4 gpurun_out/r02q_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
EGTR_BENCH_KERNELS=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err; head -c 500 gpurun_out/r02q_bench.json; echo; grep -v Warn gpurun_out/r02q_bench.err | head -12
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02q_bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'lat',d['config'].get('single_forward_latency_ms'),'check',d['output_check']['deviating'],d['output_check']['worst_rel_err'])
PY
