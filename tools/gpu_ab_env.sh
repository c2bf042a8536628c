#!/bin/bash
# A/B of dev knobs through the real forward (CUDA-graph replay): usage gpu_ab_env.sh "VAR=val ..." "VAR=val ..." ...
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for cfg in "$@"; do
echo "=== $cfg"; env $cfg timeout 600 python bench.py --cpu-sample 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],2), round(d['e2e']['value'],2), d['ms_per_step'], d['stage_ms'])"
done
