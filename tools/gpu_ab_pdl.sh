#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for m in 0 1 2 0 2; do
echo "=== bench PDL mode $m"; EGTR_B200_PDL=$m timeout 600 python bench.py --cpu-sample 0 2>gpurun_out/bench_pdl$m.err | tee gpurun_out/bench_pdl$m.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['stage_ms'])"
done
