#!/bin/bash
for b in 2 4; do
echo "=== batch $b"; timeout 600 python bench.py --cpu-sample 0 --batch-per-gpu $b --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],2), round(d['e2e']['value'],2), d['ms_per_step'], d['stage_ms'], d['config'].get('single_forward_latency_ms'))"
done
