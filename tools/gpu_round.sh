#!/bin/bash
# GPU pass: bisect the throughput-mode deviation at full size B, then bench the candidate fallback configurations
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 150 python tools/diag_throughput.py golden 2>&1 | tail -60
ab() { echo "=== $1"; shift; env "$@" timeout 60 python bench.py --cpu-sample 0 --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],2), round(d['e2e']['value'],2), 'latency ms', round(d['config']['single_forward_latency_ms'],3))"; }
{ ab unbalanced EGTR_GEMM_BALANCE=0; ab latency-knobs EGTR_THROUGHPUT_SPLITK=64 EGTR_THROUGHPUT_GRID_DIV=1; ab splitk64-div2-bal EGTR_THROUGHPUT_SPLITK=64; } 2>&1 | tee gpurun_out/fallback_ab.txt
