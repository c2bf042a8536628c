#!/bin/bash
# quick GPU pass: MSDA encoder patch A/B (8x4 vs 8x8 pixels per CTA), correctness of the variant first
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
EGTR_B200_LIB=$PWD/egtr_b200/csrc/libvar_q64.so timeout 600 python -m pytest tests -m gpu -q -x -k "msda or golden" 2>&1 | tail -5 | tee gpurun_out/q64_tests.log
for v in "" q64 "" q64; do
  if [ -n "$v" ]; then export EGTR_B200_LIB=$PWD/egtr_b200/csrc/libvar_$v.so; else unset EGTR_B200_LIB; fi
  echo "=== ${v:-default}"; timeout 600 python bench.py --cpu-sample 0 --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],2), round(d['e2e']['value'],2), 'msda_enc us', round(d['roofline_msda_enc']['avg_launch_us'],1))"
done 2>&1 | tee gpurun_out/msda_ab2.txt
