#!/bin/bash
for v in "" m5u4 m4u8 m3u16 m8u4; do
  if [ -n "$v" ]; then export EGTR_B200_LIB=$PWD/egtr_b200/csrc/libvar_$v.so; else unset EGTR_B200_LIB; fi
  echo "=== ${v:-default m6u4}"; timeout 600 python bench.py --cpu-sample 0 --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],2), round(d['e2e']['value'],2), 'msda_enc us', round(d['roofline_msda_enc']['avg_launch_us'],1))"
done
