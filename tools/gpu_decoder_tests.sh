#!/bin/bash
# each decoder test in its own process: a trap in one must not poison the CUDA context of the others
out=gpurun_out/${1:-decoder_tests}.log
: > $out
ids=$(python -m pytest tests/test_gpu_decoder.py -m gpu --collect-only -q 2>/dev/null | grep "::")
for t in $ids; do
  echo "=== $t" >> $out
  timeout 240 python -m pytest "$t" -m gpu -q -s -x 2>&1 | grep -E "^(small|A |B |external|in-kernel|\{)|AssertionError|fault|passed|failed" | cut -c1-1200 >> $out
done
cat $out
