#!/bin/bash
mkdir -p gpurun_out
for bn in 0 128 64; do for fl in "" "--noflush"; do
echo "=== BN=$bn $fl"; EGTR_GEMM_BLOCK_N=$bn timeout 300 python tools/gemm_bench.py --iters 10 --p32 --p32out $fl 2>&1 | grep -v conv | grep -v dec_ | awk '{print $1, $2,$3,$4,$5,$6,$7, $8, $9, $10, $11}'
done; done
