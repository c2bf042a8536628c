#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
for v in E A E A; do echo "== variant $v"; EGTR_B200_LIB=$PWD/egtr_b200/csrc/libvar_$v.so timeout 300 python tools/gemm_bench.py --iters 15 2>&1 | awk '{print $1, $8}' | tr '\n' ' '; echo; done
