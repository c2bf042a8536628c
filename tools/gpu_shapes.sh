#!/bin/bash
mkdir -p gpurun_out
EGTR_BENCH_SHAPES=1 timeout 600 python bench.py --cpu-sample 0 --steps 10 2>gpurun_out/shapes.txt | cut -c1-200
grep gemm_p32 gpurun_out/shapes.txt
