#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "gemm" 2>&1 | tail -3
EGTR_B200_LIB=$PWD/egtr_b200/csrc/libvar_P.so timeout 300 python tools/gemm_bench.py --iters 10 --prof 2>&1 | tee gpurun_out/gemm_prof.txt
