#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== p32 tests (WS forced)"; EGTR_GEMM_WS=2 timeout 600 python -m pytest tests/test_gpu_p32.py -m gpu -q -x 2>&1 | tail -${TAIL:-8} | tee gpurun_out/p32_tests_ws.log
echo "=== p32 tests (auto)"; timeout 600 python -m pytest tests/test_gpu_p32.py -m gpu -q -x 2>&1 | tail -${TAIL:-8}
for ws in 0 1 2; do echo "--- gemm bench WS=$ws"; EGTR_GEMM_WS=$ws python tools/gemm_bench.py --iters 10 --p32 --p32out --only _ 2>&1 | grep -E "round|enc_value|enc_fc1|dec_value|l1_conv3|l3_conv3"; done
bash tools/gpu_ab_env.sh "EGTR_GEMM_WS=0" "EGTR_GEMM_WS=1" "EGTR_GEMM_WS=2"
