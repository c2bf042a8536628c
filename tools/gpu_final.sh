#!/bin/bash
# Round-end refresh: tests, smoke, bench (both arms), launch list, DRAM traffic, ncu full of the top kernels.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
bash tools/gpu_check.sh
echo "=== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tee gpurun_out/bench_ref.json | cut -c1-300
bash tools/gpu_traffic.sh
N="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
timeout 900 $N -k regex:gemm_p32 -s 56 -c 6 -o gpurun_out/r01_gemm_p32_enc python tools/profile_forward.py > gpurun_out/ncu_a.log 2>&1; tail -1 gpurun_out/ncu_a.log
timeout 900 $N -k regex:msda_kernel -c 1 -o gpurun_out/r01_msda_enc python tools/profile_forward.py > gpurun_out/ncu_b.log 2>&1; tail -1 gpurun_out/ncu_b.log
timeout 900 $N -k regex:msda_kernel -s 6 -c 1 -o gpurun_out/r01_msda_dec python tools/profile_forward.py > gpurun_out/ncu_c.log 2>&1; tail -1 gpurun_out/ncu_c.log
timeout 900 $N -k regex:gemm_sbf16 -c 3 -o gpurun_out/r01_gemm_sbf16_rel python tools/profile_forward.py > gpurun_out/ncu_d.log 2>&1; tail -1 gpurun_out/ncu_d.log
