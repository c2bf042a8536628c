#!/bin/bash
# ncu --set full captures of the top kernels inside one real forward (workload B, batch 1).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
timeout 900 $N -k regex:gemm_p32 -s 57 -c 7 -o gpurun_out/r01_gemm_p32_enc python tools/profile_forward.py > gpurun_out/ncu_a.log 2>&1; tail -1 gpurun_out/ncu_a.log
timeout 900 $N -k regex:msda_kernel -c 1 -o gpurun_out/r01_msda_enc python tools/profile_forward.py > gpurun_out/ncu_b.log 2>&1; tail -1 gpurun_out/ncu_b.log
timeout 900 $N -k regex:msda_kernel -s 6 -c 1 -o gpurun_out/r01_msda_dec python tools/profile_forward.py > gpurun_out/ncu_c.log 2>&1; tail -1 gpurun_out/ncu_c.log
timeout 900 $N -k regex:gemm_sbf16 -c 5 -o gpurun_out/r01_gemm_sbf16_rel python tools/profile_forward.py > gpurun_out/ncu_d.log 2>&1; tail -1 gpurun_out/ncu_d.log
timeout 900 $N -k regex:gemm_p32 -s 6 -c 4 -o gpurun_out/r01_gemm_p32_layer1 python tools/profile_forward.py > gpurun_out/ncu_e.log 2>&1; tail -1 gpurun_out/ncu_e.log
ls -la gpurun_out/*.ncu-rep
