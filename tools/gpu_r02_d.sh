#!/bin/bash
# round 2, call D: H16 MSDeformAttn value layout, smem-histogram triplet select, pipelined relhead producer
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02d_gpu_tests.log; tail -12 gpurun_out/r02d_gpu_tests.log
EGTR_BENCH_KERNELS=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; tail -c 600 gpurun_out/r02d_bench.json; grep -v Warn gpurun_out/r02d_bench.err | head -24
EGTR_MSDA_VALUE=f32 EGTR_BENCH_KERNELS=1 timeout 600 python bench.py --steps 20 --warmup 5 --cpu-sample 0 --reference-gpu 0 > gpurun_out/r02d_bench_msda_f32.json 2> gpurun_out/r02d_bench_msda_f32.err; head -c 300 gpurun_out/r02d_bench_msda_f32.json; grep "msda\|relhead" gpurun_out/r02d_bench_msda_f32.err
N="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
timeout 600 $N -k regex:msda_kernel -c 1 -o gpurun_out/r02_msda_enc_h16 python tools/profile_forward.py > gpurun_out/r02d_ncu_msda.log 2>&1; tail -1 gpurun_out/r02d_ncu_msda.log
timeout 600 $N -k regex:relhead_kernel -c 1 -o gpurun_out/r02_relhead_v2 python tools/profile_forward.py > gpurun_out/r02d_ncu_relhead.log 2>&1; tail -1 gpurun_out/r02d_ncu_relhead.log
