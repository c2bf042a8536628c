#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_msda_h16.py tests/test_gpu_serving.py "tests/test_gpu_forward.py::test_forward_full_size_configs" "tests/test_gpu_forward.py::test_fused_relation_stage_matches_unfused_kernels" "tests/test_gpu_forward.py::test_forward_matches_reference_golden" tests/test_gpu_relhead.py -m gpu -q -s --tb=short 2>&1 | grep -v "^$" | cut -c1-400 > gpurun_out/r02e_tests.log; grep -n "Error\|error\|FAILED\|passed\|failed\|vs oracle\|batched" gpurun_out/r02e_tests.log | head -80
