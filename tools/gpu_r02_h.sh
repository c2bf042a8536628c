#!/bin/bash
# round 2, call H (8 GPUs of one box): image-parallel scaling of workload B and the BASELINE.json multi-GPU configs C (N=8), D (N=4), E (N=8)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() {  # workload, gpus, tag
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29500 + $2)) bench.py --gpus $2 --workload $1 --steps 20 --warmup 5 --cpu-sample 0 \
    > gpurun_out/r02h_bench_$1_n$2.json 2> gpurun_out/r02h_bench_$1_n$2.err
  tail -c 300 gpurun_out/r02h_bench_$1_n$2.err | grep -i "error\|Traceback" ; head -c 260 gpurun_out/r02h_bench_$1_n$2.json; echo
}
run B 8
run C 8
run E 8
run D 4
run B 4
run B 2
timeout 300 python bench.py --steps 20 --warmup 5 --cpu-sample 0 --reference-gpu 0 > gpurun_out/r02h_bench_B_n1.json 2> gpurun_out/r02h_bench_B_n1.err; head -c 260 gpurun_out/r02h_bench_B_n1.json; echo
