#!/bin/bash
# round 2, call V (8 GPUs of one box): image-parallel scaling of workload B in the final configuration (16 images in flight per GPU)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() {  # workload, gpus
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29500 + $2)) bench.py --gpus $2 --workload $1 --steps 20 --warmup 5 --cpu-sample 0 --reference-gpu 0 \
    > gpurun_out/r02v_bench_$1_n$2.json 2> gpurun_out/r02v_bench_$1_n$2.err
  tail -c 400 gpurun_out/r02v_bench_$1_n$2.err | grep -i "error\|Traceback" ; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02v_bench_$1_n$2.json'))
    print('$1 n$2 value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms',round(d['ms_per_step'],3),'check',d['output_check']['deviating'])
except Exception as e: print('$1 n$2 failed', e)
PY
}
run B 8
run B 4
run B 2
timeout 300 python bench.py --steps 20 --warmup 5 --cpu-sample 0 --reference-gpu 0 > gpurun_out/r02v_bench_B_n1.json 2> gpurun_out/r02v_bench_B_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r02v_bench_B_n1.json')); print('B n1 value',round(d['value'],1),'e2e',round(d['e2e']['value'],1))"
