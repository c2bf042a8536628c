#!/bin/bash
# image-parallel scaling of workload B: N = 1 and N = $1 on the same box (gpurun --gpus N)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
NG=${1:-2}
for n in 1 $NG; do
if [ $n = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n))"; fi
timeout 600 $L bench.py --gpus $n --workload B --steps 20 --warmup 5 --cpu-sample 0 --reference-gpu 0 > gpurun_out/r02_scaling_bench_B_n$n.json 2> gpurun_out/r02_scaling_bench_B_n$n.err
python - <<PY
import json
lines=[l for l in open('gpurun_out/r02_scaling_bench_B_n$n.json').read().splitlines() if l.startswith('{')]
d=json.loads(lines[-1]); print('B n$n value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms',round(d['ms_per_step'],3),'check',d['output_check']['deviating'])
PY
done
