#!/bin/bash
# launch list + ncu --set full of the fused decoder kernel (16-CTA cluster of a lone forward; 8-CTA cluster),
# GPU tests of the engine / decoder after the side-stream geometry change, bench B with all legs
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_profile_traffic.csv python tools/profile_forward.py > gpurun_out/r02_profile_traffic.log 2>&1; tail -1 gpurun_out/r02_profile_traffic.log
python tools/launch_summary.py gpurun_out/r02_profile_traffic.csv 12 > gpurun_out/r02_profile_launch_summary.txt 2>&1; head -8 gpurun_out/r02_profile_launch_summary.txt
python tools/traffic_summary.py gpurun_out/r02_profile_traffic.csv gpurun_out/r02_profile_traffic.json | head -5
N="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
timeout 600 $N -k regex:decoder_kernel -c 1 -o gpurun_out/r02_decoder_c16 python tools/profile_forward.py > gpurun_out/ncu_dec16.log 2>&1; tail -1 gpurun_out/ncu_dec16.log
EGTR_DECODER_CLUSTER=8 timeout 600 $N -k regex:decoder_kernel -c 1 -o gpurun_out/r02_decoder_c8 python tools/profile_forward.py > gpurun_out/ncu_dec8.log 2>&1; tail -1 gpurun_out/ncu_dec8.log
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_decoder.py tests/test_gpu_serving.py -m gpu -x -q 2>&1 | tail -3
EGTR_BENCH_KERNELS=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_profile_bench_B.json 2> gpurun_out/r02_profile_bench_B.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_profile_bench_B.json'))
print('B value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'lat',d['config'].get('single_forward_latency_ms'),'check',d['output_check']['deviating'],'frac',round(d['roofline']['frac'],3), 'ref_gpu', d.get('reference_gpu',{}).get('as_shipped'))
PY
