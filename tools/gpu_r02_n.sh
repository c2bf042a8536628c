#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for f in 0 128 256; do
timeout 200 python tools/decoder_profile.py 1 $f > gpurun_out/r02n_decoder_profile_$f.txt 2>&1; grep -E "debug|decoder kernel|init|qkv|ln1|oproj|mha|msda|fc" gpurun_out/r02n_decoder_profile_$f.txt
done
