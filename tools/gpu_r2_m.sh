#!/bin/bash
mkdir -p gpurun_out
echo "== ffma2 probe"; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2_probe tools/ffma2_probe.cu && /tmp/ffma2_probe
echo "== bench"; CHILD=1 DLB_CONV28=1 timeout 300 python tools/conv28_bench.py 256 8192 2>&1 | tail -3
echo "== ncu conv28"; bash tools/gpu_ncu_conv.sh c28b conv28 8192 2 2 > /dev/null 2>&1
python tools/ncu_source_mix.py gpurun_out/c28b_source.csv 12 > gpurun_out/c28b_source_mix.txt 2>&1; grep -v "^   [A-Z@]" gpurun_out/c28b_source_mix.txt | head -80
rm -f gpurun_out/c28b_source.csv
