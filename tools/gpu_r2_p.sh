#!/bin/bash
mkdir -p gpurun_out
echo "== net tests"; timeout 1200 python -m pytest tests/test_net_gpu.py tests/test_reference_gpu.py tests/test_conv28_gpu.py tests/test_reference_adapter.py -x -q -m gpu 2>&1 | tail -5
for F in 1 0; do
  echo "== DLB_HEAD_FOLD=$F"; CHILD=1 DLB_HEAD_FOLD=$F timeout 300 python tools/conv28_bench.py 128 256 512 1024 8192 2>&1 | tail -6 | cut -c1-60
done
echo "== peaks"; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/peaks tools/peaks.cu 2>/dev/null && /tmp/peaks | tee gpurun_out/peaks_r2.json
