#!/bin/bash
mkdir -p gpurun_out
echo "== conv28 tests"; timeout 900 python -m pytest tests/test_conv28_gpu.py -x -q -m gpu 2>&1 | tail -4
echo "== bench"; CHILD=1 DLB_CONV28=1 timeout 300 python tools/conv28_bench.py 128 256 512 1024 2048 8192 2>&1 | tail -7
echo "== old kernels at 64/128"; CHILD=1 DLB_CONV28=0 timeout 300 python tools/conv28_bench.py 64 128 2>&1 | tail -3
CHILD=1 DLB_CONV28=1 timeout 300 python tools/conv28_bench.py 64 2>&1 | tail -2
