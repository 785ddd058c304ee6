#!/bin/bash
mkdir -p gpurun_out
echo "== conv28 tests"; timeout 900 python -m pytest tests/test_conv28_gpu.py tests/test_net_gpu.py -x -q -m gpu 2>&1 | tail -6
for SK in 0 1200 1700 2400; do
  for ST in 0 1; do
    echo "== skew $SK stagger $ST"; CHILD=1 DLB_CONV28=1 DLB_CONV28_SKEW=$SK DLB_CONV28_STAGGER=$ST timeout 300 python tools/conv28_bench.py 256 1024 8192 2>&1 | tail -3
  done
done
