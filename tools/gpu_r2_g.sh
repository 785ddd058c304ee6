#!/bin/bash
mkdir -p gpurun_out
echo "== conv28 tests"; timeout 900 python -m pytest tests/test_conv28_gpu.py -x -q -m gpu 2>&1 | tail -12
echo "== conv28 bench"; timeout 900 python tools/conv28_bench.py 2>&1 | tail -14
echo "== mega small2"; timeout 900 python tools/mega_small2.py 2>&1 | grep -v "^$" | tail -14
