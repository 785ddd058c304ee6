#!/bin/bash
echo "== conv28 tests"; timeout 900 python -m pytest tests/test_conv28_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "== bench"; CHILD=1 timeout 300 python tools/conv28_bench.py 256 1024 8192 2>&1 | tail -4
