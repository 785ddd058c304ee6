#!/bin/bash
# parity tests with one dev library, then the A/B table
export DLB200_LIB=$PWD/devlibs/libdlb200_$1.so
timeout 900 python -m pytest tests/test_matmul_gpu.py tests/test_net_gpu.py -q -x 2>&1 | tail -4
unset DLB200_LIB
bash tools/gpu_ab.sh "$@"
