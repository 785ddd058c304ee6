#!/bin/bash
# tests with dev library $1, then alternating bench lines of the given dev libraries (kernel tables for the first round)
export DLB200_LIB=$PWD/devlibs/libdlb200_$1.so
timeout 900 python -m pytest tests/test_matmul_gpu.py tests/test_net_gpu.py -q -x 2>&1 | tail -4
for round in 1 2; do
for name in "$@"; do
  export DLB200_LIB=$PWD/devlibs/libdlb200_${name}.so
  timeout 600 python bench.py --steps 1024 > /tmp/ab_${name}.json 2> /tmp/ab_${name}.err
  if [ $round -eq 1 ]; then python tools/show_bench.py /tmp/ab_${name}.json 2>/dev/null | grep -E "^value|tc_gemm|finalize" | sed "s/^/$name: /"; else python tools/show_bench.py /tmp/ab_${name}.json 2>/dev/null | head -1 | sed "s/^/$name: /"; fi
  tail -2 /tmp/ab_${name}.err
done
done
