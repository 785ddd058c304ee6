#!/bin/bash
# A/B of library builds on the same box, bench line only (value, e2e): alternating, two rounds each.
for round in 1 2; do
for name in "$@"; do
  export DLB200_LIB=$PWD/devlibs/libdlb200_${name}.so
  timeout 600 python bench.py --steps 1024 > /tmp/ab_${name}.json 2> /tmp/ab_${name}.err; python tools/show_bench.py /tmp/ab_${name}.json 2>/dev/null | head -1 | sed "s/^/$name: /"; tail -2 /tmp/ab_${name}.err
done
done
export DLB200_LIB=$PWD/devlibs/libdlb200_$2.so
timeout 600 python -m pytest tests/test_net_gpu.py -q -x -k "async or host or train_steps" 2>&1 | tail -3
