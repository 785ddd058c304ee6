#!/bin/bash
# bench at the three contraction modes (0 packed-FMA, 1 default, 2 tcgen05 always)
mkdir -p gpurun_out
P=${1:-tc}
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for m in 0 2; do
  echo "== bench --tc $m"; timeout 600 python bench.py --tc $m > gpurun_out/${P}_bench_tc$m.json 2> gpurun_out/${P}_bench_tc$m.err; python tools/show_bench.py gpurun_out/${P}_bench_tc$m.json | grep -v "^  \[B=8192\] [03-]" ; tail -3 gpurun_out/${P}_bench_tc$m.err
done
