#!/bin/bash
# GPU tests + f64 and f32 bench
mkdir -p gpurun_out
P=${1:-d}
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== bench f64"; timeout 600 python bench.py --real f64 --steps 512 > gpurun_out/${P}_bench_f64.json 2> gpurun_out/${P}_bench_f64.err; python tools/show_bench.py gpurun_out/${P}_bench_f64.json 2>/dev/null | head -30; tail -3 gpurun_out/${P}_bench_f64.err
echo "== bench f32"; timeout 600 python bench.py > gpurun_out/${P}_bench_f32.json 2> gpurun_out/${P}_bench_f32.err; python tools/show_bench.py gpurun_out/${P}_bench_f32.json 2>/dev/null | head -4
