#!/bin/bash
# Quick GPU check: the given pytest selection (default: everything gpu-marked), then a short bench.
mkdir -p gpurun_out
P=${1:-q}; shift
echo "== pytest $@"; timeout 900 python -m pytest ${@:-tests -m gpu} -x -q 2>&1 | tail -25
echo "== bench f32"; timeout 600 python bench.py > gpurun_out/${P}_bench_f32.json 2> gpurun_out/${P}_bench_f32.err; python tools/show_bench.py gpurun_out/${P}_bench_f32.json; tail -3 gpurun_out/${P}_bench_f32.err
