#!/bin/bash
# Round 2, GPU call B: the megakernel through the test suite and the bench.
mkdir -p gpurun_out
echo "== mega_debug"; timeout 300 python tools/mega_debug.py 256 2>&1 | grep -v "^$" | tail -40
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "== bench f32"; timeout 600 python bench.py --steps 1024 --warmup 32 > gpurun_out/r2b_bench_f32.json 2> gpurun_out/r2b_bench_f32.err; python tools/show_bench.py gpurun_out/r2b_bench_f32.json; tail -3 gpurun_out/r2b_bench_f32.err
