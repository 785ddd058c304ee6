#!/bin/bash
# Round 2, GPU call A: MMA descriptor probe, the new one-hop parity tests, smoke(), a short bench line.
mkdir -p gpurun_out
echo "== mma_probe"; timeout 60 tools/_build/mma_probe 2>&1 | tail -15
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest new"; timeout 1500 python -m pytest tests/test_reference_gpu.py "tests/test_net_gpu.py::test_lenet_step_full_size" -x -q -m gpu 2>&1 | tail -15
echo "== bench f32"; timeout 600 python bench.py --steps 512 --warmup 32 > gpurun_out/r2a_bench_f32.json 2> gpurun_out/r2a_bench_f32.err; python tools/show_bench.py gpurun_out/r2a_bench_f32.json; tail -3 gpurun_out/r2a_bench_f32.err
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
