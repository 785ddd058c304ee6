#!/bin/bash
# mega on/off A/B through bench.py (short)
mkdir -p gpurun_out
for m in 1 0; do
  echo "== DLB_MEGA=$m"; DLB_MEGA=$m timeout 600 python bench.py --steps 1024 --warmup 32 > gpurun_out/r2c_mega$m.json 2> gpurun_out/r2c_mega$m.err; python tools/show_bench.py gpurun_out/r2c_mega$m.json | head -3; tail -2 gpurun_out/r2c_mega$m.err
done
