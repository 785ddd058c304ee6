#!/bin/bash
# A/B of library builds on the same box: for each devlibs/libdlb200_<name>.so given, LeNet bench kernel table + CIFAR-deep.
mkdir -p gpurun_out
for name in "$@"; do
  export DLB200_LIB=$PWD/devlibs/libdlb200_${name}.so
  echo "=========== $name"
  timeout 600 python tools/bench_cifar.py --tc 1 > gpurun_out/ab_${name}_cifar.json 2> gpurun_out/ab_${name}_cifar.err; python - <<PY
import json
for l in open("gpurun_out/ab_${name}_cifar.json"):
    d = json.loads(l); print("cifar", round(d["value"]), "samples/s", round(d["ms_per_step"], 3), "ms")
    for k in d["kernels"][:9]: print("   %-50s %9.1f us  %.3f" % (k["kernel"], k["us_per_launch"], k["share"]))
PY
  tail -2 gpurun_out/ab_${name}_cifar.err
  timeout 600 python bench.py --steps 1024 > gpurun_out/ab_${name}_bench.json 2> gpurun_out/ab_${name}_bench.err; python tools/show_bench.py gpurun_out/ab_${name}_bench.json | grep -v "roofline\|clocks"; tail -2 gpurun_out/ab_${name}_bench.err
done
