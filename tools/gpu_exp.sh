#!/bin/bash
# Quick experiment: matmul + net tests, CIFAR tc 1, LeNet B=8192 kernel table.
mkdir -p gpurun_out
P=${1:-exp}
echo "== pytest"; timeout 900 python -m pytest tests/test_matmul_gpu.py tests/test_net_gpu.py -q -x 2>&1 | tail -5
echo "== cifar"; timeout 600 python tools/bench_cifar.py --tc 1 > gpurun_out/${P}_cifar.json 2> gpurun_out/${P}_cifar.err; python - <<PY
import json
for l in open("gpurun_out/${P}_cifar.json"):
    d = json.loads(l); print(d["tc_mode"], round(d["value"]), "samples/s", round(d["ms_per_step"], 3), "ms")
    for k in d["kernels"][:14]: print("   %-50s %9.1f us  %.3f" % (k["kernel"], k["us_per_launch"], k["share"]))
PY
tail -3 gpurun_out/${P}_cifar.err
echo "== bench f32"; timeout 600 python bench.py --steps 512 > gpurun_out/${P}_bench_f32.json 2> gpurun_out/${P}_bench_f32.err; python tools/show_bench.py gpurun_out/${P}_bench_f32.json; tail -3 gpurun_out/${P}_bench_f32.err
