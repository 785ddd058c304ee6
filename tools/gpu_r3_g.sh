#!/bin/bash
# epilogue with full-block fast paths + transposed-weights input gradient from 1 024 samples up
mkdir -p gpurun_out
P=${1:-r3g}
echo "== tests"; timeout 1800 python -m pytest tests/test_net_gpu.py tests/test_matmul_gpu.py tests/test_ops_gpu.py tests/test_reference_gpu.py tests/test_step_kernel_gpu.py -x -q -m gpu 2>&1 | tail -4
echo "== bench f32"; timeout 900 python bench.py > gpurun_out/${P}_bench_f32.json 2> gpurun_out/${P}_bench_f32.err
python tools/show_bench.py gpurun_out/${P}_bench_f32.json > gpurun_out/${P}.txt 2>&1; grep -E "^value|tc_gemm|wT|cfg5" gpurun_out/${P}.txt | cut -c1-200; tail -2 gpurun_out/${P}_bench_f32.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${P}_bench_f32.json").read().splitlines()[-1])
print("config3", d.get("config3",{}).get("value"), "cfg5", d.get("config5_on_1_gpu",{}).get("value"), "config4", d.get("config4",{}).get("value"))
PY
echo "== cifar"; timeout 600 python tools/bench_cifar.py --tc 1 > gpurun_out/${P}_cifar.json 2> gpurun_out/${P}_cifar.err || timeout 600 python tools/bench_cifar.py > gpurun_out/${P}_cifar.json 2> gpurun_out/${P}_cifar.err
python - <<PY
import json
for line in open("gpurun_out/${P}_cifar.json"):
    d=json.loads(line)
    print(d["tc_mode"], round(d["value"]), [(k["kernel"], round(k["us_per_launch"])) for k in d["kernels"][:9]])
PY
