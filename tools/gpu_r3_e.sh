#!/bin/bash
# A/B: input-gradient contraction of the hidden NLayers from the per-step transposed weights (K-major plan) against the transposing plan
mkdir -p gpurun_out
P=${1:-r3e}
echo "== tests"; timeout 1500 python -m pytest tests/test_net_gpu.py tests/test_reference_gpu.py tests/test_step_kernel_gpu.py tests/test_reference_adapter.py tests/test_cpp_host_gpu.py -x -q -m gpu 2>&1 | tail -4
for m in 0 1; do
  echo "== bench f32 DLB_DX_WT=$m"; DLB_DX_WT=$m timeout 900 python bench.py > gpurun_out/${P}_wt${m}_bench_f32.json 2> gpurun_out/${P}_wt${m}_bench_f32.err
  python tools/show_bench.py gpurun_out/${P}_wt${m}_bench_f32.json > gpurun_out/${P}_wt${m}.txt 2>&1; grep -E "^value|full_dX|wT|cfg5|config3" gpurun_out/${P}_wt${m}.txt | cut -c1-220; tail -2 gpurun_out/${P}_wt${m}_bench_f32.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/${P}_wt${m}_bench_f32.json").read().splitlines()[-1])
print("config3", d.get("config3",{}).get("value"), "cfg5", d.get("config5_on_1_gpu",{}).get("value"))
PY
done
