#!/bin/bash
# strided fetch of the double contraction + graph replay of the recurrent step
mkdir -p gpurun_out
P=${1:-r3c}
echo "== tests"; timeout 1500 python -m pytest tests/test_matmul_gpu.py tests/test_mnet_gpu.py tests/test_net_gpu.py tests/test_reference_gpu.py -x -q -m gpu 2>&1 | tail -4
echo "== op bench"; timeout 600 python tools/bench_ops.py gemm64 > gpurun_out/${P}_ops.jsonl 2> gpurun_out/${P}_ops.err; cut -c1-200 gpurun_out/${P}_ops.jsonl; tail -3 gpurun_out/${P}_ops.err
echo "== bench f32"; timeout 900 python bench.py > gpurun_out/${P}_bench_f32.json 2> gpurun_out/${P}_bench_f32.err; python tools/show_bench.py gpurun_out/${P}_bench_f32.json | head -4; tail -3 gpurun_out/${P}_bench_f32.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${P}_bench_f32.json").read().splitlines()[-1])
for k in ("config3","f64","config4"):
    v=d.get(k,{})
    print(k, {kk:v.get(kk) for kk in ("value","ms_per_step","launches_per_step","error") if kk in v})
PY
echo "== bench f64"; timeout 600 python bench.py --real f64 --steps 512 > gpurun_out/${P}_bench_f64.json 2> gpurun_out/${P}_bench_f64.err; python tools/show_bench.py gpurun_out/${P}_bench_f64.json | grep -v "^  [0-9-]" | head -30
