#!/bin/bash
mkdir -p gpurun_out
P=${1:-r3d}
echo "== tests"; timeout 1500 python -m pytest tests/test_matmul_gpu.py tests/test_mnet_gpu.py tests/test_net_gpu.py -x -q -m gpu -k "f64" 2>&1 | tail -3
echo "== op bench"; timeout 600 python tools/bench_ops.py gemm64 > gpurun_out/${P}_ops.jsonl 2> gpurun_out/${P}_ops.err; cut -c1-200 gpurun_out/${P}_ops.jsonl; tail -3 gpurun_out/${P}_ops.err
echo "== bench f64"; timeout 600 python bench.py --real f64 --steps 512 > gpurun_out/${P}_bench_f64.json 2> gpurun_out/${P}_bench_f64.err; python tools/show_bench.py gpurun_out/${P}_bench_f64.json 2>/dev/null | grep -v "^  [0-9-]" | head -14
