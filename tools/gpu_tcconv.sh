#!/bin/bash
# Tensor-core convolution check: gpu tests, CIFAR-deep (BASELINE configs[2]) FMA vs tcgen05, LeNet bench.
# usage: gpu_tcconv.sh <prefix>
mkdir -p gpurun_out
P=${1:-tcconv}
echo "== pytest matmul + net"; timeout 900 python -m pytest tests/test_matmul_gpu.py tests/test_net_gpu.py -q 2>&1 | tail -30
echo "== cifar"; timeout 600 python tools/bench_cifar.py > gpurun_out/${P}_cifar.json 2> gpurun_out/${P}_cifar.err; cut -c1-1800 gpurun_out/${P}_cifar.json; tail -3 gpurun_out/${P}_cifar.err
echo "== pytest rest"; timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_matmul_gpu.py --deselect tests/test_net_gpu.py 2>&1 | tail -8
echo "== bench f32"; timeout 600 python bench.py > gpurun_out/${P}_bench_f32.json 2> gpurun_out/${P}_bench_f32.err; python tools/show_bench.py gpurun_out/${P}_bench_f32.json; tail -3 gpurun_out/${P}_bench_f32.err
