#!/bin/bash
# round 2, last session: vectorised pool kernels + the 128 x 64 double contraction
mkdir -p gpurun_out
echo "== pool + matmul tests"; timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_matmul_gpu.py -x -q -m gpu -k "player or f64" 2>&1 | tail -5
echo "== net tests"; timeout 1500 python -m pytest tests/test_net_gpu.py tests/test_reference_gpu.py tests/test_mnet_gpu.py -x -q -m gpu 2>&1 | tail -5
echo "== op bench"; timeout 600 python tools/bench_ops.py > gpurun_out/r3a_ops.jsonl 2> gpurun_out/r3a_ops.err; cat gpurun_out/r3a_ops.jsonl | cut -c1-260; tail -3 gpurun_out/r3a_ops.err
echo "== cifar"; timeout 600 python tools/bench_cifar.py > gpurun_out/r3a_cifar.json 2> gpurun_out/r3a_cifar.err; cut -c1-330 gpurun_out/r3a_cifar.json; tail -2 gpurun_out/r3a_cifar.err
echo "== bench f64"; timeout 600 python bench.py --real f64 --steps 512 > gpurun_out/r3a_bench_f64.json 2> gpurun_out/r3a_bench_f64.err; python tools/show_bench.py gpurun_out/r3a_bench_f64.json | head -30; tail -3 gpurun_out/r3a_bench_f64.err
