#!/bin/bash
# ncu --set full on the vectorised pool kernels (CIFAR-deep first PLayer, float, B = 1024) and on the double contraction (LeNet NLayer, B = 8192)
mkdir -p gpurun_out
P=${1:-r3b}
BENCH_OPS_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pool2x2 -c 8 -o /tmp/${P}_pool python tools/bench_ops.py pool > gpurun_out/${P}_pool_ncu.log 2>&1
ncu -i /tmp/${P}_pool.ncu-rep --page raw --csv > gpurun_out/${P}_pool_raw.csv 2>/dev/null
BENCH_OPS_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f64 -c 12 -o /tmp/${P}_gemm python tools/bench_ops.py gemm64 > gpurun_out/${P}_gemm_ncu.log 2>&1
ncu -i /tmp/${P}_gemm.ncu-rep --page raw --csv > gpurun_out/${P}_gemm_raw.csv 2>/dev/null
ncu -i /tmp/${P}_gemm.ncu-rep --page source --csv > gpurun_out/${P}_gemm_source.csv 2>/dev/null
ls -la gpurun_out | grep ${P}_
