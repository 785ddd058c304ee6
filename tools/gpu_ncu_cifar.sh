#!/bin/bash
# ncu on the tensor-core contractions of one CIFAR-deep training step (14 tc_gemm launches per step):
#   <prefix>_raw.csv     --set full raw page of all 14 launches of the second step
#   <prefix>_source.csv  source page (SASS + stall samples) of launch number $2 of that step (default 1 = conv2 forward)
mkdir -p gpurun_out
P=${1:-nc}
K=${2:-1}
cat > /tmp/ncu_cifar.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, deeplearning_b200 as dl
B = int(sys.argv[1])
net = dl.cifar_deep('f32'); net.init_params(0)
rng = np.random.default_rng(7)
n = 4096
x = rng.uniform(0, 1, (n, 3072)).astype(np.float32)
y = np.zeros((n, 10), np.float32); y[np.arange(n), np.arange(n) % 10] = 1
net.set_dataset(x, y); net.reset_random_generator(0); net.enable_graphs(False)
net.train_steps(B, 2, 0.01, dl.CostFunctionId.CROSS_ENTROPY, 0.1)
net.ctx.synchronize()
PY
timeout 900 ncu --set full --clock-control none --cache-control none -k regex:tc_gemm -s 14 -c 14 -o /tmp/${P}_all python /tmp/ncu_cifar.py ${3:-1024} > gpurun_out/${P}_ncu_all.log 2>&1
ncu -i /tmp/${P}_all.ncu-rep --page raw --csv > gpurun_out/${P}_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:tc_gemm -s $((14 + K)) -c 1 -o /tmp/${P}_one python /tmp/ncu_cifar.py ${3:-1024} > gpurun_out/${P}_ncu_one.log 2>&1
ncu -i /tmp/${P}_one.ncu-rep --page source --csv > gpurun_out/${P}_source.csv 2>/dev/null
ls -la gpurun_out | grep ${P}_
