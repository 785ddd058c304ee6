#!/bin/bash
# ncu --set full on selected kernels of a LeNet training step; exports the raw and source pages as CSV on the box
# (the .ncu-rep embeds the whole cubin and is too large to bring back).
# usage: gpu_ncu_conv.sh <prefix> <kernel regex> <batch> <skip> <count>
mkdir -p gpurun_out
P=${1:-n}
cat > /tmp/ncu_steps.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, deeplearning_b200 as dl, bench
B = int(sys.argv[1])
net = dl.lenet('f32'); net.init_params(0)
x, y = bench.synthetic_data(32768, 784, 10, np.float32)
net.set_dataset(x, y); net.reset_random_generator(0); net.enable_graphs(False)
net.train_steps(B, 4, 0.01, dl.CostFunctionId.CROSS_ENTROPY, 0.1)
net.ctx.synchronize()
PY
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"${2:-conv_pool}" -s ${4:-4} -c ${5:-2} -o /tmp/${P}_prof python /tmp/ncu_steps.py ${3:-8192} > gpurun_out/${P}_ncu.log 2>&1
ncu -i /tmp/${P}_prof.ncu-rep --page raw --csv > gpurun_out/${P}_raw.csv 2>/dev/null
ncu -i /tmp/${P}_prof.ncu-rep --page source --csv > gpurun_out/${P}_source.csv 2>/dev/null
ls -la gpurun_out | grep ${P}_
