#!/bin/bash
mkdir -p gpurun_out
echo "== matmul + net tests"; timeout 1500 python -m pytest tests/test_matmul_gpu.py tests/test_net_gpu.py tests/test_reference_gpu.py tests/test_reference_adapter.py tests/test_ops_gpu.py -x -q -m gpu 2>&1 | tail -6
echo "== bench"; CHILD=1 timeout 300 python tools/conv28_bench.py 256 1024 8192 2>&1 | tail -4 | cut -c1-60
python - <<'PY'
import numpy as np, deeplearning_b200 as dl
net = dl.lenet("f32"); net.init_params(seed=0)
rng = np.random.default_rng(1); n = 32768
x = rng.random((n, 784), dtype=np.float32); y = np.zeros((n, 10), np.float32); y[np.arange(n), np.arange(n) % 10] = 1
net.set_dataset(x, y); net.reset_random_generator(0)
for b in (256, 8192):
    net.train_steps(b, 4, 0.01, dl.CostFunctionId.CROSS_ENTROPY, 0.1)
    dl.timing_enable(True); net.train_steps(b, 8, 0.01, dl.CostFunctionId.CROSS_ENTROPY, 0.1); rep = dl.timing_report(); dl.timing_enable(False)
    print(b, {k.split(":",1)[1]: round(1e3 * t / c, 1) for k, (c, t) in sorted(rep.items(), key=lambda kv: -kv[1][1])[:9]})
PY
