#!/bin/bash
# Round evidence: GPU tests, both bench arms, the ncu launch list of the bench command itself, and one ncu --set full
# pass over the kernels of a training step (raw page exported to CSV on the box; the .ncu-rep is too large to bring back).
# usage: gpu_final.sh <prefix>
mkdir -p gpurun_out
P=${1:-final}
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print(\"smoke ok\")" 2>&1 | tail -3
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${P}_bench_ref.json 2> gpurun_out/${P}_bench_ref.err; cat gpurun_out/${P}_bench_ref.json | cut -c1-400
echo "== bench f32"; timeout 600 python bench.py > gpurun_out/${P}_bench_f32.json 2> gpurun_out/${P}_bench_f32.err; python tools/show_bench.py gpurun_out/${P}_bench_f32.json; tail -3 gpurun_out/${P}_bench_f32.err
echo "== bench f64"; timeout 600 python bench.py --real f64 --steps 512 > gpurun_out/${P}_bench_f64.json 2> gpurun_out/${P}_bench_f64.err; python tools/show_bench.py gpurun_out/${P}_bench_f64.json | head -16
echo "== CIFAR-deep (BASELINE configs[2]): FMA contractions vs tcgen05 implicit GEMM"; timeout 600 python tools/bench_cifar.py > gpurun_out/${P}_cifar.json 2> gpurun_out/${P}_cifar.err; cut -c1-330 gpurun_out/${P}_cifar.json; tail -2 gpurun_out/${P}_cifar.err
echo "== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${P}_bench_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${P}_ncu_bench.log 2>&1
wc -l gpurun_out/${P}_bench_launches.csv
cat > /tmp/ncu_steps.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, deeplearning_b200 as dl, bench
B = int(sys.argv[1])
net = dl.lenet('f32'); net.init_params(0)
x, y = bench.synthetic_data(32768, 784, 10, np.float32)
net.set_dataset(x, y); net.reset_random_generator(0); net.enable_graphs(False)
net.train_steps(B, 6, 0.01, dl.CostFunctionId.CROSS_ENTROPY, 0.1)
net.ctx.synchronize()
PY
for B in 256 8192; do
  echo "== ncu warm launches B=$B"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 22 -c 33 --csv --log-file gpurun_out/${P}_launches_B${B}.csv python /tmp/ncu_steps.py $B > gpurun_out/${P}_ncu_${B}.log 2>&1
  wc -l gpurun_out/${P}_launches_B${B}.csv
  echo "== ncu full B=$B"
  timeout 900 ncu --set full --clock-control none --cache-control none -s 33 -c 11 -o /tmp/${P}_prof_B${B} python /tmp/ncu_steps.py $B > gpurun_out/${P}_ncu_full_${B}.log 2>&1
  ncu -i /tmp/${P}_prof_B${B}.ncu-rep --page raw --csv > gpurun_out/${P}_full_raw_B${B}.csv 2>/dev/null
done
ls -la gpurun_out | grep ${P}_
echo "== ncu full: tensor-core contractions of one CIFAR-deep step"
bash tools/gpu_ncu_cifar.sh ${P}_cifar 1 > /dev/null 2>&1
rm -f gpurun_out/${P}_cifar_source.csv
ls -la gpurun_out | grep ${P}_cifar
