#!/bin/bash
# GPU call: tests, bench (f32/f64), ncu launch list + full capture of the top kernels of the fused path.
mkdir -p gpurun_out
P=${1:-b}
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench f32"; timeout 600 python bench.py > gpurun_out/${P}_bench_f32.json 2> gpurun_out/${P}_bench_f32.err; python tools/show_bench.py gpurun_out/${P}_bench_f32.json; tail -3 gpurun_out/${P}_bench_f32.err
echo "== bench f64"; timeout 600 python bench.py --real f64 --steps 512 > gpurun_out/${P}_bench_f64.json 2> gpurun_out/${P}_bench_f64.err; python tools/show_bench.py gpurun_out/${P}_bench_f64.json
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${P}_launches.csv python bench.py --steps 4 --warmup 3 > gpurun_out/${P}_ncu_bench.log 2>&1; wc -l gpurun_out/${P}_launches.csv
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|conv_pool|full_head|reduce_update' -s 24 -c 10 -o gpurun_out/${P}_prof python bench.py --steps 4 --warmup 3 > gpurun_out/${P}_ncu_full.log 2>&1; ls -la gpurun_out/ | tail -12
