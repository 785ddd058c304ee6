#!/bin/bash
# First GPU call: smoke, gpu tests, bench, ncu launch list + one full capture of the gemm kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_nvsmi.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== bench f32"; timeout 600 python bench.py > gpurun_out/a_bench_f32.json 2> gpurun_out/a_bench_f32.err; tail -c 1500 gpurun_out/a_bench_f32.json; tail -3 gpurun_out/a_bench_f32.err
echo "== bench f64"; timeout 600 python bench.py --real f64 --steps 512 > gpurun_out/a_bench_f64.json 2> gpurun_out/a_bench_f64.err; tail -c 600 gpurun_out/a_bench_f64.json
echo "== bench reference"; timeout 600 python bench.py --impl reference > gpurun_out/a_bench_ref.json 2>&1; tail -c 600 gpurun_out/a_bench_ref.json
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/a_launches.csv python bench.py --steps 4 --warmup 3 > gpurun_out/a_ncu_bench.log 2>&1; wc -l gpurun_out/a_launches.csv
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 30 -c 8 -o gpurun_out/a_prof_gemm python bench.py --steps 4 --warmup 3 > gpurun_out/a_ncu_full.log 2>&1; ls -la gpurun_out/
