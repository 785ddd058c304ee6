#!/bin/bash
# verification of the committed build + source-level ncu capture of the NLayer input-gradient contraction (B = 8192)
mkdir -p gpurun_out
P=${1:-r3k}
echo "== tests"; timeout 1800 python -m pytest tests/test_net_gpu.py tests/test_matmul_gpu.py tests/test_reference_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "== bench f32"; timeout 900 python bench.py > gpurun_out/${P}_bench_f32.json 2> gpurun_out/${P}_bench_f32.err
python tools/show_bench.py gpurun_out/${P}_bench_f32.json > gpurun_out/${P}.txt 2>&1; grep -E "^value|tc_gemm|cfg5" gpurun_out/${P}.txt | cut -c1-200; tail -2 gpurun_out/${P}_bench_f32.err
bash tools/gpu_ncu_conv.sh ${P}_dx "EpiStoreTimesDeriv" 8192 1 1
python tools/ncu_source_mix.py gpurun_out/${P}_dx_source.csv 25 > gpurun_out/${P}_dx_source_mix.txt 2>&1; head -60 gpurun_out/${P}_dx_source_mix.txt
rm -f gpurun_out/${P}_dx_source.csv
