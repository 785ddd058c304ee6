#!/bin/bash
# Last check of a round: smoke, the gpu tests, the bench line, and a source-level ncu capture of the TMA-fed NLayer forward.
mkdir -p gpurun_out
P=${1:-last}
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/${P}_bench_f32.json 2> gpurun_out/${P}_bench_f32.err; python tools/show_bench.py gpurun_out/${P}_bench_f32.json 2>/dev/null | head -3; tail -2 gpurun_out/${P}_bench_f32.err
bash tools/gpu_ncu_conv.sh ${P}_fwd "tc_gemm_kernel.*EpiFullForward" 8192 2 1 > /dev/null 2>&1
python tools/ncu_source_mix.py gpurun_out/${P}_fwd_source.csv 12 > gpurun_out/${P}_source_mix_tc_fwd.txt 2>&1; head -40 gpurun_out/${P}_source_mix_tc_fwd.txt
rm -f gpurun_out/${P}_fwd_source.csv
