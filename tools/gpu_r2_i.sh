#!/bin/bash
mkdir -p gpurun_out
echo "== mnet + step kernel tests"; timeout 900 python -m pytest tests/test_mnet_gpu.py tests/test_step_kernel_gpu.py -x -q -m gpu 2>&1 | tail -12
echo "== ncu conv28"; bash tools/gpu_ncu_conv.sh c28 conv28 8192 2 2 > /dev/null 2>&1
python tools/ncu_source_mix.py gpurun_out/c28_source.csv 25 > gpurun_out/c28_source_mix.txt 2>&1; head -70 gpurun_out/c28_source_mix.txt
python tools/summarize_ncu.py full gpurun_out/c28_raw.csv gpurun_out/c28_full.txt 2>&1 | head -80
rm -f gpurun_out/c28_source.csv
