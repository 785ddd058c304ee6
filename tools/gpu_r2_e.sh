#!/bin/bash
# Round 2, GPU call E: the step-kernel and reference-adapter tests, then the rest of the gpu suite.
mkdir -p gpurun_out
echo "== new tests"; timeout 900 python -m pytest tests/test_step_kernel_gpu.py tests/test_reference_adapter.py tests/test_net_gpu.py -x -q -m gpu 2>&1 | tail -12
echo "== rest"; timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_step_kernel_gpu.py --deselect tests/test_reference_adapter.py --deselect tests/test_net_gpu.py 2>&1 | tail -6
