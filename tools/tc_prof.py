"""Development only: phase clocks of tc_gemm_kernel for the LeNet step; needs a library built with tools/dev_build.sh -DDLB_TC_PROFILE
(the release build has no such hooks and no dlb_debug_tc_profile symbol)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deeplearning_b200 as dl
from deeplearning_b200 import _capi
import bench
lib = _capi.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
net = dl.lenet('f32'); net.init_params(0)
x, y = bench.synthetic_data(32768, 784, 10, np.float32)
net.set_dataset(x, y); net.reset_random_generator(0); net.enable_graphs(False)
out = (ctypes.c_ulonglong * 16)()
names = ["wait_copy", "barrier", "issue", "wait_empty", "transform", "fence+arrive", "kblocks", "-", "mma_wait_full", "mma_kblocks"]
net.train_steps(B, 2, 0.01, dl.CostFunctionId.CROSS_ENTROPY, 0.1)
lib.dlb_debug_tc_profile(out)
net.train_steps(B, 4, 0.01, dl.CostFunctionId.CROSS_ENTROPY, 0.1)
lib.dlb_debug_tc_profile(out)
v = list(out)
kb = max(v[6], 1)
print("k-blocks", v[6], "(all 3 contractions, 4 steps)")
for i in range(6):
    print("  %-14s %8.0f clk per k-block" % (names[i], v[i] / kb))
print("  %-14s %8.0f clk per k-block" % (names[8], v[8] / max(v[9], 1)))
