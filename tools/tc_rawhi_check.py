"""Development experiment: one tensor-core matmul, result written to argv[1] (compare two library builds bit for bit)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeplearning_b200 import _capi
from tests.test_ops_gpu import Dev
rng = np.random.default_rng(5)
M, N, K = 512, 128, 1024
A = rng.normal(size=(M, K)).astype(np.float32); B = rng.normal(size=(N, K)).astype(np.float32)
dev = Dev()
dA, dB, dC = dev.up(A), dev.up(B), dev.empty((M, N), np.float32)
_capi.check(dev.lib.dlb_matmul_f32(dev.ctx.h, M, N, K, dA, K, 0, dB, K, 0, dC, 1))
got = dev.down(dC, (M, N), np.float32)
np.save(sys.argv[1], got)
want = A.astype(np.float64) @ B.astype(np.float64).T
print(sys.argv[1], "max abs err vs f64", np.abs(got - want).max())
