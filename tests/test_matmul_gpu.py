"""GPU: the tcgen05 3xTF32 contraction (tc_gemm.cuh) and the FMA contraction (gemm.cuh) against float64 numpy, for every
operand layout the NLayer forward / weight-gradient / input-gradient contractions use, ragged sizes and split reductions.
Bound: |C - C64| <= 4e-6 * (|A| |B|^T) element-wise (fp32-level; plain TF32 would be ~1e-3)."""
import ctypes as C

import numpy as np
import pytest

import deeplearning_b200 as dl
from deeplearning_b200 import _capi
from tests.test_ops_gpu import Dev

pytestmark = pytest.mark.gpu

SHAPES = [
    (256, 100, 2880),    # LeNet NLayer forward (split reduction)
    (100, 2881, 256),    # LeNet weight + bias gradient
    (256, 2880, 100),    # LeNet input gradient (K = 100: ragged k-block)
    (1, 1, 1), (5, 7, 3), (128, 128, 32), (129, 130, 33), (300, 17, 1000), (1000, 260, 70),
]


@pytest.mark.parametrize("use_tc", [1, 0])
@pytest.mark.parametrize("a_t,b_t", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_matmul_layouts(M, N, K, a_t, b_t, use_tc):
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.normal(size=(M, K)).astype(np.float32)
    B = rng.normal(size=(N, K)).astype(np.float32)
    dev = Dev()
    try:
        dA = dev.up(np.ascontiguousarray(A.T) if a_t else A)
        dB = dev.up(np.ascontiguousarray(B.T) if b_t else B)
        dC = dev.empty((M, N), np.float32)
        _capi.check(dev.lib.dlb_matmul_f32(dev.ctx.h, M, N, K, dA, M if a_t else K, a_t, dB, N if b_t else K, b_t, dC, use_tc))
        got = dev.down(dC, (M, N), np.float32).astype(np.float64)
    finally:
        dev.close()
    want = A.astype(np.float64) @ B.astype(np.float64).T
    bound = 4e-6 * (np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T) + 1e-30
    err = np.abs(got - want)
    worst = float(np.max(err / bound))
    assert worst <= 1.0, f"max err/bound {worst:.3f} (max abs err {err.max():.3e})"


@pytest.mark.parametrize("a_t,b_t", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K", SHAPES + [(8192, 100, 331), (100, 700, 4099)])
def test_matmul_f64_layouts(M, N, K, a_t, b_t):
    """Real = double: the 128 x 64 DFMA micro-kernel (and the 32 x 32 one for skinny outputs) with its wave-filling reduction splits,
    every operand layout; bound 4e-15 * (|A| |B|^T) element-wise (double rounding level: the products are exact in fp64 up to 2^-53
    each, the sums are ordered differently from numpy's)."""
    rng = np.random.default_rng(M * 5 + N * 11 + K)
    A = rng.normal(size=(M, K))
    B = rng.normal(size=(N, K))
    dev = Dev()
    try:
        dA = dev.up(np.ascontiguousarray(A.T) if a_t else A)
        dB = dev.up(np.ascontiguousarray(B.T) if b_t else B)
        dC = dev.empty((M, N), np.float64)
        _capi.check(dev.lib.dlb_matmul_f64(dev.ctx.h, M, N, K, dA, M if a_t else K, a_t, dB, N if b_t else K, b_t, dC))
        got = dev.down(dC, (M, N), np.float64)
    finally:
        dev.close()
    want = A @ B.T
    bound = 4e-15 * (np.abs(A) @ np.abs(B).T) * max(1.0, np.sqrt(K) / 8) + 1e-300
    err = np.abs(got - want)
    worst = float(np.max(err / bound))
    assert worst <= 1.0, f"max err/bound {worst:.3f} (max abs err {err.max():.3e})"
