"""GPU parity of the large-batch kernels of the first LeNet stage (csrc/conv28.cu: dense FFMA2 convolution + untied bias +
activation + 2x2 max|min pooling forward, routed-gradient correlation backward, one-byte window selectors).  They take over
from mini-batch 128 up (DLB_CONV28_MIN); DLB_CONV28=1 forces them at every size so that ragged passes and sample groups are
covered at sizes the oracle finishes in seconds.  Checked against the CPU oracle (CLayer.inl:108-197, PLayer.inl:115-170),
against the small-batch kernels of fused.cuh (forward bit for bit), and for run-to-run bit-identity."""
import numpy as np
import pytest

import deeplearning_b200 as dl
from oracle import binding as ob
from tests import helpers as H

pytestmark = pytest.mark.gpu


def lenet_like(filters, hidden, act0, pool, act1, head_act=ob.SOFTMAX, n_out=10):
    return [ob.conv((1, 28, 28), (5, 5), filters, act0), ob.pool((filters, 24, 24), (2, 2), pool),
            ob.full((filters, 12, 12), hidden, act1), ob.full((1, 1, hidden), n_out, head_act)]


VARIANTS = {
    "lenet": (lenet_like(20, 100, ob.RELU, ob.POOL_MAX, ob.RELU), ob.CROSS_ENTROPY),
    "sigmoid_min": (lenet_like(6, 30, ob.SIGMOID, ob.POOL_MIN, ob.TANH), ob.CROSS_ENTROPY),
    "tanh_wide": (lenet_like(32, 48, ob.TANH, ob.POOL_MAX, ob.SIGMOID, head_act=ob.SIGMOID, n_out=7), ob.SQUARED_ERROR),
}


def make(specs, seed, mode, monkeypatch):
    if mode is None: monkeypatch.delenv("DLB_CONV28", raising=False)
    else: monkeypatch.setenv("DLB_CONV28", mode)
    orc, gpu = H.make_pair(specs, "f32", seed=seed, tc_mode=1)
    plan = gpu.describe_plan()
    assert plan.startswith("fused: conv_pool(5x5"), plan
    assert ("conv28.cu" in plan) == (mode != "0"), plan
    return orc, gpu


@pytest.mark.parametrize("name", list(VARIANTS))
@pytest.mark.parametrize("batch,n_items", [(256, 1024), (37, 150), (300, 700)])
def test_learn_matches_oracle(name, batch, n_items, monkeypatch):
    """Two epochs of Net::learn with the kernels forced on: several passes per CTA, partial passes, sample groups of unequal size,
    the tail-batch rule (Net.inl:235-236); predicted classes exactly."""
    specs, cost = VARIANTS[name]
    orc, gpu = make(specs, 3, "1", monkeypatch)
    x, y = H.synthetic_batch(orc, n_items, "f32", seed=4, lo=0.0, hi=1.0, one_hot=cost == ob.CROSS_ENTROPY)
    epochs, lr, lam, seed = 2, 0.05, 0.2, 6
    orc.learn(x, y, batch, epochs, lr, cost, lam, threads=1, perms=dl.shuffle_sequence(seed, n_items, epochs))
    gpu.reset_random_generator(seed)
    gpu.learn(x, y, batch, epochs, lr, dl.CostFunctionId(cost), lam)
    H.assert_close("parameters after learn", gpu.get_params(), orc.get_params(), "f32", rtol_mult=20)
    assert np.array_equal(gpu.predict(x[:200]), orc.eval(x[:200])), "predicted classes differ"


def test_default_threshold_large_batch_vs_oracle(monkeypatch):
    """No environment override: mini-batch 1100 is above the threshold and uses the 512-thread forward (40 passes of 28 samples over
    the CTA columns; 14 sample groups of 78-79 in the backward); three steps of LeNet against the oracle."""
    specs, cost = VARIANTS["lenet"]
    orc, gpu = make(specs, 5, None, monkeypatch)
    n, batch = 3300, 1100
    x, y = H.synthetic_batch(orc, n, "f32", seed=9, lo=0.0, hi=1.0)
    orc.learn(x, y, batch, 1, 0.05, cost, 0.3, threads=1, perms=dl.shuffle_sequence(2, n, 1))
    gpu.reset_random_generator(2)
    gpu.learn(x, y, batch, 1, 0.05, dl.CostFunctionId(cost), 0.3)
    H.assert_close("parameters after 3 steps of 1100", gpu.get_params(), orc.get_params(), "f32", rtol_mult=20)
    assert np.array_equal(gpu.predict(x[:300]), orc.eval(x[:300])), "predicted classes differ"


@pytest.mark.parametrize("batch", [64, 1000])
def test_forward_is_bit_identical_to_the_small_batch_kernel(batch, monkeypatch):
    """Same tap order, same fused multiply-adds, bias added last: the pooled values of both forward kernels are the same bits; the
    parameters after the step agree within the bar (the backward sums in a different order)."""
    specs, cost = VARIANTS["lenet"]
    pooled, params, maps = [], [], []
    x = y = None
    for mode in ("1", "0"):
        orc, gpu = make(specs, 7, mode, monkeypatch)
        if x is None: x, y = H.synthetic_batch(orc, batch, "f32", seed=8, lo=0.0, hi=1.0)
        gpu.set_dataset(x, y)
        gpu.step_indices(np.arange(batch, dtype=np.int64), 0.05, dl.CostFunctionId(cost), reg_factor=0.001)
        gpu._last_batch = batch
        pooled.append(gpu.read(dl.BUF_OUTPUT, 1)); params.append(gpu.get_params()); maps.append(gpu.read(dl.BUF_INDEX_MAP, 1))
    assert np.array_equal(pooled[0], pooled[1]), "pooled values differ between the two forward kernels"
    # the one-byte window selectors read back as the reference's Trace.IndexMapping (dlb_net_read expands them)
    assert maps[0].dtype == maps[1].dtype and np.array_equal(maps[0], maps[1]), "arg-max index maps differ"
    H.assert_close("one step, conv28 vs fused.cuh kernels", params[0], params[1], "f32", rtol_mult=20)


def test_deterministic(monkeypatch):
    specs, cost = VARIANTS["lenet"]
    results = []
    x = y = None
    for _ in range(2):
        orc, gpu = make(specs, 11, "1", monkeypatch)
        if x is None: x, y = H.synthetic_batch(orc, 1536, "f32", seed=12, lo=0.0, hi=1.0)
        gpu.reset_random_generator(5)
        gpu.learn(x, y, 512, 2, 0.05, dl.CostFunctionId(cost), 0.1)
        results.append(gpu.get_params())
    assert np.array_equal(results[0], results[1]), "two runs differ"


@pytest.mark.skipif(not ob.ref_available(), reason="oracle/_ref not on this box")
def test_batch_2048_vs_live_reference(monkeypatch):
    """Two steps of 2048 (the per-GPU share of BASELINE configs[4] at 4 GPUs) against the reference's own Net<CpuDC>::learn."""
    monkeypatch.delenv("DLB_CONV28", raising=False)
    specs = ob.lenet_specs()
    n, batch, lr, lam, seed = 4096, 2048, 0.05, 0.3, 4
    params = H.synthetic_params(301130, specs, "f32", seed=7)
    rng = np.random.default_rng(8)
    x = rng.random((n, 784)).astype(np.float32); y = ob.one_hot_labels("f32", n)
    ref = ob.RefNet(specs, "f32", seed=0)
    ref.set_params(params)
    ref.learn(x, y, batch, 1, lr, ob.CROSS_ENTROPY, lam, parallelism=-1, shuffle_seed=seed)
    gpu = dl.Net([H.to_dl_spec(s) for s in specs], real="f32")
    gpu.set_params(params)
    assert "conv28.cu" in gpu.describe_plan()
    gpu.reset_random_generator(seed)
    gpu.learn(x, y, batch, 1, lr, dl.CostFunctionId.CROSS_ENTROPY, lam)
    H.assert_close("LeNet parameters after 2 steps of 2048 vs the reference", gpu.get_params(), ref.get_params(), "f32", rtol_mult=20)
    assert np.array_equal(gpu.predict(x[:512]), ref.eval(x[:512])), "predicted classes differ"
