"""GPU parity of the step megakernel (csrc/mega.cu, dlb_net_enable_step_kernel): the whole mini-batch step of Net::learn as one
persistent cooperative kernel, all consecutive full mini-batches of an epoch inside one launch.  Checked against the CPU oracle
(and the live reference where its library is on the box) on LeNet-shaped nets: different filter counts / hidden widths /
activations / pooling operators, ragged sample groups, several passes and chunks per CTA, the tail-batch rule, host-buffer steps
with the batch cost, and run-to-run bit-identity (every sum is ordered)."""
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
    "sigmoid_min": (lenet_like(4, 36, ob.SIGMOID, ob.POOL_MIN, ob.TANH), ob.CROSS_ENTROPY),
    "tanh_sq": (lenet_like(8, 52, ob.TANH, ob.POOL_MAX, ob.SIGMOID, head_act=ob.SIGMOID, n_out=7), ob.SQUARED_ERROR),
}


def make(specs, seed):
    orc, gpu = H.make_pair(specs, "f32", seed=seed, tc_mode=1)
    gpu.enable_step_kernel(True)
    assert gpu.describe_plan().startswith("fused: one cooperative kernel per step"), gpu.describe_plan()
    return orc, gpu


@pytest.mark.parametrize("name", list(VARIANTS))
@pytest.mark.parametrize("batch,n_items", [(256, 1024), (130, 400), (37, 150), (300, 700)])
def test_learn_matches_oracle(name, batch, n_items):
    """Two epochs of Net::learn: full mini-batches run as multi-step launches, the tail (fewer than 1.5 x batch remaining items
    form one batch, Net.inl:235-236) as its own step; predicted classes exactly.  Mini-batches below 128 stay on the graph path
    (the (37, 150) case checks that hand-over)."""
    specs, cost = VARIANTS[name]
    orc, gpu = make(specs, seed=3)
    x, y = H.synthetic_batch(orc, n_items, "f32", seed=4, lo=0.0, hi=1.0, one_hot=cost == ob.CROSS_ENTROPY)
    epochs, lr, lam, seed = 2, 0.05, 0.2, 6
    orc.learn(x, y, batch, epochs, lr, cost, lam, threads=1, perms=dl.shuffle_sequence(seed, n_items, epochs))
    gpu.reset_random_generator(seed)
    gpu.learn(x, y, batch, epochs, lr, dl.CostFunctionId(cost), lam)
    H.assert_close("parameters after learn", gpu.get_params(), orc.get_params(), "f32", rtol_mult=20)
    assert np.array_equal(gpu.predict(x[:200]), orc.eval(x[:200])), "predicted classes differ"


def test_train_steps_equal_learn_and_are_deterministic():
    """train_steps (resumable, multi-step launches split at arbitrary points) == learn, bit for bit, and a repeated run is
    bit-identical."""
    specs, cost = VARIANTS["lenet"]
    n, batch = 2048, 256
    results = []
    x = y = None
    for mode in ("learn", "steps", "steps"):
        orc, gpu = make(specs, seed=11)
        if x is None:
            x, y = H.synthetic_batch(orc, n, "f32", seed=12, lo=0.0, hi=1.0)
        gpu.reset_random_generator(5)
        if mode == "learn":
            gpu.learn(x, y, batch, 2, 0.05, dl.CostFunctionId(cost), 0.1)
        else:
            gpu.set_dataset(x, y)
            assert gpu.train_steps(batch, 3, 0.05, dl.CostFunctionId(cost), 0.1) == 3 * batch
            assert gpu.train_steps(batch, 13, 0.05, dl.CostFunctionId(cost), 0.1) == 13 * batch
        results.append(gpu.get_params())
    assert np.array_equal(results[0], results[1]), "train_steps differs from learn"
    assert np.array_equal(results[1], results[2]), "the step kernel is not deterministic"


def test_step_kernel_equals_graph_path_within_the_bar():
    """The same five steps through the step kernel and through the default multi-kernel CUDA-graph path."""
    specs, cost = VARIANTS["lenet"]
    orc, mega = make(specs, seed=21)
    _, graph = H.make_pair(specs, "f32", seed=21, tc_mode=1)
    x, y = H.synthetic_batch(orc, 1280, "f32", seed=22, lo=0.0, hi=1.0)
    for net in (mega, graph):
        net.set_dataset(x, y)
        net.reset_random_generator(2)
        net.train_steps(256, 5, 0.05, dl.CostFunctionId(cost), 0.3)
    H.assert_close("step kernel vs graph path", mega.get_params(), graph.get_params(), "f32", rtol_mult=20)


def test_host_buffer_steps_and_batch_cost():
    """dlb_net_step_host through the step kernel: parameters and the batch cost (summed in the head phase) against the oracle."""
    specs, cost = VARIANTS["lenet"]
    orc, gpu = make(specs, seed=31)
    n = 256
    perm = np.arange(n, dtype=np.int64)[None, :]
    for i in range(3):
        x, y = H.synthetic_batch(orc, n, "f32", seed=40 + i, lo=0.0, hi=1.0)
        outs = orc.eval(x, want_outputs=True)[1]
        want_cost = sum(orc.cost(outs[k], y[k], cost) for k in range(n))
        got_cost = gpu.step_host(x, y, 0.05, dl.CostFunctionId(cost), reg_factor=0.001)
        assert abs(got_cost - want_cost) <= 1e-4 * abs(want_cost), (got_cost, want_cost)
        orc.learn(x, y, n, 1, 0.05, cost, 0.001 * n, threads=1, perms=perm)
    H.assert_close("parameters after three host-buffer steps", gpu.get_params(), orc.get_params(), "f32", rtol_mult=20)


@pytest.mark.skipif(not ob.ref_available(), reason="oracle/_ref not on this box")
def test_lenet_b256_vs_live_reference():
    """BASELINE configs[1] through the step kernel against the reference's own Net<CpuDC>::learn (all host threads)."""
    specs = ob.lenet_specs()
    n, batch, lr, lam, seed = 1280, 256, 0.05, 0.3, 4
    params = H.synthetic_params(301130, specs, "f32", seed=7)
    rng = np.random.default_rng(8)
    x = rng.random((n, 784)).astype(np.float32); y = ob.one_hot_labels("f32", n)
    ref = ob.RefNet(specs, "f32", seed=0)
    ref.set_params(params)
    ref.learn(x, y, batch, 1, lr, ob.CROSS_ENTROPY, lam, parallelism=-1, shuffle_seed=seed)
    gpu = dl.Net([H.to_dl_spec(s) for s in specs], real="f32")
    gpu.set_params(params)
    gpu.enable_step_kernel(True)
    gpu.reset_random_generator(seed)
    gpu.learn(x, y, batch, 1, lr, dl.CostFunctionId.CROSS_ENTROPY, lam)
    H.assert_close("LeNet parameters after 5 steps of 256 vs the reference", gpu.get_params(), ref.get_params(), "f32", rtol_mult=20)
    assert np.array_equal(gpu.predict(x[:512]), ref.eval(x[:512])), "predicted classes differ"
