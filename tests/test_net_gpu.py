"""GPU parity of whole nets against the CPU oracle (oracle/dl_oracle.c), through the C ABI (dlb_net_*).

What is compared (BASELINE.json north_star): layer outputs, activation derivatives, gradients and updated weights
within the stated relative tolerance (helpers.RTOL: 1e-5 float, 1e-12 double, atol = rtol * max|reference|);
pooling index maps and predicted classes EXACTLY.  Nothing in the reference pins Net<GpuDC>::learn
(DeepLearningCudaTest/NetCudaTest.cpp only instantiates the template), so whole-step parity is defined here.
"""
import numpy as np
import pytest

import deeplearning_b200 as dl
from oracle import binding as ob
from tests import helpers as H

pytestmark = pytest.mark.gpu

MINI_LENET = [ob.conv((1, 12, 12), (3, 3), 4, ob.RELU), ob.pool((4, 10, 10), (2, 2), ob.POOL_MAX),
              ob.full((4, 5, 5), 16, ob.RELU), ob.full((1, 1, 16), 10, ob.SOFTMAX)]
# the reference's layer-test fixtures chained into one net (Shared/NetLayersTestUtils.h:34-65)
FIXTURE_NET = [ob.conv((5, 13, 17), (3, 3), 7, ob.SIGMOID, pad=(3, 6), stride=(2, 2)), ob.pool((7, 9, 14), (3, 4), ob.POOL_MIN),
               ob.full((7, 3, 3), 23, ob.TANH), ob.full((1, 1, 23), 5, ob.LINEAR)]
AVG_NET = [ob.conv((2, 9, 8), (2, 3), 3, ob.TANH, pad=(1, 0), stride=(1, 2)), ob.pool((3, 10, 3), (3, 2), ob.POOL_AVERAGE),
           ob.conv((3, 3, 1), (2, 1), 5, ob.RELU, pad=(1, 1)), ob.full((5, 4, 3), 4, ob.SOFTMAX)]
MINI_MLP = [ob.full((1, 1, 20), 15, ob.SIGMOID), ob.full((1, 1, 15), 10, ob.SIGMOID)]

NETS = {"mini_lenet": (MINI_LENET, ob.CROSS_ENTROPY), "fixture": (FIXTURE_NET, ob.SQUARED_ERROR),
        "avg": (AVG_NET, ob.CROSS_ENTROPY), "mini_mlp": (MINI_MLP, ob.CROSS_ENTROPY)}


def compare_traces(orc, gpu, x, y, cost_id, real, scale=0.0, grads_in=None):
    want = H.oracle_batch_trace(orc, x, y, cost_id, grads_in, scale)
    if grads_in is not None:
        gpu.write_gradient(grads_in)
    gpu.forward_backward(x, y, cost_id, scale)
    out_off = H.layer_offsets(orc.out_sizes)
    in_off = H.layer_offsets(orc.in_sizes)
    for l, s in enumerate(orc.specs):
        sl = slice(out_off[l], out_off[l] + orc.out_sizes[l])
        H.assert_close(f"layer {l} output", gpu.read(dl.BUF_OUTPUT, l), want["acts"][:, sl], real)
        if s.type != ob.PULL:
            H.assert_close(f"layer {l} derivatives", gpu.read(dl.BUF_DERIVATIVES, l), want["derivs"][:, sl], real, rtol_mult=5)
        elif s.pool_type != ob.POOL_AVERAGE:
            got = gpu.read(dl.BUF_INDEX_MAP, l)
            assert np.array_equal(got, want["index_map"][:, sl]), f"layer {l}: pooling index map differs"
        if l > 0:
            isl = slice(in_off[l], in_off[l] + orc.in_sizes[l])
            H.assert_close(f"layer {l} input gradient", gpu.read(dl.BUF_INPUT_GRAD, l), want["in_grads"][:, isl], real, rtol_mult=5)
    H.assert_close("cost gradient", gpu.read(dl.BUF_COST_GRAD), want["cost_grad"], real, rtol_mult=5)
    got_g = gpu.read(dl.BUF_GRADIENT)
    for l, s in enumerate(orc.specs):   # per layer, so that each gradient is judged against its own scale
        a = gpu.layer_param_offset(l)
        b = gpu.layer_param_offset(l + 1) if l + 1 < len(orc.specs) else gpu.param_count
        if b > a:
            H.assert_close(f"layer {l} parameter gradient", got_g[a:b], want["grads"][a:b], real, rtol_mult=5)


@pytest.mark.parametrize("real", ["f32", "f64"])
@pytest.mark.parametrize("name", list(NETS))
@pytest.mark.parametrize("batch", [1, 7])
def test_forward_backward_matches_oracle(name, real, batch):
    specs, cost = NETS[name]
    orc, gpu = H.make_pair(specs, real, seed=3)
    x, y = H.synthetic_batch(orc, batch, real, seed=5, one_hot=cost == ob.CROSS_ENTROPY)
    compare_traces(orc, gpu, x, y, cost, real)


@pytest.mark.parametrize("real", ["f32", "f64"])
def test_gradient_scale_factor(real):
    """backpropagate(..., scale s) must yield s*G_in + G_pure (Shared/NetLayersTestUtils.h:70-100)."""
    specs, cost = NETS["fixture"]
    orc, gpu = H.make_pair(specs, real, seed=4)
    x, y = H.synthetic_batch(orc, 3, real, seed=6, one_hot=False)
    g_in = np.random.default_rng(9).normal(size=orc.param_count).astype(orc.dtype)
    compare_traces(orc, gpu, x, y, cost, real, scale=0.37, grads_in=g_in)


@pytest.mark.parametrize("real", ["f32", "f64"])
@pytest.mark.parametrize("name,batch,n_items,lam", [("mini_lenet", 8, 50, 0.1), ("mini_mlp", 10, 64, 0.0), ("avg", 5, 23, 1.5)])
def test_learn_matches_oracle(name, batch, n_items, lam, real):
    """Weights after several epochs of Net::learn (permutations, tail-batch rule, regularised update) and the
    predicted classes on the training set."""
    specs, cost = NETS[name]
    orc, gpu = H.make_pair(specs, real, seed=11)
    x, y = H.synthetic_batch(orc, n_items, real, seed=12, lo=0.0, hi=1.0)
    epochs, lr, seed = 3, 0.05, 5
    perms = dl.shuffle_sequence(seed, n_items, epochs)
    orc.learn(x, y, batch, epochs, lr, cost, lam, threads=1, perms=perms)
    calls = []
    gpu.reset_random_generator(seed)
    gpu.learn(x, y, batch, epochs, lr, dl.CostFunctionId(cost), lam, epoch_callback=lambda e, r: calls.append((e, r)))
    assert [c[0] for c in calls] == list(range(epochs))
    assert np.allclose([c[1] for c in calls], 0.5 * lam / n_items, rtol=1e-6)
    H.assert_close("parameters after learn", gpu.get_params(), orc.get_params(), real, rtol_mult=20)
    assert np.array_equal(gpu.predict(x), orc.eval(x)), "predicted classes differ"


WT_NETS = {
    "mini_lenet": MINI_LENET,
    # a 23-wide hidden layer: rows of W^T that are no multiple of 16 bytes, i.e. the cp.async producers instead of TMA
    "lenet_23": [ob.conv((1, 12, 12), (3, 3), 4, ob.RELU), ob.pool((4, 10, 10), (2, 2), ob.POOL_MAX),
                 ob.full((4, 5, 5), 23, ob.TANH), ob.full((1, 1, 23), 10, ob.SOFTMAX)],
    # two hidden NLayers behind the first: both take their input gradient from transposed weights
    "mlp_4": [ob.full((1, 1, 20), 15, ob.SIGMOID), ob.full((1, 1, 15), 13, ob.TANH), ob.full((1, 1, 13), 12, ob.SIGMOID),
              ob.full((1, 1, 12), 10, ob.SOFTMAX)],
}


@pytest.mark.parametrize("name,batch,n_items", [("mini_lenet", 8, 50), ("lenet_23", 6, 30), ("mlp_4", 10, 64)])
def test_input_gradient_from_transposed_weights(name, batch, n_items, monkeypatch):
    """The hidden NLayers' input-gradient contraction fed from the per-step transposed copy of the weights (K-major plan of tc_gemm.cuh;
    the library uses it from 1 024 samples per GPU up, DLB_DX_WT_MIN lowers the bar for this test): same weights after Net::learn as the
    oracle, and the transposing plan likewise."""
    specs, cost = WT_NETS[name], ob.CROSS_ENTROPY
    lam = 0.1
    results = {}
    for wt_min in ("1", "1000000"):
        monkeypatch.setenv("DLB_DX_WT_MIN", wt_min)   # read when the net is created
        orc, gpu = H.make_pair(specs, "f32", seed=21)
        x, y = H.synthetic_batch(orc, n_items, "f32", seed=22, lo=0.0, hi=1.0, one_hot=cost == ob.CROSS_ENTROPY)
        gpu.reset_random_generator(4)
        gpu.learn(x, y, batch, 2, 0.05, dl.CostFunctionId(cost), lam)
        results[wt_min] = gpu.get_params()
    orc.learn(x, y, batch, 2, 0.05, cost, lam, threads=1, perms=dl.shuffle_sequence(4, n_items, 2))
    H.assert_close("parameters (transposed-weights plan)", results["1"], orc.get_params(), "f32", rtol_mult=20)
    H.assert_close("parameters (transposing plan)", results["1000000"], orc.get_params(), "f32", rtol_mult=20)


@pytest.mark.parametrize("real,tc_mode", [("f32", 2), ("f32", 1), ("f32", 0), ("f64", 1)])
def test_lenet_step_full_size(real, tc_mode):
    """BASELINE configs[1] at its real size: one mini-batch of 256 through LeNet; gradient = sum over the batch of the
    oracle's per-sample gradients; one update.  f32 with the contractions on tcgen05 (modes 1, 2) and on the packed-FMA
    kernels (mode 0)."""
    specs = ob.lenet_specs()
    orc, gpu = H.make_pair(specs, real, seed=21, tc_mode=tc_mode)
    n = 256
    x, y = H.synthetic_batch(orc, n, real, seed=22, lo=0.0, hi=1.0)
    want = H.oracle_batch_trace(orc, x, y, ob.CROSS_ENTROPY)
    gpu.forward_backward(x, y, dl.CostFunctionId.CROSS_ENTROPY)
    got = gpu.read(dl.BUF_GRADIENT)
    for l in range(4):
        a = gpu.layer_param_offset(l); b = gpu.layer_param_offset(l + 1) if l < 3 else gpu.param_count
        if b > a:
            H.assert_close(f"lenet layer {l} gradient", got[a:b], want["grads"][a:b], real, rtol_mult=10)
    assert np.array_equal(gpu.read(dl.BUF_INDEX_MAP, 1), want["index_map"][:, 11520:11520 + 2880])
    # the same batch as one training step on host buffers, with the batch cost read back
    p0 = orc.get_params()
    cost = gpu.step_host(x, y, 0.05, dl.CostFunctionId.CROSS_ENTROPY, reg_factor=0.001)
    want_cost = sum(orc.cost(want["acts"][i, -10:], y[i]) for i in range(n))
    assert abs(cost - want_cost) <= 1e-4 * abs(want_cost)
    perm = np.arange(n, dtype=np.int64)[None, :]
    orc.learn(x, y, n, 1, 0.05, ob.CROSS_ENTROPY, 0.001 * n, threads=1, perms=perm)
    assert not np.array_equal(p0, orc.get_params())
    H.assert_close("parameters after one step", gpu.get_params(), orc.get_params(), real, rtol_mult=10)


def test_tail_batch_rule_and_step_indices():
    """n_items = 24, batch 10: batches of 10 and 14, because fewer than 1.5 * batch remaining items form one batch
    (Net.inl:235-236).  Reproduced with explicit step_indices calls."""
    specs, cost = NETS["mini_mlp"]
    real = "f64"
    orc, gpu = H.make_pair(specs, real, seed=31)
    x, y = H.synthetic_batch(orc, 24, real, seed=32)
    perms = dl.shuffle_sequence(2, 24, 1)
    orc.learn(x, y, 10, 1, 0.1, cost, 0.0, threads=1, perms=perms)
    gpu.set_dataset(x, y)
    gpu.step_indices(perms[0, :10], 0.1, dl.CostFunctionId(cost))
    gpu.step_indices(perms[0, 10:], 0.1, dl.CostFunctionId(cost))
    H.assert_close("parameters (10 + 14 split)", gpu.get_params(), orc.get_params(), real, rtol_mult=10)


def test_eval_outputs_and_classes():
    specs, _ = NETS["mini_lenet"]
    orc, gpu = H.make_pair(specs, "f32", seed=41)
    x, _ = H.synthetic_batch(orc, 33, "f32", seed=42)
    cls, outs = orc.eval(x, want_outputs=True)
    H.assert_close("net outputs", gpu.act(x), outs, "f32")
    assert np.array_equal(gpu.predict(x), cls)


def test_argument_errors_are_loud():
    with pytest.raises(dl.DlbError):
        bad = dl.Net([dl.NLayer(10, 5), dl.NLayer(7, 3)], real="f32")   # 5 != 7
        bad.param_count
    net = dl.Net([dl.NLayer(4, 3)], real="f32")
    with pytest.raises(dl.DlbError):
        net.set_params(np.zeros(5, np.float32))
    with pytest.raises(dl.DlbError):
        net.learn(np.zeros((3, 4)), np.zeros((2, 3)), 1, 1, 0.1, dl.CostFunctionId.SQUARED_ERROR)


def test_plans():
    """LeNet and the MLP take the fused path; nets outside its grammar take the generic operators."""
    assert dl.lenet("f32").describe_plan().startswith("fused: conv_pool(5x5")
    net = dl.lenet("f32"); net.enable_step_kernel(True)
    assert net.describe_plan().startswith("fused: one cooperative kernel per step")
    net64 = dl.lenet("f64"); net64.enable_step_kernel(True)      # the step kernel is float-only: nothing changes
    assert net64.describe_plan().startswith("fused: conv_pool(5x5")
    assert dl.Net([H.to_dl_spec(s) for s in MINI_MLP], real="f64").describe_plan().startswith("fused:")
    assert dl.Net([H.to_dl_spec(s) for s in MINI_LENET], real="f32").describe_plan().startswith("fused: conv_pool(3x3")
    assert dl.Net([H.to_dl_spec(s) for s in FIXTURE_NET], real="f32").describe_plan().startswith("unfused")
    assert dl.Net([H.to_dl_spec(s) for s in AVG_NET], real="f32").describe_plan().startswith("unfused")


@pytest.mark.parametrize("real", ["f32", "f64"])
def test_train_steps_equals_learn_and_graph_replay_is_exact(real):
    """train_steps (resumable) walks the same epoch sequence as learn; replaying the captured CUDA graph of the step
    gives bit-identical parameters to launching the kernels one by one."""
    specs, cost = NETS["mini_lenet"]
    x, y = None, None
    results = []
    for mode in ("learn", "steps", "steps_nograph"):
        orc, gpu = H.make_pair(specs, real, seed=51)
        if x is None:
            x, y = H.synthetic_batch(orc, 64, real, seed=52, lo=0.0, hi=1.0)
        gpu.reset_random_generator(9)
        if mode == "learn":
            gpu.learn(x, y, 8, 3, 0.05, dl.CostFunctionId(cost), 0.2)
        else:
            gpu.set_dataset(x, y)
            gpu.enable_graphs(mode == "steps")
            assert gpu.train_steps(8, 10, 0.05, dl.CostFunctionId(cost), 0.2) == 80
            assert gpu.train_steps(8, 14, 0.05, dl.CostFunctionId(cost), 0.2) == 112
        results.append(gpu.get_params())
    assert np.array_equal(results[0], results[1]), "train_steps differs from learn"
    assert np.array_equal(results[1], results[2]), "graph replay differs from eager launches"
    perms = dl.shuffle_sequence(9, 64, 3)
    orc.learn(x, y, 8, 3, 0.05, cost, 0.2, threads=1, perms=perms)
    H.assert_close("parameters", results[0], orc.get_params(), real, rtol_mult=20)


@pytest.mark.parametrize("real", ["f32", "f64"])
def test_async_host_steps_equal_sync_host_steps(real):
    """dlb_net_step_host_async (copy of step i+1 overlapped with step i, graph replay per staging half) must give the
    same parameters and the same per-step costs as the synchronous calls."""
    specs, cost = NETS["mini_lenet"]
    orc, gpu_a = H.make_pair(specs, real, seed=61)
    _, gpu_s = H.make_pair(specs, real, seed=61)
    n, steps = 16, 9
    batches = [H.synthetic_batch(orc, n, real, seed=70 + i, lo=0.0, hi=1.0) for i in range(steps)]
    hx = dl.pinned_array((steps, n, orc.in_sizes[0]), gpu_a.dtype); hy = dl.pinned_array((steps, n, orc.out_sizes[-1]), gpu_a.dtype)
    for i, (x, y) in enumerate(batches):
        hx[i] = x; hy[i] = y
    costs = dl.pinned_array((steps,), np.float64)
    costs[:] = -1
    for i in range(steps):
        gpu_a.step_host_async(hx[i], hy[i], 0.05, dl.CostFunctionId(cost), 0.01, costs[i:i + 1])
    gpu_a.ctx.synchronize()
    sync_costs = [gpu_s.step_host(hx[i], hy[i], 0.05, dl.CostFunctionId(cost), 0.01) for i in range(steps)]
    assert np.array_equal(gpu_a.get_params(), gpu_s.get_params())
    assert np.allclose(costs, sync_costs, rtol=1e-12)
    perm = np.arange(n, dtype=np.int64)[None, :]
    for x, y in batches:   # oracle: one learn call per batch with reg_factor = lambda / n
        orc.learn(x, y, n, 1, 0.05, cost, 0.01 * n, threads=1, perms=perm)
    H.assert_close("parameters after 9 host steps", gpu_a.get_params(), orc.get_params(), real, rtol_mult=20)


@pytest.mark.parametrize("real", ["f32", "f64"])
def test_cifar_deep_net_matches_oracle(real):
    """BASELINE configs[2] (3 conv+pool stages on 3x32x32, ReLU, L2) at a small batch: layer traces of the unfused
    operators and two regularised training steps (generic path: the net is outside the fused grammar)."""
    specs = ob.cifar_specs()
    orc, gpu = H.make_pair(specs, real, seed=81)
    assert gpu.describe_plan().startswith("unfused")
    x, y = H.synthetic_batch(orc, 3, real, seed=82, lo=0.0, hi=1.0)
    compare_traces(orc, gpu, x, y, ob.CROSS_ENTROPY, real)
    gpu.set_dataset(x, y)
    perm = np.array([[2, 0, 1], [1, 2, 0]], np.int64)
    orc.learn(x, y, 3, 2, 0.05, ob.CROSS_ENTROPY, 0.3, threads=1, perms=perm)
    for e in range(2):
        gpu.step_indices(perm[e], 0.05, dl.CostFunctionId.CROSS_ENTROPY, reg_factor=0.3 / 3)
    H.assert_close("cifar parameters after two steps", gpu.get_params(), orc.get_params(), real, rtol_mult=20)
    assert np.array_equal(gpu.predict(x), orc.eval(x))


def test_tensor_core_convolutions_match_fma_path_at_batch_96():
    """Implicit-GEMM convolutions on tcgen05 (tc_gemm.cuh + tc_tiles.cuh) against the packed-FMA contraction, which the
    tests above pin to the oracle: every layer buffer of the CIFAR-deep net at a batch that fills several 128-pixel tiles
    per layer and splits the filter-gradient reduction; pooling index maps exactly."""
    specs = ob.cifar_specs()
    _, fma = H.make_pair(specs, "f32", seed=91, tc_mode=0)
    orc, tcn = H.make_pair(specs, "f32", seed=91, tc_mode=2)
    x, y = H.synthetic_batch(orc, 96, "f32", seed=92, lo=0.0, hi=1.0)
    for net in (fma, tcn):
        net.forward_backward(x, y, ob.CROSS_ENTROPY, 0.0)
    for l, s in enumerate(specs):
        H.assert_close(f"layer {l} output", tcn.read(dl.BUF_OUTPUT, l), fma.read(dl.BUF_OUTPUT, l), "f32")
        if s.type == ob.PULL:
            a, b = tcn.read(dl.BUF_INDEX_MAP, l), fma.read(dl.BUF_INDEX_MAP, l)
            # an arg-max may legitimately move between two inputs that agree to rounding; the values must still agree (above)
            assert np.mean(a != b) < 1e-4, f"layer {l}: {np.sum(a != b)} pooling indices differ"
        elif s.type == ob.CONVOLUTION:
            # ReLU'(z) is a step: a pre-activation within rounding of 0 may land on either side
            da, db = tcn.read(dl.BUF_DERIVATIVES, l), fma.read(dl.BUF_DERIVATIVES, l)
            assert np.mean(da != db) < 1e-5, f"layer {l}: {np.sum(da != db)} ReLU derivatives differ"
        else:
            H.assert_close(f"layer {l} derivatives", tcn.read(dl.BUF_DERIVATIVES, l), fma.read(dl.BUF_DERIVATIVES, l), "f32", rtol_mult=5)
        if l > 0:
            H.assert_close(f"layer {l} input gradient", tcn.read(dl.BUF_INPUT_GRAD, l), fma.read(dl.BUF_INPUT_GRAD, l), "f32", rtol_mult=20)
    ga, gb = tcn.read(dl.BUF_GRADIENT), fma.read(dl.BUF_GRADIENT)
    for l in range(len(specs)):
        lo = tcn.layer_param_offset(l)
        hi = tcn.layer_param_offset(l + 1) if l + 1 < len(specs) else tcn.param_count
        if hi > lo:
            H.assert_close(f"layer {l} parameter gradient", ga[lo:hi], gb[lo:hi], "f32", rtol_mult=20)


@pytest.mark.parametrize("real", ["f32", "f64"])
def test_evaluate_cost_and_correct_answers(real):
    """Net::evaluate_cost_function_and_correct_answers (Net.inl:411-439): mean cost + l2 * sum w^2, fraction correct."""
    specs, cost = NETS["mini_lenet"]
    orc, gpu = H.make_pair(specs, real, seed=91)
    n = 50
    x, y = H.synthetic_batch(orc, n, real, seed=92, lo=0.0, hi=1.0)
    cls, outs = orc.eval(x, want_outputs=True)
    params = orc.get_params().astype(np.float64)
    sq = 0.0
    for l, s in enumerate(orc.specs):   # weights / filters only
        if s.type == ob.PULL:
            continue
        a = gpu.layer_param_offset(l) + orc.out_sizes[l]
        b = gpu.layer_param_offset(l + 1) if l + 1 < len(orc.specs) else gpu.param_count
        sq += float(np.sum(params[a:b] ** 2))
    want_cost = sum(orc.cost(outs[i], y[i], cost) for i in range(50)) / 50 + 0.01 * sq
    want_correct = float(np.mean(cls == np.argmax(y, axis=1)))
    got_cost, got_correct = gpu.evaluate_cost_function_and_correct_answers(x, y, dl.CostFunctionId(cost), 0.01)
    assert abs(got_cost - want_cost) <= (1e-4 if real == "f32" else 1e-10) * abs(want_cost)
    assert got_correct == want_correct
    assert gpu.count_correct_answers(x, y) == int(round(want_correct * 50))


DROPOUT_NETS = {"mini_lenet": (MINI_LENET, [0.8, 1.0, 0.5, 0.9], ob.CROSS_ENTROPY),
                "fixture": (FIXTURE_NET, [1.0, 0.7, 0.6, 1.0], ob.SQUARED_ERROR)}


def make_dropout_pair(name, real):
    specs, keep, cost = DROPOUT_NETS[name]
    orc = ob.OracleNet(specs, real, keep_rates=keep)
    params = H.synthetic_params(orc.param_count, specs, real, seed=5)
    orc.set_params(params)
    gpu = dl.Net([H.to_dl_spec(s) for s in specs], real=real)
    gpu.set_params(params)
    for l, k in enumerate(keep):
        gpu.set_keep_rate(l, k)
    assert gpu.mask_size == orc.mask_size > 0
    assert gpu.describe_plan().startswith("unfused")
    return orc, gpu, keep, cost


@pytest.mark.parametrize("real", ["f32", "f64"])
@pytest.mark.parametrize("name", list(DROPOUT_NETS))
def test_dropout_traces_match_oracle(name, real):
    """ALayer::ApplyDropout (ALayer.inl:46-55) in the batched operators: keep masks on layer inputs (Net.inl:126) and input
    gradients (Net.inl:274-275), every trace buffer against the oracle under the same explicit masks; inference = inputs
    times keep_rate (Net.inl:143), predicted classes exactly."""
    orc, gpu, keep, cost = make_dropout_pair(name, real)
    x, y = H.synthetic_batch(orc, 6, real, seed=21, one_hot=cost == ob.CROSS_ENTROPY)
    masks = dl.dropout_masks(9, orc.in_sizes, keep, 1, real)[0]
    orc.bind_masks(masks)
    gpu.set_dropout_masks(masks)
    compare_traces(orc, gpu, x, y, cost, real)
    orc.bind_masks(None)
    want_cls, want_out = orc.eval(x, True)
    H.assert_close("inference outputs with keep rates", gpu.act(x), want_out, real, rtol_mult=5)
    assert np.array_equal(gpu.predict(x), want_cls)


@pytest.mark.parametrize("real", ["f32", "f64"])
def test_learn_with_dropout_matches_oracle(real):
    """dlb_net_learn draws its own masks per mini-batch like Net::learn (Net.inl:228-230): after the same seed the updated
    parameters must equal the oracle's, which is fed the mask stream (tests/test_oracle_vs_reference.py pins both the stream
    and the oracle's dropout to the reference)."""
    orc, gpu, keep, cost = make_dropout_pair("mini_lenet", real)
    n, batch, epochs, seed = 60, 8, 2, 13
    x, y = H.synthetic_batch(orc, n, real, seed=22, lo=0.0, hi=1.0)
    gpu.reset_random_generator(seed)
    gpu.learn(x, y, batch, epochs, 0.05, dl.CostFunctionId.CROSS_ENTROPY, 0.1)
    masks = dl.dropout_masks(seed, orc.in_sizes, keep, ob.count_batches(n, batch) * epochs, real)
    orc.learn(x, y, batch, epochs, 0.05, cost, 0.1, 1, dl.shuffle_sequence(seed, n, epochs), masks)
    H.assert_close("parameters after learn with dropout", gpu.get_params(), orc.get_params(), real, rtol_mult=20)


# Geometries on the edge of what the tensor-core convolutions accept (tc_conv_ok, layers.cuh): a strided CLayer behind another
# layer (its input gradient stays on the FMA contraction, forward and filter gradient go to tcgen05), a 16-tap window (beyond
# the 15-bit tap masks: the whole layer stays on the FMA path), a 15-tap window with heavy padding (the largest masks).
EDGE_NETS = {
    "strided_inner": [ob.conv((1, 11, 10), (3, 3), 3, ob.TANH, pad=(1, 1)), ob.conv((3, 11, 10), (3, 2), 4, ob.SIGMOID, pad=(1, 0), stride=(2, 2)),
                      ob.full((4, 6, 5), 5, ob.LINEAR)],
    "window_16": [ob.conv((2, 18, 4), (16, 1), 3, ob.TANH, pad=(2, 0)), ob.conv((3, 7, 4), (2, 2), 2, ob.RELU), ob.full((2, 6, 3), 4, ob.LINEAR)],
    "window_15": [ob.conv((2, 9, 17), (3, 15), 3, ob.TANH, pad=(1, 7)), ob.conv((3, 9, 17), (15, 3), 2, ob.TANH, pad=(7, 1)),
                  ob.full((2, 9, 17), 4, ob.LINEAR)],
}


@pytest.mark.parametrize("name", list(EDGE_NETS))
@pytest.mark.parametrize("batch", [2, 9])
def test_tensor_core_convolution_edge_geometries(name, batch):
    specs = EDGE_NETS[name]
    orc, gpu = H.make_pair(specs, "f32", seed=31, tc_mode=2)
    x, y = H.synthetic_batch(orc, batch, "f32", seed=32, one_hot=False)
    compare_traces(orc, gpu, x, y, ob.SQUARED_ERROR, "f32")


def test_cifar_deep_full_size_steps_tensor_cores_vs_fma_and_determinism():
    """BASELINE configs[2] at its full mini-batch (1024): two regularised training steps with the contractions on tcgen05 and
    on the FMA path end at the same parameters within the float bar (the FMA path is pinned to the oracle at small sizes),
    the predicted classes of a sample agree, and a repeated run from the same state is bit-identical (every reduction is
    ordered)."""
    rng = np.random.default_rng(5)
    n = 2048
    x = rng.uniform(0, 1, (n, 3 * 32 * 32)).astype(np.float32)
    y = np.zeros((n, 10), np.float32); y[np.arange(n), rng.integers(0, 10, n)] = 1
    results = []
    for tc_mode in (0, 1, 1):
        net = dl.cifar_deep("f32")
        net.init_params(seed=3)
        net.enable_tensor_cores(tc_mode)
        net.set_dataset(x, y)
        net.reset_random_generator(9)
        net.train_steps(1024, 2, 0.01, dl.CostFunctionId.CROSS_ENTROPY, 0.1)
        results.append((net.get_params(), net.predict(x[:256])))
        net.close()
    (p_fma, c_fma), (p_tc, c_tc), (p_tc2, c_tc2) = results
    assert np.array_equal(p_tc, p_tc2) and np.array_equal(c_tc, c_tc2), "the tensor-core path is not deterministic"
    H.assert_close("parameters after two full-size steps", p_tc, p_fma, "f32", rtol_mult=20)
    assert np.mean(c_tc != c_fma) <= 0.01
