"""GPU parity in ONE hop: the CUDA path (through the C ABI) against what the reference itself computed.

  * tests/golden/*.npz  - outputs of the unmodified reference (oracle/_ref, built from /root/reference by
    tools/make_golden.py in the build container); inputs and initial parameters are regenerated from numpy seeds.
  * oracle/_ref live    - the prebuilt reference libraries travel to the GPU box with the snapshot; Net<CpuDC>::learn with
    all host threads is fast enough to check full-size workloads (LeNet B=256, CIFAR-deep B=1024) in seconds.  Where the
    libraries are absent the C oracle (pinned to the reference by tests/test_golden_oracle.py) stands in.

The pattern is the reference's own GPU-vs-CPU comparison, DeepLearningCudaTest/NetLayersCudaTest.cpp:77-172.
"""
import numpy as np
import pytest

import deeplearning_b200 as dl
from oracle import binding as ob
from tests import golden_cases as gc
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _gpu_net(case, real, tc_mode=2):
    specs = case["specs"]
    gpu = dl.Net([H.to_dl_spec(s) for s in specs], real=real)
    gpu.set_params(H.synthetic_params(gpu.param_count, specs, real, seed=case["param_seed"]))
    if real == "f32":
        gpu.enable_tensor_cores(tc_mode)
    in_n = gpu.layer_in_size(0)
    out_n = gpu.layer_out_size(len(specs) - 1)
    x, y = gc.case_data(case, in_n, out_n, real)
    return gpu, x, y, gc.load(case, real)


@pytest.mark.parametrize("real", ["f32", "f64"])
@pytest.mark.parametrize("case", gc.CASES, ids=[c["name"] for c in gc.CASES])
def test_traces_match_reference_fixture(case, real):
    """Every trace buffer of the first n_trace samples against the reference's fixture (index maps exactly)."""
    gpu, x, y, gold = _gpu_net(case, real)
    nt = case["n_trace"]
    specs = case["specs"]
    gpu.forward_backward(x[:nt], y[:nt], case["cost"], 0.0)
    L = len(specs)
    out_sizes = [gpu.layer_out_size(l) for l in range(L)]
    in_sizes = [gpu.layer_in_size(l) for l in range(L)]
    out_off, in_off = H.layer_offsets(out_sizes), H.layer_offsets(in_sizes)
    imap_gold = gold["trace_index_map"]
    for l, s in enumerate(specs):
        sl = slice(out_off[l], out_off[l] + out_sizes[l])
        if s.type == ob.PULL and s.pool_type != ob.POOL_AVERAGE:
            assert np.array_equal(gpu.read(dl.BUF_INDEX_MAP, l), imap_gold[:, sl]), f"layer {l}: pooling index map differs from the reference's"
        if not case["store_traces"]:
            continue
        H.assert_close(f"layer {l} output", gpu.read(dl.BUF_OUTPUT, l), gold["trace_acts"][:, sl], real)
        if s.type != ob.PULL:
            H.assert_close(f"layer {l} derivatives", gpu.read(dl.BUF_DERIVATIVES, l), gold["trace_derivs"][:, sl], real, rtol_mult=5)
        if l > 0:
            isl = slice(in_off[l], in_off[l] + in_sizes[l])
            H.assert_close(f"layer {l} input gradient", gpu.read(dl.BUF_INPUT_GRAD, l), gold["trace_in_grads"][:, isl], real, rtol_mult=5)
    got_g = gpu.read(dl.BUF_GRADIENT)
    if case["store_traces"]:
        H.assert_close("cost gradient", gpu.read(dl.BUF_COST_GRAD), gold["trace_cost_grad"], real, rtol_mult=5)
        H.assert_close("gradient", got_g, gold["trace_grads"], real, rtol_mult=5)
    else:
        H.assert_close("gradient (strided)", got_g[::gc.STRIDE], gold["trace_grads_strided"], real, rtol_mult=10)
        H.assert_close("net output", gpu.read(dl.BUF_OUTPUT, L - 1), gold["trace_out"], real)


@pytest.mark.parametrize("real", ["f32", "f64"])
@pytest.mark.parametrize("case", gc.CASES, ids=[c["name"] for c in gc.CASES])
def test_learn_matches_reference_fixture(case, real):
    """dlb_net_learn from the fixture's seed: parameters after learn, predicted classes (exactly) and outputs."""
    gpu, x, y, gold = _gpu_net(case, real, tc_mode=1)
    ln = case["learn"]
    assert np.array_equal(dl.shuffle_sequence(ln["seed"], x.shape[0], ln["epochs"]), gold["learn_perms"])
    gpu.reset_random_generator(ln["seed"])
    gpu.learn(x, y, ln["batch"], ln["epochs"], ln["lr"], dl.CostFunctionId(case["cost"]), ln["lam"])
    after = gpu.get_params()
    H.assert_close("parameters after learn", after if case["store_traces"] else after[::gc.STRIDE], gold["learn_params"], real, rtol_mult=20)
    s = np.array([after.astype(np.float64).sum(), (after.astype(np.float64) ** 2).sum()])
    assert np.allclose(s, gold["learn_param_sum"], rtol=1e-5 if real == "f32" else 1e-11)
    assert np.array_equal(gpu.predict(x), gold["learn_classes"]), "predicted classes differ from the reference's"
    H.assert_close("outputs after learn", gpu.act(x[:16]), gold["learn_outputs"], real, rtol_mult=20)


def _reference_or_oracle(specs, real, params):
    """The live reference (all host threads) when its prebuilt library is on this box, else the C oracle."""
    if ob.ref_available():
        net = ob.RefNet(specs, real, seed=0)
        net.set_params(params)
        return net, "reference"
    net = ob.OracleNet(specs, real)
    net.set_params(params)
    return net, "oracle"


def _cpu_learn(net, kind, x, y, batch, epochs, lr, cost, lam, seed):
    if kind == "reference":
        net.learn(x, y, batch, epochs, lr, cost, lam, parallelism=-1, shuffle_seed=seed)
    else:
        net.learn(x, y, batch, epochs, lr, cost, lam, threads=1, perms=dl.shuffle_sequence(seed, x.shape[0], epochs))


@pytest.mark.parametrize("real,mode", [("f32", "steps"), ("f32", "learn"), ("f64", "steps")])
def test_lenet_b256_train_mode_vs_reference(real, mode):
    """BASELINE configs[1] at full size through the path the bench times: train_steps at mini-batch 256 (eager first step,
    graph capture, graph replays) against Net<CpuDC>::learn of the reference on the same epoch permutations."""
    specs = ob.lenet_specs()
    n, batch, lr, lam, seed = 1280, 256, 0.05, 0.3, 4
    params = H.synthetic_params(301130, specs, real, seed=7)
    rng = np.random.default_rng(8)
    x = rng.random((n, 784)).astype(ob.dtype_of(real)); y = ob.one_hot_labels(real, n)
    cpu, kind = _reference_or_oracle(specs, real, params)
    _cpu_learn(cpu, kind, x, y, batch, 1, lr, ob.CROSS_ENTROPY, lam, seed)
    gpu = dl.Net([H.to_dl_spec(s) for s in specs], real=real)
    gpu.set_params(params)
    gpu.reset_random_generator(seed)
    if mode == "steps":
        gpu.set_dataset(x, y)
        assert gpu.train_steps(batch, 5, lr, dl.CostFunctionId.CROSS_ENTROPY, lam) == n
    else:
        gpu.learn(x, y, batch, 1, lr, dl.CostFunctionId.CROSS_ENTROPY, lam)
    H.assert_close(f"LeNet parameters after 5 steps of 256 vs the {kind}", gpu.get_params(), cpu.get_params(), real, rtol_mult=20)
    assert np.array_equal(gpu.predict(x[:512]), cpu.eval(x[:512])), "predicted classes differ"


def test_cifar_deep_b1024_step_vs_reference():
    """BASELINE configs[2] at its full mini-batch: one regularised step of 1024 samples with every contraction on tcgen05
    (split filter-gradient reductions, multi-tile paths) against the reference's Net<CpuDC>::learn; predicted classes of
    256 samples exactly."""
    specs = ob.cifar_specs()
    real = "f32"
    n, lr, lam, seed = 1024, 0.01, 0.1, 2
    rng = np.random.default_rng(5)
    x = rng.uniform(0, 1, (n, 3 * 32 * 32)).astype(np.float32)
    y = np.zeros((n, 10), np.float32); y[np.arange(n), rng.integers(0, 10, n)] = 1
    gpu = dl.cifar_deep(real)
    params = H.synthetic_params(gpu.param_count, specs, real, seed=3)
    gpu.set_params(params)
    gpu.enable_tensor_cores(1)
    cpu, kind = _reference_or_oracle(specs, real, params)
    _cpu_learn(cpu, kind, x, y, n, 1, lr, ob.CROSS_ENTROPY, lam, seed)
    gpu.reset_random_generator(seed)
    gpu.learn(x, y, n, 1, lr, dl.CostFunctionId.CROSS_ENTROPY, lam)
    got, want = gpu.get_params(), cpu.get_params()
    for l in range(len(specs)):   # per layer, so that each container is judged against its own scale
        a = gpu.layer_param_offset(l)
        b = gpu.layer_param_offset(l + 1) if l + 1 < len(specs) else gpu.param_count
        if b > a:
            H.assert_close(f"CIFAR-deep layer {l} parameters after one step of 1024 vs the {kind}", got[a:b], want[a:b], real, rtol_mult=20)
    # the update itself must be resolved, not just the unchanged bulk of the parameters
    H.assert_close("CIFAR-deep parameter update", got - params, want - params, real, rtol_mult=2000)
    assert np.array_equal(gpu.predict(x[:256]), cpu.eval(x[:256])), "predicted classes differ"
