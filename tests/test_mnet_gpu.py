"""GPU parity of the recurrent stack (csrc/mnet.cu: MNet of RMLayer batched over the mini-batch) through the C
ABI (dlb_mnet_*): against the reference-generated fixtures tests/golden/mnet_*.npz, against the numpy oracle on ragged sizes, and -
where oracle/_ref/libdlrefm_*.so is on the box - against the reference's own MNet<CpuDC>::learn at BASELINE configs[3]'s shape
(length 64, hidden 256) on a bounded mini-batch.  Float: ~1e-5 relative; double: ~1e-12 (the sums over the mini-batch run in the
contraction's tile order instead of the reference's sample order)."""
import os

import numpy as np
import pytest

import deeplearning_b200 as dl
from oracle import binding as ob
from oracle.mnet_oracle import MNetOracle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = {"f32": (4e-5, 4e-6), "f64": (1e-10, 1e-12)}


def close(name, got, want, real, mult=1.0):
    rtol, atol = TOL[real]
    assert np.all(np.isfinite(got)) and np.all(np.isfinite(want)), f"{name}: non-finite values"
    scale = float(np.max(np.abs(want))) if want.size else 1.0
    err = np.abs(got.astype(np.float64) - want.astype(np.float64))
    bad = err > mult * (atol * max(scale, 1.0) + rtol * np.abs(want))
    assert not bad.any(), f"{name}: {int(bad.sum())} of {got.size} out of tolerance, max abs err {err.max():.3e} (scale {scale:.3e})"


def gpu_net(depth, dims, acts, clips, real):
    return dl.MNet(depth, [dl.RMLayer(dims[l], dims[l + 1], dl.ActivationFunctionId(acts[l]), clips[l]) for l in range(len(acts))], real=real)


@pytest.mark.parametrize("real", ["f32", "f64"])
def test_reference_fixtures(real):
    z = np.load(os.path.join(GOLDEN, f"mnet_{real}.npz"))
    net = gpu_net(int(z["depth"]), z["dims"].tolist(), z["acts"].tolist(), z["clips"].tolist(), real)
    assert net.param_count == z["p0"].size
    net.set_params(z["p0"])
    close("MNet::act", net.act(z["x"]), z["out"], real)
    close("gradient sum, squared error", net.calc_gradient_sum(z["x"], z["y"], dl.CostFunctionId.SQUARED_ERROR), z["g_sq"], real, mult=4)
    close("gradient sum, cross entropy", net.calc_gradient_sum(z["x"], z["y"], dl.CostFunctionId.CROSS_ENTROPY), z["g_ce"], real, mult=4)
    net.learn(z["x"], z["y"], dl.CostFunctionId.SQUARED_ERROR, 0.3, 0.0)
    close("parameters after learn (clipped first layer)", net.get_params(), z["p1"], real, mult=4)
    net.learn(z["x"], z["y"], dl.CostFunctionId.CROSS_ENTROPY, 0.2, 0.01)
    close("parameters after the second learn (regularised)", net.get_params(), z["p2"], real, mult=8)


@pytest.mark.parametrize("real", ["f32", "f64"])
@pytest.mark.parametrize("depth,dims,acts,n", [(1, [3, 5], (ob.TANH,), 4), (6, [5, 9, 7, 4], (ob.SIGMOID, ob.TANH, ob.SIGMOID), 11),
                                                (13, [17, 70, 33], (ob.RELU, ob.LINEAR), 37), (5, [40, 130], (ob.TANH,), 150),
                                                (4, [20, 256, 8], (ob.TANH, ob.SIGMOID), 300)])   # (the last one: single-launch recurrence contractions)
def test_against_the_oracle(real, depth, dims, acts, n):
    """Ragged sizes (tiles of the contractions partly filled, split reductions), depth 1 (no recurrence), every element-wise activation."""
    dt = np.float32 if real == "f32" else np.float64
    clips = [0.0, 0.5, 5.0][:len(acts)]
    orc = MNetOracle(depth, dims, list(acts), clips, real)
    net = gpu_net(depth, dims, list(acts), clips, real)
    rng = np.random.default_rng(7)
    p = (rng.standard_normal(orc.param_count) * 0.25 / np.sqrt(max(dims) / 8)).astype(dt)
    orc.set_params(p); net.set_params(p)
    x = rng.standard_normal((n, depth, dims[0])).astype(dt); y = rng.random((n, depth, dims[-1])).astype(dt)
    close("act", net.act(x), orc.act(x), real)
    close("gradient sum", net.calc_gradient_sum(x, y, dl.CostFunctionId.SQUARED_ERROR), orc.calc_gradient_sum(x, y, ob.SQUARED_ERROR), real, mult=4)
    for step in range(3):
        orc.learn(x, y, ob.SQUARED_ERROR, 0.25, 0.02 * step); net.learn(x, y, dl.CostFunctionId.SQUARED_ERROR, 0.25, 0.02 * step)
    close("parameters after three learn steps", net.get_params(), orc.get_params(), real, mult=16)


def test_resident_steps_equal_host_steps_and_are_deterministic():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((64, 16, 24)).astype(np.float32); y = rng.random((64, 16, 12)).astype(np.float32)
    results = []
    for mode in ("host", "resident", "resident"):
        net = gpu_net(16, [24, 48, 12], [ob.TANH, ob.SIGMOID], [1.0, 5.0], "f32")
        net.init_params(seed=3)
        if mode == "host":
            for _ in range(4): net.learn(x, y, dl.CostFunctionId.SQUARED_ERROR, 0.1, 0.001)
        else:
            net.set_batch(x, y); net.learn_resident(dl.CostFunctionId.SQUARED_ERROR, 0.1, 0.001, n_steps=4)
        results.append(net.get_params())
    assert np.array_equal(results[0], results[1]) and np.array_equal(results[1], results[2])


def test_argument_errors():
    net = gpu_net(4, [3, 5], [ob.TANH], [5.0], "f32")
    with pytest.raises(ValueError): net.act(np.zeros((2, 5, 3), np.float32))          # wrong depth ("Invalid input data", RMLayer.inl:171-172)
    with pytest.raises(dl.DlbError): dl.MNet(4, [dl.RMLayer(3, 5), dl.RMLayer(6, 2)])  # "Incompatible input size of the new layer." (MNet.inl:300-301)
    with pytest.raises(dl.DlbError): dl.MNet(4, [dl.RMLayer(3, 5, dl.ActivationFunctionId.SOFTMAX)])


@pytest.mark.skipif(not ob.mnet_ref_available(), reason="oracle/_ref/libdlrefm_*.so not on this box")
@pytest.mark.parametrize("real", ["f32", "f64"])
def test_baseline_shape_vs_live_reference(real):
    """BASELINE configs[3]'s shape - sequences of length 64, hidden size 256 - on 24 sequences (the reference needs ~0.2 s per sequence):
    two learn steps against the reference's own MNet<CpuDC>::learn."""
    dt = np.float32 if real == "f32" else np.float64
    depth, dims, acts, clips = 64, [32, 256, 16], [ob.TANH, ob.TANH], [5.0, 5.0]
    ref = ob.RefMNet(depth, dims, acts, clips, real, seed=1)
    p = (ref.get_params() * 0.08).astype(dt)
    ref.set_params(p)
    net = gpu_net(depth, dims, acts, clips, real); net.set_params(p)
    rng = np.random.default_rng(4)
    n = 24
    x = rng.standard_normal((n, depth, dims[0])).astype(dt); y = (rng.random((n, depth, dims[-1])) * 2 - 1).astype(dt)
    close("act", net.act(x), ref.act(x), real, mult=4)
    for _ in range(2):
        ref.learn(x, y, ob.SQUARED_ERROR, 0.05, 0.001); net.learn(x, y, dl.CostFunctionId.SQUARED_ERROR, 0.05, 0.001)
    close("parameters after two learn steps vs the reference", net.get_params(), ref.get_params(), real, mult=16)
