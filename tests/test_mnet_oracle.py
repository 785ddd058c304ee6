"""Pins the numpy restatement of the recurrent stack (oracle/mnet_oracle.py) to the reference: against the reference-generated
fixtures tests/golden/mnet_{f32,f64}.npz (tools/make_golden_mnet.py) and, where oracle/_ref/libdlrefm_*.so is present, against the
reference's own MNet<CpuDC> / RMLayer<CpuDC> run live (act, calc_gradient_sum, learn with and without clipping / regularisation)."""
import os

import numpy as np
import pytest

from oracle import binding as ob
from oracle.mnet_oracle import MNetOracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = {"f32": (2e-5, 2e-6), "f64": (1e-11, 1e-13)}   # (rtol, atol): the order of the per-sample sums is the reference's own


def close(name, got, want, real, mult=1.0):
    rtol, atol = TOL[real]
    scale = float(np.max(np.abs(want))) if want.size else 1.0
    assert np.all(np.isfinite(got)) and np.all(np.isfinite(want)), f"{name}: non-finite values"
    err = np.abs(got.astype(np.float64) - want.astype(np.float64))
    bad = err > mult * (atol * max(scale, 1.0) + rtol * np.abs(want))
    assert not bad.any(), f"{name}: {int(bad.sum())} of {got.size} out of tolerance, max abs err {err.max():.3e} (scale {scale:.3e})"


@pytest.mark.parametrize("real", ["f32", "f64"])
def test_oracle_matches_reference_fixtures(real):
    z = np.load(os.path.join(GOLDEN, f"mnet_{real}.npz"))
    net = MNetOracle(int(z["depth"]), z["dims"].tolist(), z["acts"].tolist(), z["clips"].tolist(), real)
    assert net.param_count == z["p0"].size
    net.set_params(z["p0"])
    close("MNet::act", net.act(z["x"]), z["out"], real)
    close("gradient sum, squared error", net.calc_gradient_sum(z["x"], z["y"], ob.SQUARED_ERROR), z["g_sq"], real, mult=4)
    close("gradient sum, cross entropy", net.calc_gradient_sum(z["x"], z["y"], ob.CROSS_ENTROPY), z["g_ce"], real, mult=4)
    net.learn(z["x"], z["y"], ob.SQUARED_ERROR, 0.3, 0.0)
    close("parameters after learn (clipped first layer)", net.get_params(), z["p1"], real, mult=4)
    net.learn(z["x"], z["y"], ob.CROSS_ENTROPY, 0.2, 0.01)
    close("parameters after the second learn (regularised)", net.get_params(), z["p2"], real, mult=8)


@pytest.mark.skipif(not ob.mnet_ref_available(), reason="oracle/_ref/libdlrefm_*.so not built")
@pytest.mark.parametrize("real", ["f32", "f64"])
@pytest.mark.parametrize("acts", [(ob.TANH,), (ob.RELU, ob.LINEAR), (ob.SIGMOID, ob.TANH, ob.SIGMOID)])
def test_oracle_matches_live_reference(real, acts):
    dt = np.float32 if real == "f32" else np.float64
    dims = [5, 9, 7, 4][:len(acts) + 1]
    clips = [0.0, 0.5, 5.0][:len(acts)]
    ref = ob.RefMNet(6, dims, list(acts), clips, real, seed=5)
    net = MNetOracle(6, dims, list(acts), clips, real)
    p = (ref.get_params() * 0.4).astype(dt)   # (unit-variance weights let ReLU / linear recurrences overflow within the sequence)
    ref.set_params(p); net.set_params(p)
    rng = np.random.default_rng(1)
    x = rng.standard_normal((11, 6, dims[0])).astype(dt); y = rng.random((11, 6, dims[-1])).astype(dt)
    close("act", net.act(x), ref.act(x), real)
    close("gradient sum", net.calc_gradient_sum(x, y, ob.SQUARED_ERROR), ref.calc_gradient_sum(x, y, ob.SQUARED_ERROR), real, mult=4)
    for step in range(3):
        ref.learn(x, y, ob.SQUARED_ERROR, 0.25, 0.02 * step); net.learn(x, y, ob.SQUARED_ERROR, 0.25, 0.02 * step)
    close("parameters after three learn steps", net.get_params(), ref.get_params(), real, mult=16)
