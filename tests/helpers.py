"""Shared helpers of the parity tests: build the same net in the CPU oracle and in libdlb200, compare buffers."""
from __future__ import annotations

import numpy as np

import deeplearning_b200 as dl
from oracle import binding as ob

# Relative tolerances of the parity bar (BASELINE.json north_star: "about 1e-5 for float, tighter for double").
RTOL = {"f32": 1e-5, "f64": 1e-12}


def eps(real: str) -> float:
    return float(np.finfo(ob.dtype_of(real)).eps)


def to_dl_spec(s: ob.LayerSpec) -> dl.LayerSpec:
    return dl.LayerSpec(*[getattr(s, f) for f, _ in ob.LayerSpec._fields_])


def synthetic_params(n: int, specs, real: str, seed: int = 0) -> np.ndarray:
    """Parameters with the reference's distributions (biases U(-1,1), weights N(0,1/sqrt(fan_in))) from numpy;
    used where the reference library (oracle/_ref) is not needed to draw them."""
    rng = np.random.default_rng(seed)
    flat = np.zeros(n, ob.dtype_of(real))
    off = 0
    for s in specs:
        if s.type == ob.PULL:
            continue
        o = ob.out_shape(s)
        if s.type == ob.FULL:
            nb, fan = s.out_c, s.in_c * s.in_h * s.in_w
            nw = nb * fan
        else:
            nb, fan = o[0] * o[1] * o[2], s.in_c * s.win_h * s.win_w
            nw = s.out_c * fan
        flat[off:off + nb] = rng.uniform(-1, 1, nb); off += nb
        flat[off:off + nw] = rng.normal(0, 1 / np.sqrt(fan), nw); off += nw
    assert off == n
    return flat


def make_pair(specs, real: str, seed: int = 0, tc_mode: int = 2):
    """(oracle net, gpu net) with identical parameters.  tc_mode: 0 = packed-FMA contractions, 1 = the library default
    (tcgen05 from 128 samples per batch), 2 = tcgen05 at every batch size (the default here, so that the small parity
    cases exercise the tensor-core kernels)."""
    orc = ob.OracleNet(specs, real)
    params = synthetic_params(orc.param_count, specs, real, seed)
    orc.set_params(params)
    gpu = dl.Net([to_dl_spec(s) for s in specs], real=real)
    assert gpu.param_count == orc.param_count
    gpu.set_params(params)
    if real == "f32":
        gpu.enable_tensor_cores(tc_mode)
    return orc, gpu


def synthetic_batch(orc, n: int, real: str, seed: int = 1, lo=-1.0, hi=1.0, one_hot=True):
    rng = np.random.default_rng(seed)
    x = rng.uniform(lo, hi, (n, orc.in_sizes[0])).astype(ob.dtype_of(real))
    k = orc.out_sizes[-1]
    if one_hot:
        y = np.zeros((n, k), ob.dtype_of(real))
        y[np.arange(n), rng.integers(0, k, n)] = 1
    else:
        y = rng.uniform(0.05, 0.95, (n, k)).astype(ob.dtype_of(real))
    return x, y


def oracle_batch_trace(orc, x, y, cost_id, grads_in=None, scale=0.0):
    """Runs the oracle sample by sample the way Net::learn does inside one thread (first sample scale, then 1)
    and stacks the per-sample traces."""
    n = x.shape[0]
    acts, derivs, imaps, in_grads, cost_grads = [], [], [], [], []
    grads = None if grads_in is None else np.array(grads_in, orc.dtype)
    for i in range(n):
        s = scale if i == 0 else 1.0
        t = orc.trace(x[i], y[i], cost_id, s, grads if (i > 0 or scale != 0.0) else None)
        grads = t["grads"]
        acts.append(t["acts"]); derivs.append(t["derivs"]); imaps.append(t["index_map"])
        in_grads.append(t["in_grads"]); cost_grads.append(t["cost_grad"])
    return dict(acts=np.stack(acts), derivs=np.stack(derivs), index_map=np.stack(imaps), in_grads=np.stack(in_grads),
                cost_grad=np.stack(cost_grads), grads=grads)


def assert_close(name, got, want, real, rtol_mult=1.0, atol_scale=None):
    """|got - want| <= atol + rtol * |want| with rtol = RTOL[real] * rtol_mult and
    atol = rtol * max|want| (summation-order noise scales with the largest term)."""
    got = np.asarray(got, np.float64); want = np.asarray(want, np.float64)
    assert got.shape == want.shape, f"{name}: shape {got.shape} != {want.shape}"
    rtol = RTOL[real] * rtol_mult
    scale = atol_scale if atol_scale is not None else (np.max(np.abs(want)) if want.size else 0.0)
    atol = rtol * max(scale, 1e-30)
    err = np.abs(got - want)
    bad = err > atol + rtol * np.abs(want)
    if np.any(bad) or not np.all(np.isfinite(got)):
        i = int(np.argmax(err))
        raise AssertionError(f"{name}: {int(bad.sum())} of {got.size} elements out of tolerance; max abs err {err.max():.3e} "
                             f"at flat index {i} (got {got.ravel()[i]!r}, want {want.ravel()[i]!r}), rtol {rtol:g}, atol {atol:.3e}")


def layer_offsets(sizes):
    off, out = 0, []
    for s in sizes:
        out.append(off); off += s
    return out
