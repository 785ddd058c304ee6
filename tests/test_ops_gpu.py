"""GPU parity of the op-level C ABI (dlb_conv_forward_* ... dlb_sgd_update_*) on raw device pointers, restating the
reference's GPU-vs-CPU layer tests (DeepLearningCudaTest/NetLayersCudaTest.cpp:77-172) with its fixtures
(Shared/NetLayersTestUtils.h:34-65): CLayer {5,13,17} 3x3x7 pad {0,3,6} stride 2, PLayer {5,10,7} window 3x4,
NLayer 10 -> 23; sigmoid / softmax / relu / tanh; max / min / average.  Index maps must be exactly equal."""
import ctypes as C

import numpy as np
import pytest

import deeplearning_b200 as dl
from deeplearning_b200 import _capi
from oracle import binding as ob
from tests import helpers as H

pytestmark = pytest.mark.gpu


class Dev:
    """Tiny RAII helper around dlb_malloc / dlb_memcpy for the tests."""

    def __init__(self):
        self.ctx = dl.Context(0)
        self.lib = self.ctx.lib
        self.ptrs = []

    def up(self, a):
        a = np.ascontiguousarray(a)
        p = C.c_void_p()
        _capi.check(self.lib.dlb_malloc(self.ctx.h, max(a.nbytes, 1), C.byref(p)))
        _capi.check(self.lib.dlb_memcpy_h2d(self.ctx.h, p, a.ctypes.data, a.nbytes))
        self.ctx.synchronize()
        self.ptrs.append(p)
        return p

    def empty(self, shape, dtype):
        return self.up(np.zeros(shape, dtype))

    def down(self, p, shape, dtype):
        a = np.empty(shape, dtype)
        _capi.check(self.lib.dlb_memcpy_d2h(self.ctx.h, a.ctypes.data, p, a.nbytes))
        return a

    def close(self):
        for p in self.ptrs:
            self.lib.dlb_free(self.ctx.h, p)


def fn(dev, name, real):
    return getattr(dev.lib, f"dlb_{name}_{real}")


LAYERS = {
    "clayer_sigmoid": ob.conv((5, 13, 17), (3, 3), 7, ob.SIGMOID, pad=(3, 6), stride=(2, 2)),
    "clayer_softmax": ob.conv((5, 13, 17), (3, 3), 7, ob.SOFTMAX, pad=(3, 6), stride=(2, 2)),
    "clayer_relu_s1": ob.conv((3, 9, 11), (5, 5), 6, ob.RELU, pad=(2, 2)),
    "nlayer_sigmoid": ob.full((1, 1, 10), 23, ob.SIGMOID),
    "nlayer_softmax": ob.full((1, 1, 10), 23, ob.SOFTMAX),
    "nlayer_tanh": ob.full((1, 1, 10), 23, ob.TANH),
}


@pytest.mark.parametrize("real", ["f32", "f64"])
@pytest.mark.parametrize("name", list(LAYERS))
@pytest.mark.parametrize("use_gather", [False, True])
def test_parametric_layer_forward_backward(name, real, use_gather):
    spec = LAYERS[name]
    orc = ob.OracleNet([spec], real)
    dt = orc.dtype
    rng = np.random.default_rng(3)
    params = H.synthetic_params(orc.param_count, [spec], real, seed=7)
    orc.set_params(params)
    B, rows = 5, 9
    in_n, out_n = orc.in_sizes[0], orc.out_sizes[0]
    store = rng.uniform(-1, 1, (rows, in_n)).astype(dt)
    gather = np.array([7, 2, 2, 0, 8], np.int32)
    x = store[gather] if use_gather else store[:B]
    delta = rng.uniform(-1, 1, (B, out_n)).astype(dt)
    g_in = rng.normal(size=orc.param_count).astype(dt)
    scale = 0.37
    bias_n = out_n
    dev = Dev()
    try:
        dspec = H.to_dl_spec(spec)
        d_store, d_gather = dev.up(store), dev.up(gather)
        d_x = d_store if use_gather else dev.up(x)
        d_g = d_gather if use_gather else None
        d_bias, d_w = dev.up(params[:bias_n]), dev.up(params[bias_n:])
        d_y, d_deriv = dev.empty((B, out_n), dt), dev.empty((B, out_n), dt)
        kind = "conv" if spec.type == ob.CONVOLUTION else "full"
        _capi.check(fn(dev, kind + "_forward", real)(dev.ctx.h, C.byref(dspec), B, d_x, d_g, d_bias, d_w, d_y, d_deriv))
        y, deriv = dev.down(d_y, (B, out_n), dt), dev.down(d_deriv, (B, out_n), dt)
        # oracle: forward via eval-style trace of a one-layer net; backward via a hand-made delta is not part of the
        # net API, so the oracle's per-sample trace is used with the LINEAR cost whose derivative is 1 and the deltas
        # are applied by linearity: run the oracle once per sample with cost = LINEAR and scale the comparison instead.
        want_y = np.stack([orc.trace(x[i], np.zeros(out_n, dt), ob.COST_LINEAR)["acts"] for i in range(B)])
        want_d = np.stack([orc.trace(x[i], np.zeros(out_n, dt), ob.COST_LINEAR)["derivs"] for i in range(B)])
        H.assert_close("output", y, want_y, real)
        H.assert_close("derivatives", deriv, want_d, real, rtol_mult=5)
        # backward with delta == 1 (the LINEAR cost's derivative), gradient_scale_factor on a random previous gradient
        ones = np.ones((B, out_n), dt)
        d_delta = dev.up(ones)
        d_gs, d_db, d_dw = dev.empty((B, out_n), dt), dev.up(g_in[:bias_n]), dev.up(g_in[bias_n:])
        d_dx = dev.empty((B, in_n), dt)
        _capi.check(fn(dev, kind + "_backward", real)(dev.ctx.h, C.byref(dspec), B, d_x, d_g, d_delta, d_deriv, d_w, d_gs, d_db, d_dw, d_dx, scale))
        got = np.concatenate([dev.down(d_db, (bias_n,), dt), dev.down(d_dw, (orc.param_count - bias_n,), dt)])
        want = g_in.copy()
        for i in range(B):
            want = orc.trace(x[i], np.zeros(out_n, dt), ob.COST_LINEAR, scale if i == 0 else 1.0, want)["grads"]
        H.assert_close("parameter gradient (scaled accumulation)", got, want, real, rtol_mult=5)
        # input gradient against a finite check through the oracle of a two-layer net is covered in test_net_gpu;
        # here: linearity in delta (dx(2*delta) == 2*dx(delta)) and zero delta -> zero dx
        dx1 = dev.down(d_dx, (B, in_n), dt)
        d_delta2 = dev.up(2 * ones)
        _capi.check(fn(dev, kind + "_backward", real)(dev.ctx.h, C.byref(dspec), B, d_x, d_g, d_delta2, d_deriv, d_w, d_gs, d_db, d_dw, d_dx, 0.0))
        H.assert_close("dx linearity", dev.down(d_dx, (B, in_n), dt), 2 * dx1, real, rtol_mult=2)
    finally:
        dev.close()


@pytest.mark.parametrize("real", ["f32", "f64"])
@pytest.mark.parametrize("pool_type", [ob.POOL_MAX, ob.POOL_MIN, ob.POOL_AVERAGE])
@pytest.mark.parametrize("in_size,window", [((5, 10, 7), (3, 4)),     # generic kernel: trailing row / columns dropped
                                            ((3, 8, 12), (2, 2)),     # vectorised 2 x 2 kernels (even height, width a multiple of 4)
                                            ((6, 4, 4), (2, 2)),      # ... one thread per output row
                                            ((2, 6, 6), (2, 2)),      # width not a multiple of 4: generic kernel
                                            ((3, 5, 8), (2, 2))])     # odd height: generic kernel, last row uncovered
def test_player_forward_backward(in_size, window, pool_type, real):
    """PLayer forward / backward, every kernel variant (layers.cuh); small integer inputs give many exact ties inside the windows, which
    must resolve to the first element in (ky, kx) scan order: outputs, index maps and input gradients exactly equal to the oracle's."""
    spec = ob.pool(in_size, window, pool_type)
    Cn, n_in = in_size[0], in_size[0] * in_size[1] * in_size[2]
    # a conv front-end only to give the oracle net a first layer whose input gradient is evaluated for the pool
    net_specs = [ob.conv(in_size, (1, 1), Cn, ob.LINEAR), spec]
    orc = ob.OracleNet(net_specs, real)
    dt = orc.dtype
    p = np.zeros(orc.param_count, dt)
    p[n_in:] = np.eye(Cn, dtype=dt).ravel()            # identity 1x1 convolution, zero bias: the pool sees the input itself
    orc.set_params(p)
    rng = np.random.default_rng(4)
    B = 5
    x = rng.integers(-3, 4, (B, n_in)).astype(dt)      # small integers: many exact ties inside the windows
    out_n = orc.out_sizes[1]
    dev = Dev()
    try:
        dspec = H.to_dl_spec(spec)
        d_x, d_y, d_map = dev.up(x), dev.empty((B, out_n), dt), dev.empty((B, out_n), np.int32)
        _capi.check(fn(dev, "pool_forward", real)(dev.ctx.h, C.byref(dspec), B, d_x, d_y, d_map))
        traces = [orc.trace(x[i], np.zeros(out_n, dt), ob.COST_LINEAR) for i in range(B)]
        want_y = np.stack([t["acts"][n_in:] for t in traces])
        assert np.array_equal(dev.down(d_y, (B, out_n), dt), want_y), "pool output must be exactly equal"
        if pool_type != ob.POOL_AVERAGE:
            want_map = np.stack([t["index_map"][n_in:] for t in traces])
            assert np.array_equal(dev.down(d_map, (B, out_n), np.int32), want_map), "index map must be exactly equal"
        d_delta, d_dx = dev.up(np.ones((B, out_n), dt)), dev.empty((B, n_in), dt)
        _capi.check(fn(dev, "pool_backward", real)(dev.ctx.h, C.byref(dspec), B, d_delta, d_map, d_dx))
        want_dx = np.stack([t["in_grads"][n_in:] for t in traces])
        assert np.array_equal(dev.down(d_dx, (B, n_in), dt), want_dx), "pool input gradient must be exactly equal"
    finally:
        dev.close()


@pytest.mark.parametrize("real", ["f32", "f64"])
def test_cost_gradient_and_update(real):
    dt = ob.dtype_of(real)
    dev = Dev()
    try:
        outs = np.array([[0.0, 1.0, 0.3, 1.0, 0.0, 0.7], [0.2, 0.9, 0.5, 0.1, 0.6, 0.4]], dt)
        refs = np.array([[0.0, 1.0, 0.4, 0.0, 1.0, 1.0], [0.0, 1.0, 0.0, 0.3, 1.0, 0.0]], dt)
        orc = ob.OracleNet([ob.full((1, 1, 3), 6, ob.LINEAR)], real)
        for cost in (ob.SQUARED_ERROR, ob.CROSS_ENTROPY, ob.COST_LINEAR):
            d_o, d_r, d_g = dev.up(outs), dev.up(refs), dev.empty((2, 6), dt)
            _capi.check(fn(dev, "cost_gradient", real)(dev.ctx.h, cost, 2, 6, d_o, d_r, None, d_g))
            got = dev.down(d_g, (2, 6), dt)
            for i in range(2):
                p = np.zeros(orc.param_count, dt); p[:6] = outs[i]
                orc.set_params(p)
                want = orc.trace(np.zeros(3, dt), refs[i], cost)["cost_grad"]
                assert np.allclose(got[i], want, rtol=1e-6, atol=0), (cost, got[i], want)
        # ALayer::update on one container: w <- w (1 + reg*l) + g * l * gsf  /  b <- b + g * l * gsf
        rng = np.random.default_rng(5)
        w, g = rng.normal(size=1000).astype(dt), rng.normal(size=1000).astype(dt)
        for reg_flag, reg in ((1, 0.01), (0, 0.01), (1, 0.0)):
            d_w, d_g = dev.up(w), dev.up(g)
            _capi.check(fn(dev, "sgd_update", real)(dev.ctx.h, d_w, d_g, 1000, reg_flag, -0.1, reg, 1 / 8))
            l = dt(-0.1)
            want = w * (dt(1) + dt(reg) * l) + g * (l * dt(1 / 8)) if (reg_flag and reg > 0) else w + g * (l * dt(1 / 8))
            H.assert_close("updated container", dev.down(d_w, (1000,), dt), want, real)
    finally:
        dev.close()


def test_op_argument_errors():
    dev = Dev()
    try:
        spec = H.to_dl_spec(ob.full((1, 1, 10), 23, ob.SIGMOID))
        with pytest.raises(dl.DlbError):
            _capi.check(dev.lib.dlb_conv_forward_f32(dev.ctx.h, C.byref(spec), 1, None, None, None, None, None, None))
        bad = H.to_dl_spec(ob.conv((1, 4, 4), (3, 3), 2, ob.RELU)); bad.stride_h = 0
        d = dev.empty((64,), np.float32)
        with pytest.raises(dl.DlbError):
            _capi.check(dev.lib.dlb_conv_forward_f32(dev.ctx.h, C.byref(bad), 1, d, None, d, d, d, d))
    finally:
        dev.close()
