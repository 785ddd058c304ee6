"""ctypes bindings for the CPU checkers under oracle/ (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may import this module.
The product package (deeplearning_b200) never does.

Two back ends with one interface:
  * ``OracleNet``  — the plain-C restatement  (oracle/dl_oracle.c  -> oracle/_build/liboracle_{f32,f64}.so)
  * ``RefNet``     — the UNMODIFIED reference (oracle/ref_driver.cpp + /root/reference sources
                                               -> oracle/_ref/libdlref_{f32,f64}.so, prebuilt in the build container)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = "/root/reference"


class LayerSpec(C.Structure):
    """Mirror of dlb_layer_spec (include/dlb200_spec.h)."""
    _fields_ = [(n, C.c_int32) for n in (
        "type", "activation", "in_c", "in_h", "in_w", "win_h", "win_w", "out_c",
        "pad_h", "pad_w", "stride_h", "stride_w", "pool_type", "r0", "r1", "r2")]


FULL, CONVOLUTION, PULL = 1, 2, 3
SIGMOID, TANH, RELU, SOFTMAX, LINEAR = 1, 2, 3, 4, 5
SQUARED_ERROR, CROSS_ENTROPY, COST_LINEAR = 1, 2, 3
POOL_MAX, POOL_MIN, POOL_AVERAGE = 1, 2, 3


def conv(in_chw, win, filters, act, pad=(0, 0), stride=(1, 1)) -> LayerSpec:
    return LayerSpec(CONVOLUTION, act, in_chw[0], in_chw[1], in_chw[2], win[0], win[1], filters,
                     pad[0], pad[1], stride[0], stride[1], 0, 0, 0, 0)


def pool(in_chw, win, pool_type=POOL_MAX) -> LayerSpec:
    return LayerSpec(PULL, 0, in_chw[0], in_chw[1], in_chw[2], win[0], win[1], 0, 0, 0, win[0], win[1], pool_type, 0, 0, 0)


def full(in_chw, out, act) -> LayerSpec:
    return LayerSpec(FULL, act, in_chw[0], in_chw[1], in_chw[2], 0, 0, out, 0, 0, 1, 1, 0, 0, 0, 0)


def conv_out(in_hw, win, pad, stride):
    def one(i, k, p, s):
        t = i + 2 * p - k
        return 0 if t < 0 else t // s + 1
    return one(in_hw[0], win[0], pad[0], stride[0]), one(in_hw[1], win[1], pad[1], stride[1])


def out_shape(spec: LayerSpec):
    if spec.type == FULL:
        return (1, 1, spec.out_c)
    if spec.type == CONVOLUTION:
        h, w = conv_out((spec.in_h, spec.in_w), (spec.win_h, spec.win_w), (spec.pad_h, spec.pad_w), (spec.stride_h, spec.stride_w))
        return (spec.out_c, h, w)
    h, w = conv_out((spec.in_h, spec.in_w), (spec.win_h, spec.win_w), (0, 0), (spec.win_h, spec.win_w))
    return (spec.in_c, h, w)


def lenet_specs():
    """BASELINE.json configs[1]: {1,28,28} -> C 5x5x20 ReLU -> P max 2x2 -> N 2880->100 ReLU -> N 100->10 softmax."""
    return [conv((1, 28, 28), (5, 5), 20, RELU), pool((20, 24, 24), (2, 2), POOL_MAX),
            full((20, 12, 12), 100, RELU), full((1, 1, 100), 10, SOFTMAX)]


def mlp_specs():
    """BASELINE.json configs[0]: 784-100-10 sigmoid MLP."""
    return [full((1, 1, 784), 100, SIGMOID), full((1, 1, 100), 10, SIGMOID)]


def cifar_specs():
    """BASELINE.json configs[2] as concretised in SURVEY.md §8(d)."""
    return [conv((3, 32, 32), (5, 5), 32, RELU, pad=(2, 2)), pool((32, 32, 32), (2, 2)),
            conv((32, 16, 16), (5, 5), 64, RELU, pad=(2, 2)), pool((64, 16, 16), (2, 2)),
            conv((64, 8, 8), (3, 3), 128, RELU, pad=(1, 1)), pool((128, 8, 8), (2, 2)),
            full((128, 4, 4), 256, RELU), full((1, 1, 256), 10, SOFTMAX)]


def dtype_of(real: str):
    return {"f32": np.float32, "f64": np.float64}[real]


def build_oracle() -> None:
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def build_ref() -> bool:
    """Builds oracle/_ref when the reference sources are present; returns whether the libraries exist afterwards."""
    if os.path.isdir(REFERENCE_ROOT):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])
    return ref_available()


def ref_available() -> bool:
    return all(os.path.exists(os.path.join(HERE, "_ref", f"libdlref_{r}.so")) for r in ("f32", "f64"))


_libs = {}


def _load(kind: str, real: str) -> C.CDLL:
    key = (kind, real)
    if key in _libs:
        return _libs[key]
    if kind == "oracle":
        path = os.path.join(HERE, "_build", f"liboracle_{real}.so")
        if not os.path.exists(path):
            build_oracle()
    else:
        path = os.path.join(HERE, "_ref", f"libdlref_{real}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} is missing: run `make -C oracle ref` in the build container")
    lib = C.CDLL(path)
    p = "orc_" if kind == "oracle" else "dlref_"
    vp, i64, dbl, i32 = C.c_void_p, C.c_int64, C.c_double, C.c_int
    getattr(lib, p + "last_error").restype = C.c_char_p
    getattr(lib, p + "real_bytes").restype = C.c_int
    f = getattr(lib, p + "net_create"); f.restype = vp
    f.argtypes = [C.POINTER(LayerSpec), i32] + ([C.c_uint] if kind == "ref" else [])
    getattr(lib, p + "net_destroy").argtypes = [vp]
    f = getattr(lib, p + "net_param_count"); f.restype = i64; f.argtypes = [vp]
    for name in ("net_layer_in_size", "net_layer_out_size"):
        f = getattr(lib, p + name); f.restype = i64; f.argtypes = [vp, i32]
    getattr(lib, p + "net_set_params").argtypes = [vp, vp]
    getattr(lib, p + "net_get_params").argtypes = [vp, vp]
    getattr(lib, p + "net_trace").argtypes = [vp, vp, vp, i32, dbl, vp, vp, vp, vp, vp, vp]
    getattr(lib, p + "net_eval").argtypes = [vp, vp, i64, vp, vp]
    f = getattr(lib, p + "cost"); f.restype = dbl; f.argtypes = [vp, vp, i64, i32]
    if kind == "oracle":
        lib.orc_net_learn.argtypes = [vp, vp, vp, i64, i64, i64, dbl, i32, dbl, i32, vp]
        lib.orc_net_learn_dropout.argtypes = [vp, vp, vp, i64, i64, i64, dbl, i32, dbl, i32, vp, vp]
        lib.orc_net_set_keep_rate.argtypes = [vp, i32, dbl]
        lib.orc_net_mask_size.restype = i64; lib.orc_net_mask_size.argtypes = [vp]
        lib.orc_net_bind_masks.argtypes = [vp, vp]
    else:
        lib.dlref_net_create_dropout.restype = vp
        lib.dlref_net_create_dropout.argtypes = [C.POINTER(LayerSpec), i32, C.c_uint, vp]
        lib.dlref_dropout_masks.argtypes = [C.c_uint, vp, vp, i32, i64, vp]
        lib.dlref_net_learn.argtypes = [vp, vp, vp, i64, i64, i64, dbl, i32, dbl, i32, i64, vp]
        lib.dlref_uniform_items.argtypes = [vp, i64, i32, i32, i32, dbl, dbl, C.c_uint]
        lib.dlref_reset_random_generator.argtypes = [C.c_uint]
        lib.dlref_shuffle_sequence.argtypes = [C.c_uint, i64, i64, vp]
    _libs[key] = lib
    return lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


class _NetBase:
    kind = ""
    prefix = ""

    def __init__(self, specs: Sequence[LayerSpec], real: str = "f32", seed: int = 0, keep_rates=None):
        """keep_rates (optional): one keep_rate per layer (ALayer::_keep_rate, 1 = no dropout)."""
        self.real = real
        self.dtype = dtype_of(real)
        self.lib = _load(self.kind, real)
        assert getattr(self.lib, self.prefix + "real_bytes")() == np.dtype(self.dtype).itemsize
        self.specs = list(specs)
        arr = (LayerSpec * len(self.specs))(*self.specs)
        self.keep_rates = [1.0] * len(self.specs) if keep_rates is None else [float(k) for k in keep_rates]
        assert len(self.keep_rates) == len(self.specs)
        if self.kind == "ref":
            kr = np.asarray(self.keep_rates, np.float64)
            self.h = self.lib.dlref_net_create_dropout(arr, len(self.specs), seed, _ptr(kr))
        else:
            self.h = self.lib.orc_net_create(arr, len(self.specs))
        if not self.h:
            raise ValueError(self._err())
        if self.kind == "oracle":
            for l, k in enumerate(self.keep_rates):
                self._check(self.lib.orc_net_set_keep_rate(self.h, l, k))
        self.n_layers = len(self.specs)
        self.out_sizes = [self._f("net_layer_out_size")(self.h, l) for l in range(self.n_layers)]
        self.in_sizes = [self._f("net_layer_in_size")(self.h, l) for l in range(self.n_layers)]
        self.param_count = self._f("net_param_count")(self.h)

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def _err(self) -> str:
        return self._f("last_error")().decode()

    def _check(self, rc):
        if rc:
            raise RuntimeError(self._err())

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self._f("net_destroy")(self.h)
                self.h = None
        except Exception:
            pass

    def get_params(self) -> np.ndarray:
        p = np.empty(self.param_count, self.dtype)
        self._check(self._f("net_get_params")(self.h, _ptr(p)))
        return p

    def set_params(self, p: np.ndarray) -> None:
        p = np.ascontiguousarray(p, self.dtype)
        assert p.size == self.param_count
        self._check(self._f("net_set_params")(self.h, _ptr(p)))

    def trace(self, x, target, cost_id=CROSS_ENTROPY, scale=0.0, grads_in=None):
        """One sample forward/backward; returns dict(acts, derivs, index_map, grads, in_grads, cost_grad)."""
        x = np.ascontiguousarray(x, self.dtype).ravel()
        target = np.ascontiguousarray(target, self.dtype).ravel()
        na, ni = sum(self.out_sizes), sum(self.in_sizes)
        acts = np.zeros(na, self.dtype); derivs = np.zeros(na, self.dtype)
        imap = np.full(na, -7, np.int32)
        grads = np.zeros(self.param_count, self.dtype) if grads_in is None else np.array(grads_in, self.dtype)
        in_grads = np.zeros(ni, self.dtype)
        cost_grad = np.zeros(self.out_sizes[-1], self.dtype)
        self._check(self._f("net_trace")(self.h, _ptr(x), _ptr(target), cost_id, float(scale), _ptr(acts), _ptr(derivs),
                                         _ptr(imap), _ptr(grads), _ptr(in_grads), _ptr(cost_grad)))
        return dict(acts=acts, derivs=derivs, index_map=imap, grads=grads, in_grads=in_grads, cost_grad=cost_grad)

    def eval(self, inputs, want_outputs=False):
        inputs = np.ascontiguousarray(inputs, self.dtype).reshape(-1, self.in_sizes[0])
        n = inputs.shape[0]
        classes = np.zeros(n, np.int32)
        outs = np.zeros((n, self.out_sizes[-1]), self.dtype) if want_outputs else None
        self._check(self._f("net_eval")(self.h, _ptr(inputs), n, _ptr(classes), _ptr(outs)))
        return (classes, outs) if want_outputs else classes

    def cost(self, output, reference, cost_id=CROSS_ENTROPY) -> float:
        o = np.ascontiguousarray(output, self.dtype).ravel(); r = np.ascontiguousarray(reference, self.dtype).ravel()
        return self._f("cost")(_ptr(o), _ptr(r), o.size, cost_id)


class OracleNet(_NetBase):
    kind, prefix = "oracle", "orc_"

    def learn(self, inputs, refs, batch, epochs, lr, cost_id, lam, threads, perms, masks=None):
        """masks: [n_batches, mask_size] keep masks (layers with keep_rate < 1 in layer order), one row per mini-batch."""
        inputs = np.ascontiguousarray(inputs, self.dtype).reshape(-1, self.in_sizes[0])
        refs = np.ascontiguousarray(refs, self.dtype).reshape(-1, self.out_sizes[-1])
        perms = np.ascontiguousarray(perms, np.int64).reshape(epochs, inputs.shape[0])
        if masks is not None:
            masks = np.ascontiguousarray(masks, self.dtype)
            assert masks.ndim == 2 and masks.shape[1] == self.mask_size
        self._check(self.lib.orc_net_learn_dropout(self.h, _ptr(inputs), _ptr(refs), inputs.shape[0], batch, epochs, float(lr),
                                                   cost_id, float(lam), threads, _ptr(perms), _ptr(masks)))

    @property
    def mask_size(self) -> int:
        return int(self.lib.orc_net_mask_size(self.h))

    def bind_masks(self, masks) -> None:
        """One mini-batch's masks for trace(); kept alive here.  None clears them."""
        self._masks = None if masks is None else np.ascontiguousarray(masks, self.dtype).ravel()
        if self._masks is not None:
            assert self._masks.size == self.mask_size
        self._check(self.lib.orc_net_bind_masks(self.h, _ptr(self._masks)))


class RefNet(_NetBase):
    kind, prefix = "ref", "dlref_"

    def learn(self, inputs, refs, batch, epochs, lr, cost_id, lam, parallelism=-1, shuffle_seed=0) -> float:
        """Net<CpuDC>::learn; returns seconds spent inside learn()."""
        inputs = np.ascontiguousarray(inputs, self.dtype).reshape(-1, self.in_sizes[0])
        refs = np.ascontiguousarray(refs, self.dtype).reshape(-1, self.out_sizes[-1])
        sec = C.c_double(0)
        self._check(self.lib.dlref_net_learn(self.h, _ptr(inputs), _ptr(refs), inputs.shape[0], batch, epochs, float(lr),
                                             cost_id, float(lam), parallelism, shuffle_seed, C.byref(sec)))
        return sec.value


def ref_uniform_items(real: str, n: int, chw, lo=0.0, hi=1.0, seed=7) -> np.ndarray:
    """Synthetic inputs exactly as the reference's Tensor(size, lo, hi, &mt19937(seed))."""
    lib = _load("ref", real)
    out = np.zeros((n, chw[0] * chw[1] * chw[2]), dtype_of(real))
    rc = lib.dlref_uniform_items(_ptr(out), n, chw[0], chw[1], chw[2], lo, hi, seed)
    if rc:
        raise RuntimeError(lib.dlref_last_error().decode())
    return out


def one_hot_labels(real: str, n: int, classes: int = 10) -> np.ndarray:
    y = np.zeros((n, classes), dtype_of(real))
    y[np.arange(n), np.arange(n) % classes] = 1
    return y


def ref_shuffle_sequence(seed: int, n: int, epochs: int) -> np.ndarray:
    """Epoch permutations of Net<CpuDC>::learn after reset_random_generator(seed) (libstdc++ std::ranges::shuffle)."""
    lib = _load("ref", "f32")
    out = np.zeros((epochs, n), np.int64)
    if lib.dlref_shuffle_sequence(seed, n, epochs, _ptr(out)):
        raise RuntimeError(lib.dlref_last_error().decode())
    return out


def count_batches(n: int, batch: int) -> int:
    """Mini-batches per epoch under the tail rule of Net.inl:235-236."""
    k, start = 0, 0
    while start < n:
        start = n if (n - start) < 1.5 * batch else start + batch
        k += 1
    return k


def ref_dropout_masks(real: str, seed: int, in_sizes, keep_rates, n_batches: int) -> np.ndarray:
    """The keep masks Net<CpuDC>::learn draws after reset_random_generator(seed): [n_batches, sum(in_size for keep < 1)]
    from the reference's own Vector::fill_with_random_selection_map (libstdc++ mt19937 + uniform_real_distribution + sort)."""
    lib = _load("ref", real)
    sizes = np.asarray(in_sizes, np.int64); keep = np.asarray(keep_rates, np.float64)
    width = int(sum(s for s, k in zip(in_sizes, keep_rates) if np.dtype(dtype_of(real)).type(k) < 1))
    out = np.zeros((n_batches, width), dtype_of(real))
    if lib.dlref_dropout_masks(seed, _ptr(sizes), _ptr(keep), len(in_sizes), n_batches, _ptr(out)):
        raise RuntimeError(lib.dlref_last_error().decode())
    return out


# ---- the recurrent stack (oracle/ref_mnet_driver.cpp -> oracle/_ref/libdlrefm_{f32,f64}.so) -----------------------------------
def mnet_ref_available() -> bool:
    return all(os.path.exists(os.path.join(HERE, "_ref", f"libdlrefm_{r}.so")) for r in ("f32", "f64"))


class RefMNet:
    """The reference's own MNet<CpuDC> of RMLayer<CpuDC> (unmodified sources), driven through oracle/ref_mnet_driver.cpp."""

    def __init__(self, depth, dims, activations, clips, real="f32", seed=0):
        key = ("refm", real)
        if key not in _libs:
            lib = C.CDLL(os.path.join(HERE, "_ref", f"libdlrefm_{real}.so"))
            lib.dlrefm_last_error.restype = C.c_char_p
            lib.dlrefm_create.restype = C.c_void_p
            lib.dlrefm_create.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint]
            lib.dlrefm_destroy.argtypes = [C.c_void_p]
            lib.dlrefm_param_count.argtypes = [C.c_void_p]; lib.dlrefm_param_count.restype = C.c_int64
            lib.dlrefm_get_params.argtypes = [C.c_void_p, C.c_void_p]
            lib.dlrefm_set_params.argtypes = [C.c_void_p, C.c_void_p]
            lib.dlrefm_act.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
            lib.dlrefm_gradient_sum.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
            lib.dlrefm_learn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_double]
            _libs[key] = lib
        self.lib = _libs[key]
        self.dtype = np.float32 if real == "f32" else np.float64
        assert self.lib.dlrefm_real_bytes() == np.dtype(self.dtype).itemsize
        self.depth, self.dims = int(depth), list(dims)
        d = np.asarray(dims, np.int32); a = np.asarray(activations, np.int32); c = np.asarray(clips, np.float64)
        self.h = self.lib.dlrefm_create(self.depth, len(a), d.ctypes.data, a.ctypes.data, c.ctypes.data, seed)
        if not self.h: raise RuntimeError(self.lib.dlrefm_last_error().decode())
        self.param_count = int(self.lib.dlrefm_param_count(self.h))

    def __del__(self):
        try:
            if getattr(self, "h", None): self.lib.dlrefm_destroy(self.h); self.h = None
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0: raise RuntimeError(self.lib.dlrefm_last_error().decode())

    def get_params(self):
        p = np.empty(self.param_count, self.dtype); self._check(self.lib.dlrefm_get_params(self.h, p.ctypes.data)); return p

    def set_params(self, p):
        p = np.ascontiguousarray(p, self.dtype); assert p.size == self.param_count
        self._check(self.lib.dlrefm_set_params(self.h, p.ctypes.data))

    def act(self, xs):
        xs = np.ascontiguousarray(xs, self.dtype)
        out = np.empty((xs.shape[0], self.depth, self.dims[-1]), self.dtype)
        self._check(self.lib.dlrefm_act(self.h, xs.ctypes.data, xs.shape[0], out.ctypes.data)); return out

    def calc_gradient_sum(self, xs, ys, cost_id):
        xs = np.ascontiguousarray(xs, self.dtype); ys = np.ascontiguousarray(ys, self.dtype)
        g = np.empty(self.param_count, self.dtype)
        self._check(self.lib.dlrefm_gradient_sum(self.h, xs.ctypes.data, ys.ctypes.data, xs.shape[0], cost_id, g.ctypes.data)); return g

    def learn(self, xs, ys, cost_id, learning_rate, reg_factor=0.0):
        xs = np.ascontiguousarray(xs, self.dtype); ys = np.ascontiguousarray(ys, self.dtype)
        self._check(self.lib.dlrefm_learn(self.h, xs.ctypes.data, ys.ctypes.data, xs.shape[0], cost_id, float(learning_rate), float(reg_factor)))
