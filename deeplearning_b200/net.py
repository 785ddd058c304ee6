"""Host-side mirror (Python) of the reference's Net / CLayer / PLayer / NLayer interface for the GPU data context.

Same names, argument meaning and error behaviour as the reference classes for the training path:
  Net        DeepLearning/NeuralNet/Net.h:35-420, Net.inl          (append_layer, learn, act, count_correct_answers ...)
  CLayer     DeepLearning/NeuralNet/CLayer.h:145-148               (in_size, filter_window_size, filters_count, func_id, paddings, strides)
  PLayer     DeepLearning/NeuralNet/PLayer.h:104-105               (in_size, pool_window_size, pool_operator_id)
  NLayer     DeepLearning/NeuralNet/NLayer.h:130-132               (in_dim, out_dim, func_id)
Every computation goes through the C ABI (include/dlb200.h) to libdlb200.so; this module holds no arithmetic.
"""
from __future__ import annotations

import ctypes as C
import enum
from typing import Callable, Optional, Sequence

import numpy as np

from . import _capi
from ._capi import DlbError, LayerSpec


class LayerTypeId(enum.IntEnum):      # NeuralNet/LayerTypeId.h:27-32
    FULL = 1
    CONVOLUTION = 2
    PULL = 3


class ActivationFunctionId(enum.IntEnum):   # Math/ActivationFunctionId.h:26-33
    SIGMOID = 1
    TANH = 2
    RELU = 3
    SOFTMAX = 4
    LINEAR = 5


class CostFunctionId(enum.IntEnum):   # Math/CostFunctionId.h:26-32
    SQUARED_ERROR = 1
    CROSS_ENTROPY = 2
    LINEAR = 3


class PoolTypeId(enum.IntEnum):       # Math/PoolOperator.h:28-33
    MAX = 1
    MIN = 2
    AVERAGE = 3


# dlb_net_read buffer ids
BUF_OUTPUT, BUF_DERIVATIVES, BUF_INDEX_MAP, BUF_INPUT_GRAD, BUF_COST_GRAD, BUF_GRADIENT = 1, 2, 3, 4, 5, 6


def _i3(v) -> tuple:
    if isinstance(v, int):
        return (v, v, v)
    v = tuple(int(t) for t in v)
    if len(v) != 3:
        raise ValueError("expected an Index3d (layers, rows, columns)")
    return v


def _i2(v) -> tuple:
    if isinstance(v, int):
        return (v, v)
    v = tuple(int(t) for t in v)
    if len(v) != 2:
        raise ValueError("expected an Index2d (rows, columns)")
    return v


def CLayer(in_size, filter_window_size, filters_count: int, func_id: ActivationFunctionId,
           paddings=(0, 0, 0), strides=(1, 1, 1)) -> LayerSpec:
    """CLayer(in_size, filter_window_size, filters_count, func_id, paddings = {0}, strides = {1}) — CLayer.h:145-148.

    The filter spans all input channels (CLayer.inl:34-40), so the channel padding must be 0 and the channel stride
    is irrelevant; anything else makes the reference constructor throw "Unexpected channel size" as well.
    """
    c, h, w = _i3(in_size); kh, kw = _i2(filter_window_size); p = _i3(paddings); s = _i3(strides)
    if p[0] != 0:
        raise ValueError("Unexpected channel size")
    return LayerSpec(LayerTypeId.CONVOLUTION, int(func_id), c, h, w, kh, kw, int(filters_count), p[1], p[2], s[1], s[2], 0, 0, 0, 0)


def PLayer(in_size, pool_window_size, pool_operator_id: PoolTypeId) -> LayerSpec:
    """PLayer(in_size, pool_window_size, pool_operator_id) — PLayer.h:104-105; stride == window, no padding."""
    c, h, w = _i3(in_size); wh, ww = _i2(pool_window_size)
    return LayerSpec(LayerTypeId.PULL, 0, c, h, w, wh, ww, 0, 0, 0, wh, ww, int(pool_operator_id), 0, 0, 0)


def NLayer(in_dim, out_dim: int, func_id: ActivationFunctionId = ActivationFunctionId.SIGMOID) -> LayerSpec:
    """NLayer(in_dim, out_dim, func_id) — NLayer.h:130-132.  in_dim may be an Index3d; it is flattened."""
    c, h, w = (1, 1, int(in_dim)) if isinstance(in_dim, (int, np.integer)) else _i3(in_dim)
    return LayerSpec(LayerTypeId.FULL, int(func_id), c, h, w, 0, 0, int(out_dim), 0, 0, 1, 1, 0, 0, 0, 0)


def _conv_out(i, k, p, s):
    t = i + 2 * p - k
    return 0 if t < 0 else t // s + 1


def out_size(spec: LayerSpec) -> tuple:
    """ALayer::out_size() as an Index3d."""
    if spec.type == LayerTypeId.FULL:
        return (1, 1, spec.out_c)
    if spec.type == LayerTypeId.CONVOLUTION:
        return (spec.out_c, _conv_out(spec.in_h, spec.win_h, spec.pad_h, spec.stride_h), _conv_out(spec.in_w, spec.win_w, spec.pad_w, spec.stride_w))
    return (spec.in_c, _conv_out(spec.in_h, spec.win_h, 0, spec.win_h), _conv_out(spec.in_w, spec.win_w, 0, spec.win_w))


class Context:
    """One device + one stream (dlb_context).  stream: an existing cudaStream_t handle (int) or None."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self.lib = _capi.load()
        h = C.c_void_p()
        _capi.check(self.lib.dlb_context_create(device, C.c_void_p(stream) if stream else None, C.byref(h)))
        self.h = h
        self.device = device

    def synchronize(self) -> None:
        _capi.check(self.lib.dlb_context_synchronize(self.h))

    def launch_count(self) -> int:
        return int(self.lib.dlb_context_launch_count(self.h))

    @property
    def stream(self) -> int:
        return int(self.lib.dlb_context_stream(self.h) or 0)

    def close(self) -> None:
        if getattr(self, "h", None):
            self.lib.dlb_context_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Event:
    """A CUDA event recorded on a Context's stream (device-side timing)."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        h = C.c_void_p()
        _capi.check(ctx.lib.dlb_event_create(C.byref(h)))
        self.h = h

    def record(self) -> None:
        _capi.check(self.ctx.lib.dlb_event_record(self.ctx.h, self.h))

    def elapsed_ms(self, stop: "Event") -> float:
        ms = C.c_double(0.0)
        _capi.check(self.ctx.lib.dlb_event_elapsed_ms(self.h, stop.h, C.byref(ms)))
        return ms.value

    def __del__(self):
        try:
            self.ctx.lib.dlb_event_destroy(self.h)
        except Exception:
            pass


def timing_enable(on: bool) -> None:
    _capi.check(_capi.load().dlb_timing_enable(1 if on else 0))


def timing_report() -> dict:
    """{"<layer>:<operator>:<kernel>": (launches, total_ms)} since the last report."""
    buf = C.create_string_buffer(1 << 16)
    _capi.check(_capi.load().dlb_timing_report(buf, len(buf)))
    out = {}
    for line in buf.value.decode().splitlines():
        k, n, ms = line.split("\t")
        out[k] = (int(n), float(ms))
    return out


def rank_slice(count: int, rank: int, world: int) -> tuple:
    first, mine = C.c_int64(0), C.c_int64(0)
    _capi.check(_capi.load().dlb_rank_slice(count, rank, world, C.byref(first), C.byref(mine)))
    return int(first.value), int(mine.value)


def pinned_array(shape, dtype) -> np.ndarray:
    """A numpy array backed by page-locked host memory (cudaMallocHost); freed when the process exits."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    _capi.check(_capi.load().dlb_host_alloc_pinned(n, C.byref(p)))
    buf = (C.c_char * n).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


class Net:
    """Net<GpuDC> — the batched B200 implementation behind the reference's Net interface.

    real: "f32" (the reference's USE_SINGLE_PRECISION builds) or "f64" (its default `Real`).
    """

    def __init__(self, layers: Sequence[LayerSpec] = (), real: str = "f32", context: Optional[Context] = None, device: int = 0):
        self.real = real
        self.dtype = {"f32": np.float32, "f64": np.float64}[real]
        self.ctx = context or Context(device)
        self.lib = self.ctx.lib
        self._specs = list(layers)
        self._h = None
        self._keep = []     # host arrays the device may still be reading

    # ---- construction -------------------------------------------------------------------------------------------
    def append_layer(self, spec: LayerSpec) -> tuple:
        """Net::append_layer<L>(args...) — returns the output size of the appended layer (Net.h:281-286)."""
        if self._h is not None:
            raise DlbError("layers cannot be appended after the net has been used")
        self._specs.append(spec)
        return out_size(spec)

    def _handle(self):
        if self._h is None:
            arr = (LayerSpec * len(self._specs))(*self._specs)
            h = C.c_void_p()
            _capi.check(self.lib.dlb_net_create(self.ctx.h, arr, len(self._specs), np.dtype(self.dtype).itemsize, C.byref(h)))
            self._h = h
        return self._h

    def close(self) -> None:
        if self._h is not None:
            self.lib.dlb_net_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- introspection ------------------------------------------------------------------------------------------
    def layers_count(self) -> int:
        return len(self._specs)

    def in_size(self) -> tuple:
        s = self._specs[0]
        return (s.in_c, s.in_h, s.in_w)

    def out_size(self) -> tuple:
        return out_size(self._specs[-1])

    @property
    def param_count(self) -> int:
        return int(self.lib.dlb_net_param_count(self._handle()))

    def layer_in_size(self, l: int) -> int:
        return int(self.lib.dlb_net_layer_in_size(self._handle(), l))

    def layer_out_size(self, l: int) -> int:
        return int(self.lib.dlb_net_layer_out_size(self._handle(), l))

    def layer_param_offset(self, l: int) -> int:
        return int(self.lib.dlb_net_layer_param_offset(self._handle(), l))

    def describe_plan(self) -> str:
        buf = C.create_string_buffer(512)
        _capi.check(self.lib.dlb_net_describe_plan(self._handle(), buf, len(buf)))
        return buf.value.decode()

    def enable_tensor_cores(self, mode) -> None:
        """fp32 NLayer contractions: 0 / False = packed-FMA kernels, 1 / True = tcgen05 3xTF32 for mini-batches of at
        least 128 samples per GPU (the default), 2 = tcgen05 at every batch size."""
        _capi.check(self.lib.dlb_net_enable_tensor_cores(self._handle(), int(mode)))

    def enable_step_kernel(self, on: bool = True) -> None:
        """The whole step as one persistent cooperative kernel (csrc/mega.cu) where the net qualifies (LeNet-shaped float nets)."""
        _capi.check(self.lib.dlb_net_enable_step_kernel(self._handle(), 1 if on else 0))

    def set_keep_rate(self, layer: int, keep_rate: float) -> None:
        """ALayer::set_keep_rate / the keep_rate constructor argument of CLayer, PLayer and NLayer (one minus the dropout rate
        of the layer's input, ALayer.inl:25-55)."""
        _capi.check(self.lib.dlb_net_set_keep_rate(self._handle(), int(layer), float(keep_rate)))

    @property
    def mask_size(self) -> int:
        return int(self.lib.dlb_net_mask_size(self._handle()))

    def set_dropout_masks(self, masks) -> None:
        """Explicit keep masks of one mini-batch for forward_backward() (layers with keep_rate < 1 in layer order)."""
        m = np.ascontiguousarray(masks, self.dtype).ravel()
        if m.size != self.mask_size:
            raise ValueError(f"expected {self.mask_size} mask elements")
        _capi.check(self.lib.dlb_net_set_dropout_masks(self._handle(), m.ctypes.data))

    def enable_graphs(self, on: bool) -> None:
        _capi.check(self.lib.dlb_net_enable_graphs(self._handle(), 1 if on else 0))

    # ---- parameters (GpuDC::from_host / to_host) ----------------------------------------------------------------
    def set_params(self, flat) -> None:
        p = np.ascontiguousarray(flat, self.dtype).ravel()
        if p.size != self.param_count:
            raise DlbError("Collections must be of the same size")
        _capi.check(self.lib.dlb_net_set_params(self._handle(), p.ctypes.data))

    def get_params(self) -> np.ndarray:
        p = np.empty(self.param_count, self.dtype)
        _capi.check(self.lib.dlb_net_get_params(self._handle(), p.ctypes.data))
        return p

    def init_params(self, seed: int = 0) -> None:
        """Random initialisation with the reference's distributions (NLayer.inl:27-39, CLayer.inl:42-48): biases
        U(-1, 1), weights / filters N(0, 1/sqrt(fan_in)).  numpy's generator, so statistically - not bit - equal to a
        reference net; parity tests copy the oracle's parameters with set_params instead."""
        rng = np.random.default_rng(seed)
        flat = np.zeros(self.param_count, self.dtype)
        off = 0
        for s in self._specs:
            if s.type == LayerTypeId.PULL:
                continue
            o = out_size(s)
            if s.type == LayerTypeId.FULL:
                nb, fan_in = s.out_c, s.in_c * s.in_h * s.in_w
                nw = nb * fan_in
            else:
                nb, fan_in = o[0] * o[1] * o[2], s.in_c * s.win_h * s.win_w
                nw = s.out_c * fan_in
            flat[off:off + nb] = rng.uniform(-1, 1, nb); off += nb
            flat[off:off + nw] = rng.normal(0, 1 / np.sqrt(fan_in), nw); off += nw
        self.set_params(flat)

    @staticmethod
    def reset_random_generator_static(seed: int) -> None:  # pragma: no cover - kept for API symmetry
        raise NotImplementedError("use the instance method reset_random_generator(seed)")

    def reset_random_generator(self, seed: int) -> None:
        """Net::reset_random_generator(seed), Net.inl:599-603 (the generator that shuffles the epochs)."""
        _capi.check(self.lib.dlb_net_reset_random_generator(self._handle(), seed))

    # ---- training -----------------------------------------------------------------------------------------------
    def _as_items(self, a, width: int) -> np.ndarray:
        a = np.ascontiguousarray(a, self.dtype)
        return a.reshape(-1, width)

    def set_dataset(self, training_items, reference_items) -> int:
        x = self._as_items(training_items, self.layer_in_size(0))
        y = self._as_items(reference_items, self.layer_out_size(self.layers_count() - 1))
        if x.shape[0] != y.shape[0]:
            raise DlbError("Incompatible collection of training and reference items.")
        _capi.check(self.lib.dlb_net_set_dataset(self._handle(), x.ctypes.data, y.ctypes.data, x.shape[0], 0))
        return x.shape[0]

    def learn(self, training_items, reference_items, batch_size: int, epochs_count: int, learning_rate: float,
              cost_func_id: CostFunctionId, lambda_: float = 0.0,
              epoch_callback: Optional[Callable[[int, float], None]] = None) -> None:
        """Net::learn(training_items, reference_items, batch_size, epochs_count, learning_rate, cost_func_id, lambda,
        epoch_callback) — Net.inl:193-324.  training_items / reference_items may be None to reuse the resident data set."""
        if training_items is not None:
            if len(training_items) != len(reference_items):
                raise DlbError("Incompatible collection of training and reference items.")
            if len(training_items) == 0 or batch_size == 0:
                return
            self.set_dataset(training_items, reference_items)
        cb = _capi.EPOCH_CALLBACK(lambda e, r, _u: epoch_callback(int(e), float(r))) if epoch_callback else None
        _capi.check(self.lib.dlb_net_learn(self._handle(), batch_size, epochs_count, float(learning_rate), int(cost_func_id),
                                           float(lambda_), C.cast(cb, C.c_void_p) if cb else None, None))

    def train_steps(self, batch_size: int, n_steps: int, learning_rate: float, cost_func_id: CostFunctionId, lambda_: float = 0.0) -> int:
        """The next n_steps mini-batches of learn()'s epoch sequence on the resident data set; returns items consumed."""
        items = C.c_int64(0)
        _capi.check(self.lib.dlb_net_train_steps(self._handle(), batch_size, n_steps, float(learning_rate), int(cost_func_id),
                                                 float(lambda_), C.byref(items)))
        return int(items.value)

    def step_indices(self, indices, learning_rate: float, cost_func_id: CostFunctionId, reg_factor: float = 0.0) -> None:
        idx = np.ascontiguousarray(indices, np.int64)
        _capi.check(self.lib.dlb_net_step_indices(self._handle(), idx.ctypes.data, idx.size, float(learning_rate), int(cost_func_id), float(reg_factor)))

    def step_host(self, inputs: np.ndarray, refs: np.ndarray, learning_rate: float, cost_func_id: CostFunctionId,
                  reg_factor: float = 0.0, want_cost: bool = True) -> Optional[float]:
        """One mini-batch step on host buffers (already contiguous arrays of the net's dtype, ideally pinned)."""
        n = inputs.shape[0]
        cost = C.c_double(0.0)
        _capi.check(self.lib.dlb_net_step_host(self._handle(), inputs.ctypes.data, refs.ctypes.data, n, float(learning_rate),
                                               int(cost_func_id), float(reg_factor), C.byref(cost) if want_cost else None))
        return cost.value if want_cost else None

    def step_host_async(self, inputs: np.ndarray, refs: np.ndarray, learning_rate: float, cost_func_id: CostFunctionId,
                        reg_factor: float = 0.0, cost_slot: Optional[np.ndarray] = None) -> None:
        """Enqueue one mini-batch step on host buffers without waiting: the copy of this batch overlaps the previous step.
        cost_slot: a one-element float64 array in pinned memory that receives the batch cost; valid after ctx.synchronize()."""
        _capi.check(self.lib.dlb_net_step_host_async(self._handle(), inputs.ctypes.data, refs.ctypes.data, inputs.shape[0],
                                                     float(learning_rate), int(cost_func_id), float(reg_factor),
                                                     cost_slot.ctypes.data if cost_slot is not None else None))

    def forward_backward(self, inputs, refs, cost_func_id: CostFunctionId, scale: float = 0.0) -> int:
        x = self._as_items(inputs, self.layer_in_size(0))
        y = self._as_items(refs, self.layer_out_size(self.layers_count() - 1))
        _capi.check(self.lib.dlb_net_forward_backward_host(self._handle(), x.ctypes.data, y.ctypes.data, x.shape[0], int(cost_func_id), float(scale)))
        self._last_batch = x.shape[0]
        return x.shape[0]

    def read(self, what: int, layer: int = 0) -> np.ndarray:
        n = getattr(self, "_last_batch", 0)
        if what == BUF_GRADIENT:
            a = np.empty(self.param_count, self.dtype)
        elif what == BUF_INDEX_MAP:
            a = np.empty((n, self.layer_out_size(layer)), np.int32)
        elif what == BUF_INPUT_GRAD:
            a = np.empty((n, self.layer_in_size(layer)), self.dtype)
        elif what == BUF_COST_GRAD:
            a = np.empty((n, self.layer_out_size(self.layers_count() - 1)), self.dtype)
        else:
            a = np.empty((n, self.layer_out_size(layer)), self.dtype)
        _capi.check(self.lib.dlb_net_read(self._handle(), what, layer, a.ctypes.data, a.nbytes))
        return a

    def write_gradient(self, flat) -> None:
        g = np.ascontiguousarray(flat, self.dtype).ravel()
        if g.size != self.param_count:
            raise DlbError("Collections must be of the same size")
        _capi.check(self.lib.dlb_net_write_gradient(self._handle(), g.ctypes.data))

    # ---- inference ----------------------------------------------------------------------------------------------
    def act(self, inputs) -> np.ndarray:
        """Net::act(input) for one item or a batch of items; returns the outputs."""
        x = self._as_items(inputs, self.layer_in_size(0))
        out = np.empty((x.shape[0], self.layer_out_size(self.layers_count() - 1)), self.dtype)
        _capi.check(self.lib.dlb_net_eval_host(self._handle(), x.ctypes.data, x.shape[0], None, out.ctypes.data))
        return out

    def predict(self, inputs) -> np.ndarray:
        """max_element_id of the output of every item (Net.inl:432-433)."""
        x = self._as_items(inputs, self.layer_in_size(0))
        cls = np.empty(x.shape[0], np.int32)
        _capi.check(self.lib.dlb_net_eval_host(self._handle(), x.ctypes.data, x.shape[0], cls.ctypes.data, None))
        return cls

    def count_correct_answers(self, test_input, labels) -> int:
        """Net::count_correct_answers, Net.inl:441-472."""
        y = self._as_items(labels, self.layer_out_size(self.layers_count() - 1))
        x = self._as_items(test_input, self.layer_in_size(0))
        if x.shape[0] != y.shape[0]:
            raise DlbError("Invalid input.")
        return int(np.sum(self.predict(x) == np.argmax(y, axis=1)))

    def evaluate_cost_function_and_correct_answers(self, test_input, reference_output, cost_func_id: CostFunctionId,
                                                   l2_reg_factor: float = 0.0) -> tuple:
        """Net::evaluate_cost_function_and_correct_answers, Net.inl:411-439 -> (Cost, CorrectAnswers)."""
        x = self._as_items(test_input, self.layer_in_size(0))
        y = self._as_items(reference_output, self.layer_out_size(self.layers_count() - 1))
        if x.shape[0] != y.shape[0]:
            raise DlbError("Invalid input.")
        cost, correct = C.c_double(0.0), C.c_double(0.0)
        _capi.check(self.lib.dlb_net_evaluate_host(self._handle(), x.ctypes.data, y.ctypes.data, x.shape[0], int(cost_func_id),
                                                   float(l2_reg_factor), C.byref(cost), C.byref(correct)))
        return cost.value, correct.value

    # ---- data parallelism ---------------------------------------------------------------------------------------
    def set_distributed(self, rank: int, world_size: int, unique_id: Optional[bytes]) -> None:
        buf = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        _capi.check(self.lib.dlb_net_set_distributed(self._handle(), rank, world_size, buf))


def _p2p_export(net: "Net") -> bytes:
    buf = C.create_string_buffer(64)
    _capi.check(net.lib.dlb_net_p2p_export(net._handle(), buf))
    return buf.raw


def _p2p_connect(net: "Net", rank: int, world_size: int, handles) -> None:
    blob = b"".join(handles)
    assert len(blob) == 64 * world_size
    _capi.check(net.lib.dlb_net_set_distributed_p2p(net._handle(), rank, world_size, C.create_string_buffer(blob, len(blob))))


Net.p2p_export = _p2p_export
Net.set_distributed_p2p = _p2p_connect


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _capi.check(_capi.load().dlb_nccl_get_unique_id(buf))
    return buf.raw


def shuffle_sequence(seed: int, n: int, epochs: int) -> np.ndarray:
    out = np.zeros((epochs, n), np.int64)
    _capi.check(_capi.load().dlb_shuffle_sequence(seed, n, epochs, out.ctypes.data))
    return out


def dropout_masks(seed: int, in_sizes, keep_rates, n_batches: int, real: str = "f32") -> np.ndarray:
    """The keep masks the training entry points draw after reset_random_generator(seed): [n_batches, mask_size]."""
    dtype = {"f32": np.float32, "f64": np.float64}[real]
    sizes = np.asarray(in_sizes, np.int64); keep = np.asarray(keep_rates, np.float64)
    width = int(sum(int(s) for s, k in zip(sizes, keep) if dtype(k) < 1))
    out = np.zeros((n_batches, width), dtype)
    _capi.check(_capi.load().dlb_dropout_masks(seed, sizes.ctypes.data, keep.ctypes.data, len(sizes), n_batches, np.dtype(dtype).itemsize,
                                               out.ctypes.data))
    return out


def lenet(real: str = "f32", context: Optional[Context] = None) -> Net:
    """BASELINE.json configs[1]: {1,28,28} -> C 5x5x20 ReLU -> P max 2x2 -> N 2880->100 ReLU -> N 100->10 softmax."""
    net = Net(real=real, context=context)
    sz = net.append_layer(CLayer((1, 28, 28), (5, 5), 20, ActivationFunctionId.RELU))
    sz = net.append_layer(PLayer(sz, (2, 2), PoolTypeId.MAX))
    sz = net.append_layer(NLayer(sz, 100, ActivationFunctionId.RELU))
    net.append_layer(NLayer(sz, 10, ActivationFunctionId.SOFTMAX))
    return net


def cifar_deep(real: str = "f32", context: Optional[Context] = None) -> Net:
    """BASELINE.json configs[2]: {3,32,32} -> C 5x5x32 pad 2 ReLU -> P max 2x2 -> C 5x5x64 pad 2 ReLU -> P max 2x2
    -> C 3x3x128 pad 1 ReLU -> P max 2x2 -> N 2048->256 ReLU -> N 256->10 softmax (711 786 parameters)."""
    net = Net(real=real, context=context)
    relu = ActivationFunctionId.RELU
    sz = net.append_layer(CLayer((3, 32, 32), (5, 5), 32, relu, paddings=(0, 2, 2)))
    sz = net.append_layer(PLayer(sz, (2, 2), PoolTypeId.MAX))
    sz = net.append_layer(CLayer(sz, (5, 5), 64, relu, paddings=(0, 2, 2)))
    sz = net.append_layer(PLayer(sz, (2, 2), PoolTypeId.MAX))
    sz = net.append_layer(CLayer(sz, (3, 3), 128, relu, paddings=(0, 1, 1)))
    sz = net.append_layer(PLayer(sz, (2, 2), PoolTypeId.MAX))
    sz = net.append_layer(NLayer(sz, 256, relu))
    net.append_layer(NLayer(sz, 10, ActivationFunctionId.SOFTMAX))
    return net
