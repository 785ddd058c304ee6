"""Host mirror of the reference's recurrent stack over the C ABI (include/dlb200.h, csrc/mnet.cu): MNet of RMLayer
(NeuralNet/MNet.h:31-160, NeuralNet/RMLayer.h:30-160) - same method names, argument meaning and error behaviour; sequences are
numpy arrays [n][depth][dim] of the net's Real."""
import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _capi
from .net import ActivationFunctionId, Context, CostFunctionId


def RMLayer(in_sub_dim: int, out_sub_dim: int, func_id: ActivationFunctionId = ActivationFunctionId.TANH,
            gradient_clip_threshold: float = 5.0) -> _capi.RMLayerSpec:
    """RMLayer(rec_depth, in_sub_dim, out_sub_dim, init_strategy, func_id = TANH, gradient_clip_threshold = 5) (RMLayer.h:131-133); the
    recurrence depth belongs to the net here (every layer of an MNet has the same one, RMLayer.inl:118-119)."""
    return _capi.RMLayerSpec(int(in_sub_dim), int(out_sub_dim), int(func_id), 0, float(gradient_clip_threshold))


class MNet:
    """MNet<GpuDC>: append_layer'ed RMLayers, act / calc_gradient_sum / learn on whole mini-batches of sequences."""

    def __init__(self, rec_depth: int, layers: Sequence[_capi.RMLayerSpec], real: str = "f32", context: Optional[Context] = None):
        self.lib = _capi.load()
        self.ctx = context or Context()
        self.real = real
        self.dtype = np.float32 if real == "f32" else np.float64
        self.rec_depth = int(rec_depth)
        self.specs = list(layers)
        arr = (_capi.RMLayerSpec * len(self.specs))(*self.specs)
        h = C.c_void_p()
        _capi.check_mnet(self.lib.dlb_mnet_create(self.ctx.h, self.rec_depth, arr, len(self.specs), 4 if real == "f32" else 8, C.byref(h)))
        self.h = h
        self.param_count = int(self.lib.dlb_mnet_param_count(self.h))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.dlb_mnet_destroy(self.h); self.h = None
        except Exception:
            pass

    def in_dim(self) -> int: return self.specs[0].in_dim
    def out_dim(self) -> int: return self.specs[-1].out_dim
    def layer_param_offset(self, layer: int) -> int: return int(self.lib.dlb_mnet_layer_param_offset(self.h, layer))

    def init_params(self, seed: int = 0, scale: float = 0.3) -> np.ndarray:
        rng = np.random.default_rng(seed)
        p = (rng.standard_normal(self.param_count) * scale).astype(self.dtype)
        self.set_params(p)
        return p

    def set_params(self, params: np.ndarray) -> None:
        p = np.ascontiguousarray(params, self.dtype)
        if p.size != self.param_count: raise ValueError("Invalid input")
        _capi.check_mnet(self.lib.dlb_mnet_set_params(self.h, p.ctypes.data))

    def get_params(self) -> np.ndarray:
        p = np.empty(self.param_count, self.dtype)
        _capi.check_mnet(self.lib.dlb_mnet_get_params(self.h, p.ctypes.data))
        return p

    def _seq(self, a, dim) -> np.ndarray:
        a = np.ascontiguousarray(a, self.dtype)
        if a.ndim != 3 or a.shape[1] != self.rec_depth or a.shape[2] != dim: raise ValueError("Invalid input data")   # RMLayer.inl:171-172
        return a

    def act(self, inputs) -> np.ndarray:
        x = self._seq(inputs, self.in_dim())
        out = np.empty((x.shape[0], self.rec_depth, self.out_dim()), self.dtype)
        _capi.check_mnet(self.lib.dlb_mnet_act_host(self.h, x.ctypes.data, x.shape[0], out.ctypes.data))
        return out

    def calc_gradient_sum(self, inputs, refs, cost_func_id: CostFunctionId) -> np.ndarray:
        x, y = self._seq(inputs, self.in_dim()), self._seq(refs, self.out_dim())
        if x.shape[0] != y.shape[0]: raise ValueError("Inconsistent input data.")   # MNet.inl:113-114
        g = np.empty(self.param_count, self.dtype)
        _capi.check_mnet(self.lib.dlb_mnet_gradient_sum_host(self.h, x.ctypes.data, y.ctypes.data, x.shape[0], int(cost_func_id), g.ctypes.data))
        return g

    def learn(self, inputs, refs, cost_func_id: CostFunctionId, learning_rate: float, reg_factor: float = 0.0) -> None:
        x, y = self._seq(inputs, self.in_dim()), self._seq(refs, self.out_dim())
        if x.shape[0] != y.shape[0]: raise ValueError("Inconsistent input data.")
        _capi.check_mnet(self.lib.dlb_mnet_learn_host(self.h, x.ctypes.data, y.ctypes.data, x.shape[0], int(cost_func_id), float(learning_rate), float(reg_factor)))

    def set_batch(self, inputs, refs) -> None:
        x, y = self._seq(inputs, self.in_dim()), self._seq(refs, self.out_dim())
        _capi.check_mnet(self.lib.dlb_mnet_set_batch(self.h, x.ctypes.data, y.ctypes.data, x.shape[0]))

    def learn_resident(self, cost_func_id: CostFunctionId, learning_rate: float, reg_factor: float = 0.0, n_steps: int = 1) -> None:
        _capi.check_mnet(self.lib.dlb_mnet_learn_resident(self.h, int(cost_func_id), float(learning_rate), float(reg_factor), int(n_steps)))
