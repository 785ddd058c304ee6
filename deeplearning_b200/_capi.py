"""ctypes binding of libdlb200.so (include/dlb200.h).

There is no CPU fallback: if the library is missing, or no CUDA device is present when a context is created,
the failure is loud (ImportError / DlbError).  Nothing under oracle/ is ever imported from here.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DLB200_LIB") or os.path.join(HERE, "libdlb200.so")   # DLB200_LIB: another build of the same library (development)


class DlbError(RuntimeError):
    """A non-zero status from the C ABI; the message is dlb_last_error()."""


class LayerSpec(C.Structure):
    """dlb_layer_spec (include/dlb200_spec.h)."""
    _fields_ = [(n, C.c_int32) for n in (
        "type", "activation", "in_c", "in_h", "in_w", "win_h", "win_w", "out_c",
        "pad_h", "pad_w", "stride_h", "stride_w", "pool_type", "r0", "r1", "r2")]

    def key(self):
        return tuple(getattr(self, f) for f, _ in self._fields_)


EPOCH_CALLBACK = C.CFUNCTYPE(None, C.c_int64, C.c_double, C.c_void_p)

_lib = None

_OPS = {
    "conv_forward": lambda R: [C.c_void_p, C.POINTER(LayerSpec), C.c_int] + [C.c_void_p] * 6,
    "conv_backward": lambda R: [C.c_void_p, C.POINTER(LayerSpec), C.c_int] + [C.c_void_p] * 9 + [C.c_double],
    "pool_forward": lambda R: [C.c_void_p, C.POINTER(LayerSpec), C.c_int] + [C.c_void_p] * 3,
    "pool_backward": lambda R: [C.c_void_p, C.POINTER(LayerSpec), C.c_int] + [C.c_void_p] * 3,
    "full_forward": lambda R: [C.c_void_p, C.POINTER(LayerSpec), C.c_int] + [C.c_void_p] * 6,
    "full_backward": lambda R: [C.c_void_p, C.POINTER(LayerSpec), C.c_int] + [C.c_void_p] * 9 + [C.c_double],
    "cost_gradient": lambda R: [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 4,
    "sgd_update": lambda R: [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_double],
}

# every symbol include/dlb200.h declares (tests/test_capi_symbols.py checks the header against this list and the .so)
SYMBOLS = [
    "dlb_version", "dlb_last_error", "dlb_context_create", "dlb_context_destroy", "dlb_context_synchronize",
    "dlb_context_stream", "dlb_malloc", "dlb_free", "dlb_memcpy_h2d", "dlb_memcpy_d2h", "dlb_memcpy_d2d", "dlb_memset",
    "dlb_host_alloc_pinned", "dlb_host_free_pinned", "dlb_context_launch_count",
    *[f"dlb_{op}_{sfx}" for op in _OPS for sfx in ("f32", "f64")],
    "dlb_net_create", "dlb_net_destroy", "dlb_net_param_count", "dlb_net_layer_count", "dlb_net_layer_in_size",
    "dlb_net_layer_out_size", "dlb_net_layer_param_offset", "dlb_net_set_params", "dlb_net_get_params",
    "dlb_net_device_params", "dlb_net_device_grads", "dlb_net_set_dataset", "dlb_net_reset_random_generator",
    "dlb_net_learn", "dlb_net_step_indices", "dlb_net_step_host", "dlb_net_forward_backward_host", "dlb_net_read",
    "dlb_net_write_gradient", "dlb_net_eval_host", "dlb_nccl_get_unique_id", "dlb_net_set_distributed",
    "dlb_shuffle_sequence", "dlb_net_train_steps", "dlb_rank_slice", "dlb_event_create", "dlb_event_destroy",
    "dlb_event_record", "dlb_event_elapsed_ms", "dlb_timing_enable", "dlb_timing_report", "dlb_net_describe_plan",
    "dlb_net_enable_graphs", "dlb_net_enable_tensor_cores", "dlb_net_enable_step_kernel", "dlb_matmul_f32", "dlb_matmul_f64", "dlb_net_step_host_async",
    "dlb_net_p2p_export", "dlb_net_set_distributed_p2p", "dlb_net_evaluate_host",
    "dlb_net_set_keep_rate", "dlb_net_mask_size", "dlb_net_set_dropout_masks", "dlb_dropout_masks", "dlb_context_device",
    "dlb_mnet_last_error", "dlb_mnet_create", "dlb_mnet_destroy", "dlb_mnet_param_count", "dlb_mnet_layer_param_offset", "dlb_mnet_set_params",
    "dlb_mnet_get_params", "dlb_mnet_act_host", "dlb_mnet_gradient_sum_host", "dlb_mnet_learn_host", "dlb_mnet_set_batch", "dlb_mnet_learn_resident",
]


class RMLayerSpec(C.Structure):   # dlb_rmlayer_spec (include/dlb200.h)
    _fields_ = [("in_dim", C.c_int32), ("out_dim", C.c_int32), ("activation", C.c_int32), ("reserved", C.c_int32), ("clip_threshold", C.c_double)]


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(deeplearning_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
    lib.dlb_last_error.restype = C.c_char_p
    lib.dlb_context_create.argtypes = [i32, vp, C.POINTER(vp)]
    lib.dlb_context_destroy.argtypes = [vp]
    lib.dlb_context_synchronize.argtypes = [vp]
    lib.dlb_context_stream.argtypes = [vp]; lib.dlb_context_stream.restype = vp
    lib.dlb_context_launch_count.argtypes = [vp]; lib.dlb_context_launch_count.restype = i64
    lib.dlb_malloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    lib.dlb_free.argtypes = [vp, vp]
    for n in ("dlb_memcpy_h2d", "dlb_memcpy_d2h", "dlb_memcpy_d2d"):
        getattr(lib, n).argtypes = [vp, vp, vp, C.c_size_t]
    lib.dlb_memset.argtypes = [vp, vp, i32, C.c_size_t]
    lib.dlb_host_alloc_pinned.argtypes = [C.c_size_t, C.POINTER(vp)]
    lib.dlb_host_free_pinned.argtypes = [vp]
    for op, mk in _OPS.items():
        for sfx in ("f32", "f64"):
            getattr(lib, f"dlb_{op}_{sfx}").argtypes = mk(sfx)
    lib.dlb_net_create.argtypes = [vp, C.POINTER(LayerSpec), i32, i32, C.POINTER(vp)]
    lib.dlb_net_destroy.argtypes = [vp]
    lib.dlb_net_param_count.argtypes = [vp]; lib.dlb_net_param_count.restype = i64
    lib.dlb_net_layer_count.argtypes = [vp]
    for n in ("dlb_net_layer_in_size", "dlb_net_layer_out_size", "dlb_net_layer_param_offset"):
        getattr(lib, n).argtypes = [vp, i32]; getattr(lib, n).restype = i64
    lib.dlb_net_set_params.argtypes = [vp, vp]
    lib.dlb_net_get_params.argtypes = [vp, vp]
    lib.dlb_net_device_params.argtypes = [vp]; lib.dlb_net_device_params.restype = vp
    lib.dlb_net_device_grads.argtypes = [vp]; lib.dlb_net_device_grads.restype = vp
    lib.dlb_net_set_dataset.argtypes = [vp, vp, vp, i64, i32]
    lib.dlb_net_reset_random_generator.argtypes = [vp, C.c_uint]
    lib.dlb_net_learn.argtypes = [vp, i64, i64, dbl, i32, dbl, vp, vp]
    lib.dlb_net_step_indices.argtypes = [vp, vp, i64, dbl, i32, dbl]
    lib.dlb_net_step_host.argtypes = [vp, vp, vp, i64, dbl, i32, dbl, C.POINTER(dbl)]
    lib.dlb_net_step_host_async.argtypes = [vp, vp, vp, i64, dbl, i32, dbl, vp]
    lib.dlb_net_forward_backward_host.argtypes = [vp, vp, vp, i64, i32, dbl]
    lib.dlb_net_read.argtypes = [vp, i32, i32, vp, i64]
    lib.dlb_net_write_gradient.argtypes = [vp, vp]
    lib.dlb_net_eval_host.argtypes = [vp, vp, i64, vp, vp]
    lib.dlb_nccl_get_unique_id.argtypes = [vp]
    lib.dlb_net_set_distributed.argtypes = [vp, i32, i32, vp]
    lib.dlb_shuffle_sequence.argtypes = [C.c_uint, i64, i64, vp]
    lib.dlb_net_train_steps.argtypes = [vp, i64, i64, dbl, i32, dbl, C.POINTER(i64)]
    lib.dlb_rank_slice.argtypes = [i64, i32, i32, C.POINTER(i64), C.POINTER(i64)]
    lib.dlb_event_create.argtypes = [C.POINTER(vp)]
    lib.dlb_event_destroy.argtypes = [vp]
    lib.dlb_event_record.argtypes = [vp, vp]
    lib.dlb_event_elapsed_ms.argtypes = [vp, vp, C.POINTER(dbl)]
    lib.dlb_timing_enable.argtypes = [i32]
    lib.dlb_timing_report.argtypes = [C.c_char_p, i64]
    lib.dlb_net_describe_plan.argtypes = [vp, C.c_char_p, i64]
    lib.dlb_net_evaluate_host.argtypes = [vp, vp, vp, i64, i32, dbl, C.POINTER(dbl), C.POINTER(dbl)]
    lib.dlb_net_p2p_export.argtypes = [vp, vp]
    lib.dlb_net_set_distributed_p2p.argtypes = [vp, i32, i32, vp]
    lib.dlb_net_enable_graphs.argtypes = [vp, i32]
    lib.dlb_net_enable_tensor_cores.argtypes = [vp, i32]
    lib.dlb_net_enable_step_kernel.argtypes = [vp, i32]
    lib.dlb_net_set_keep_rate.argtypes = [vp, i32, dbl]
    lib.dlb_net_mask_size.argtypes = [vp]; lib.dlb_net_mask_size.restype = i64
    lib.dlb_net_set_dropout_masks.argtypes = [vp, vp]
    lib.dlb_dropout_masks.argtypes = [C.c_uint, vp, vp, i32, i64, i32, vp]
    lib.dlb_matmul_f32.argtypes = [vp, i32, i32, i32, vp, i64, i32, vp, i64, i32, vp, i32]
    lib.dlb_matmul_f64.argtypes = [vp, i32, i32, i32, vp, i64, i32, vp, i64, i32, vp]
    lib.dlb_context_device.argtypes = [vp]
    lib.dlb_mnet_last_error.restype = C.c_char_p
    lib.dlb_mnet_create.argtypes = [vp, i32, C.POINTER(RMLayerSpec), i32, i32, C.POINTER(vp)]
    lib.dlb_mnet_destroy.argtypes = [vp]
    lib.dlb_mnet_param_count.argtypes = [vp]; lib.dlb_mnet_param_count.restype = i64
    lib.dlb_mnet_layer_param_offset.argtypes = [vp, i32]; lib.dlb_mnet_layer_param_offset.restype = i64
    lib.dlb_mnet_set_params.argtypes = [vp, vp]
    lib.dlb_mnet_get_params.argtypes = [vp, vp]
    lib.dlb_mnet_act_host.argtypes = [vp, vp, i64, vp]
    lib.dlb_mnet_gradient_sum_host.argtypes = [vp, vp, vp, i64, i32, vp]
    lib.dlb_mnet_learn_host.argtypes = [vp, vp, vp, i64, i32, dbl, dbl]
    lib.dlb_mnet_set_batch.argtypes = [vp, vp, vp, i64]
    lib.dlb_mnet_learn_resident.argtypes = [vp, i32, dbl, dbl, i64]
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise DlbError(load().dlb_last_error().decode(errors="replace"))


def check_mnet(rc: int) -> None:
    if rc != 0:
        raise DlbError(load().dlb_mnet_last_error().decode(errors="replace"))
