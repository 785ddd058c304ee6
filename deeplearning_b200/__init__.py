"""deeplearning_b200 — B200-native (sm_100a) mini-batch training path of dragunovdenis/DeepLearning.

Python host mirror of the reference's Net / CLayer / PLayer / NLayer interface over the C ABI of libdlb200.so
(include/dlb200.h).  The CUDA library is mandatory: there is no CPU fallback.
"""
from ._capi import DlbError, LayerSpec  # noqa: F401
from .net import (  # noqa: F401
    ActivationFunctionId, CLayer, Context, CostFunctionId, LayerTypeId, NLayer, Net, PLayer, PoolTypeId,
    BUF_COST_GRAD, BUF_DERIVATIVES, BUF_GRADIENT, BUF_INDEX_MAP, BUF_INPUT_GRAD, BUF_OUTPUT,
    Event, cifar_deep, dropout_masks, lenet, nccl_unique_id, out_size, pinned_array, rank_slice, shuffle_sequence, timing_enable, timing_report,
)
from .mnet import MNet, RMLayer  # noqa: F401
