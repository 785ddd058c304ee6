#!/usr/bin/env python
"""Development: the step kernel at small / odd mini-batches against the oracle, per layer and per step count."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deeplearning_b200 as dl
from oracle import binding as ob
from tests import helpers as H

def seg_err(gpu, got, want):
    out = []
    for l in range(4):
        a = gpu.layer_param_offset(l); b = gpu.layer_param_offset(l + 1) if l < 3 else gpu.param_count
        if b > a:
            d = np.abs(got[a:b].astype(np.float64) - want[a:b]); out.append("L%d %.2e" % (l, d.max() / (np.abs(want[a:b]).max() + 1e-30)))
    return " ".join(out)

specs = ob.lenet_specs()
for batch, steps in [(37, 1), (37, 3), (39, 1), (64, 1), (64, 2), (100, 2), (128, 2), (16, 2), (1, 3)]:
    orc, gpu = H.make_pair(specs, "f32", seed=3, tc_mode=1)
    gpu.enable_step_kernel(True)
    n = batch * steps
    x, y = H.synthetic_batch(orc, n, "f32", seed=4, lo=0.0, hi=1.0)
    perm = np.arange(n, dtype=np.int64)[None, :]
    orc.learn(x, y, batch, 1, 0.05, ob.CROSS_ENTROPY, 0.2, threads=1, perms=perm)
    gpu.set_dataset(x, y)
    # one launch of `steps` steps over the identity order
    gpu.reset_random_generator(0)
    # step_indices runs single steps; train_steps would shuffle - use step_indices in a loop and train via learn-like multi-step through train_order if there is one
    for i in range(steps):
        gpu.step_indices(np.arange(i * batch, (i + 1) * batch, dtype=np.int64), 0.05, dl.CostFunctionId.CROSS_ENTROPY, reg_factor=0.2 / n)
    print("batch %4d x %d single-step launches: %s" % (batch, steps, seg_err(gpu, gpu.get_params(), orc.get_params())))
