#!/usr/bin/env python
"""Generates tests/golden/mnet_{f32,f64}.npz from the REFERENCE itself (oracle/_ref/libdlrefm_*.so = the unmodified MNet<CpuDC> /
RMLayer<CpuDC>): parameters drawn by the reference's own initialisers, outputs of MNet::act, the gradient sum of
MNet::calc_gradient_sum and the parameters after two MNet::learn steps (one of them with the gradient-norm clipping active), for a
two-layer recurrent net.  Run in the build container (needs /root/reference built into oracle/_ref); the fixtures travel."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import binding as ob

DEPTH, DIMS, ACTS, CLIPS = 7, [6, 12, 5], [ob.TANH, ob.SIGMOID], [0.05, 5.0]   # the first layer's threshold is low enough to clip
N = 9

for real in ("f32", "f64"):
    dt = np.float32 if real == "f32" else np.float64
    ref = ob.RefMNet(DEPTH, DIMS, ACTS, CLIPS, real, seed=11)
    p0 = ref.get_params()
    rng = np.random.default_rng(3)
    x = rng.standard_normal((N, DEPTH, DIMS[0])).astype(dt)
    y = rng.random((N, DEPTH, DIMS[-1])).astype(dt)
    out = ref.act(x)
    g_sq = ref.calc_gradient_sum(x, y, ob.SQUARED_ERROR)
    g_ce = ref.calc_gradient_sum(x, y, ob.CROSS_ENTROPY)
    ref.learn(x, y, ob.SQUARED_ERROR, 0.3, 0.0); p1 = ref.get_params()
    ref.learn(x, y, ob.CROSS_ENTROPY, 0.2, 0.01); p2 = ref.get_params()
    path = os.path.join(ROOT, "tests", "golden", f"mnet_{real}.npz")
    np.savez_compressed(path, depth=DEPTH, dims=DIMS, acts=ACTS, clips=CLIPS, x=x, y=y, p0=p0, out=out, g_sq=g_sq, g_ce=g_ce, p1=p1, p2=p2)
    print(path, os.path.getsize(path), "bytes; |g_sq| layer0 =", float(np.sqrt((g_sq[:12 + 72 + 144].astype(np.float64) ** 2).sum())))
