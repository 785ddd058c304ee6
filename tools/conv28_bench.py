#!/usr/bin/env python
"""Development: LeNet step time and per-kernel table with the conv28 kernels off / on at several mini-batches."""
import os, sys, subprocess, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def run(batches):
    import deeplearning_b200 as dl
    net = dl.lenet("f32"); net.init_params(seed=0)
    rng = np.random.default_rng(1)
    n = 4 * max(batches)
    x = rng.random((n, 784), dtype=np.float32); y = np.zeros((n, 10), np.float32); y[np.arange(n), np.arange(n) % 10] = 1
    net.set_dataset(x, y); net.reset_random_generator(0)
    ctx = net.ctx
    cid = dl.CostFunctionId.CROSS_ENTROPY
    for b in batches:
        steps = 200 if b <= 1024 else 30
        net.train_steps(b, 5, 0.01, cid, 0.1); ctx.synchronize()
        e0, e1 = dl.Event(ctx), dl.Event(ctx)
        e0.record(); net.train_steps(b, steps, 0.01, cid, 0.1); e1.record()
        us = 1e3 * e0.elapsed_ms(e1) / steps
        dl.timing_enable(True); net.train_steps(b, 8, 0.01, cid, 0.1); rep = dl.timing_report(); dl.timing_enable(False)
        conv = {k.split(":")[1] if ":" in k else k: 1e3 * t / c for k, (c, t) in rep.items() if "conv" in k}
        print("DLB_CONV28=%s batch %5d: %8.2f us/step  %s" % (os.environ.get("DLB_CONV28", "auto"), b, us, {k: round(v, 1) for k, v in conv.items()}), flush=True)

if __name__ == "__main__":
    if os.environ.get("CHILD"):
        run([int(v) for v in sys.argv[1:]])
    else:
        for mode in ("0", "1"):
            r = subprocess.run([sys.executable, os.path.abspath(__file__)] + (sys.argv[1:] or ["256", "512", "1024", "2048", "8192"]),
                               env=dict(os.environ, CHILD="1", DLB_CONV28=mode), capture_output=True, text=True, timeout=900)
            print(r.stdout[-3000:]); print(r.stderr[-1500:])
