"""BASELINE.json configs[2] (CIFAR-shaped deep CNN, mini-batch 1024) on one B200: training samples/s with the contractions on
the FMA path (--tc 0) and on the tensor cores (--tc 1), plus the per-kernel table of each.  Not the bench line (bench.py
measures configs[1]); a measured parity-test case, results go to profiles/."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deeplearning_b200 as dl  # noqa: E402


def run(real, tc, batch, steps, warmup, n_items):
    ctx = dl.Context(device=0)
    net = dl.cifar_deep(real=real, context=ctx)
    net.init_params(seed=0)
    if real == "f32":
        net.enable_tensor_cores(tc)
    rng = np.random.default_rng(7)
    x = rng.uniform(0, 1, (n_items, 3 * 32 * 32)).astype(net.dtype)
    y = np.zeros((n_items, 10), net.dtype); y[np.arange(n_items), np.arange(n_items) % 10] = 1
    net.set_dataset(x, y)
    net.reset_random_generator(0)
    cost = dl.CostFunctionId.CROSS_ENTROPY
    e0, e1 = dl.Event(ctx), dl.Event(ctx)
    net.train_steps(batch, warmup, 0.01, cost, 0.1)
    ctx.synchronize()
    l0 = ctx.launch_count()
    e0.record(); items = net.train_steps(batch, steps, 0.01, cost, 0.1); e1.record()
    ms = e0.elapsed_ms(e1)
    launches = ctx.launch_count() - l0
    dl.timing_enable(True)
    net.train_steps(batch, 4, 0.01, cost, 0.1)
    rep = dl.timing_report()
    dl.timing_enable(False)
    tot = sum(v[1] for v in rep.values()) or 1.0
    kernels = [{"kernel": k, "n_per_step": n / 4, "us_per_launch": 1e3 * t / n, "share": t / tot}
               for k, (n, t) in sorted(rep.items(), key=lambda kv: -kv[1][1])[:16]]
    p = net.get_params()
    return {"workload": "CIFAR-deep (BASELINE configs[2])", "dtype": real, "tc_mode": tc, "batch": batch, "steps": steps,
            "value": items / (ms / 1e3), "unit": "samples/s", "ms_per_step": ms / steps, "launches_per_step": launches / steps,
            "flop_per_sample": 119946240, "tflops": 119946240 * items / (ms / 1e3) / 1e12,
            "params_finite": bool(np.all(np.isfinite(p))), "plan": net.describe_plan(), "kernels": kernels}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--real", default="f32")
    ap.add_argument("--tc", type=int, nargs="*", default=[0, 1])
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--items", type=int, default=8192)
    a = ap.parse_args()
    for tc in a.tc:
        print(json.dumps(run(a.real, tc, a.batch, a.steps, a.warmup, a.items)), flush=True)
