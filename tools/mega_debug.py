#!/usr/bin/env python
"""Development check of the step megakernel (mega.cu) on a GPU box: one LeNet step through dlb_net_step_indices, every
buffer the kernel leaves behind compared with the CPU oracle, buffer by buffer with error figures (not asserts), for the
one-launch-per-phase mode (DLB_MEGA_SPLIT=1) and the single cooperative launch.

    python tools/mega_debug.py [batch]
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel_err(got, want):
    got = np.asarray(got, np.float64); want = np.asarray(want, np.float64)
    scale = np.max(np.abs(want)) if want.size else 1.0
    err = np.abs(got - want) / (np.abs(want) + scale + 1e-30)
    i = int(np.argmax(err)) if err.size else 0
    return float(err.max()) if err.size else 0.0, i


def run(batch: int):
    import deeplearning_b200 as dl
    from oracle import binding as ob
    from tests import helpers as H

    real = "f32"
    specs = ob.lenet_specs()
    orc, gpu = H.make_pair(specs, real, seed=1, tc_mode=1)
    print("plan:", gpu.describe_plan())
    x, y = H.synthetic_batch(orc, batch, real, seed=2, lo=0.0, hi=1.0)
    want = H.oracle_batch_trace(orc, x, y, ob.CROSS_ENTROPY)
    p0 = gpu.get_params()
    gpu.set_dataset(x, y)
    lr, reg = 0.05, 0.001
    gpu.step_indices(np.arange(batch, dtype=np.int64), lr, dl.CostFunctionId.CROSS_ENTROPY, reg_factor=reg)
    gpu._last_batch = batch
    out_off = H.layer_offsets(orc.out_sizes); in_off = H.layer_offsets(orc.in_sizes)
    pooled = gpu.read(dl.BUF_OUTPUT, 1)
    e, i = rel_err(pooled, want["acts"][:, out_off[1]:out_off[1] + 2880]); print(f"pooled           max rel err {e:.3e} at {i}")
    dpool = gpu.read(dl.BUF_INPUT_GRAD, 2)
    e, i = rel_err(dpool, want["in_grads"][:, in_off[2]:in_off[2] + 2880]); print(f"dpool            max rel err {e:.3e} at {i} (sample {i // 2880}, k {i % 2880})")
    # the megakernel's scratch (development buffer ids of dlb_net_read)
    def raw(what, count):
        import ctypes
        a = np.empty(count, np.float32)
        from deeplearning_b200 import _capi
        _capi.check(gpu.lib.dlb_net_read(gpu._handle(), what, 0, a.ctypes.data, a.nbytes))
        return a
    F, H1 = 20, 100
    z1part = raw(101, F * batch * H1).reshape(F, batch, H1)
    pooled_want = want["acts"][:, out_off[1]:out_off[1] + 2880].astype(np.float64)
    W1 = p0[12120:300120].reshape(100, 2880).astype(np.float64)
    z1_want = np.stack([pooled_want[:, f * 144:(f + 1) * 144] @ W1[:, f * 144:(f + 1) * 144].T for f in range(F)])
    e, i = rel_err(z1part, z1_want); print(f"z1part           max rel err {e:.3e} at {i} (f {i // (batch * H1)}, b {(i // H1) % batch}, n {i % H1}) got {z1part.ravel()[i]:.6g} want {z1_want.ravel()[i]:.6g}")
    bad = np.abs(z1part - z1_want) > 1e-4 * (np.abs(z1_want) + np.abs(z1_want).max())
    print("  bad fraction by filter:", np.round(bad.mean(axis=(1, 2)), 2))
    print("  bad fraction by n (blocks of 10):", np.round(bad.mean(axis=(0, 1)).reshape(10, 10).mean(axis=1), 2))
    print("  bad fraction by sample (blocks of 8):", np.round(bad.mean(axis=(0, 2)).reshape(-1, 8).mean(axis=1), 2))
    if bad.any():
        f_, b_, n_ = [int(v[0]) for v in np.nonzero(bad)]
        print(f"  first bad: f {f_} b {b_} n {n_}: got {z1part[f_, b_, n_]:.6g} want {z1_want[f_, b_, n_]:.6g}")
        # is the value some other product?  compare with the full table for this filter
        got = float(z1part[f_, b_, n_])
        tab = z1_want[f_]
        j = np.argmin(np.abs(tab - got)); print(f"  closest value of the same filter's table: b {j // H1} n {j % H1} ({tab.ravel()[j]:.6g})")
    z1 = z1part.astype(np.float64).sum(axis=0) + p0[12020:12120]
    e, i = rel_err(np.maximum(z1, 0), want["acts"][:, out_off[2]:out_off[2] + 100]); print(f"a1 from z1part   max rel err {e:.3e} at {i}")
    g1 = raw(102, batch * 128).reshape(batch, 128)
    # g1 = dC/dz of the hidden layer = (input gradient of the head) * relu'
    d1 = (want["acts"][:, out_off[2]:out_off[2] + 100] > 0)
    g1_want = want["in_grads"][:, in_off[3]:in_off[3] + 100] * d1
    e, i = rel_err(g1[:, :100], g1_want); print(f"g1               max rel err {e:.3e} at {i} (b {i // 100}, n {i % 100}); pad columns max {np.abs(g1[:, 100:]).max():.3g}")
    dw1part = raw(103, 7 * 100 * 2880).reshape(7, 100, 2880)
    dW1_want = g1_want.astype(np.float64).T @ pooled_want
    e, i = rel_err(dw1part.astype(np.float64).sum(axis=0), dW1_want); print(f"sum of dw1part   max rel err {e:.3e} at {i} (n {i // 2880}, k {i % 2880})")
    g = gpu.read(dl.BUF_GRADIENT)
    names = ["conv (bias | filters)", "pool", "hidden NLayer (bias | W)", "head (bias | W)"]
    for l in range(4):
        a = gpu.layer_param_offset(l); b = gpu.layer_param_offset(l + 1) if l < 3 else gpu.param_count
        if b > a:
            e, i = rel_err(g[a:b], want["grads"][a:b]); print(f"grad layer {l} {names[l]:26s} max rel err {e:.3e} at {i} (got {g[a + i]:.6g} want {want['grads'][a + i]:.6g})")
    e, i = rel_err(g[:11520], want["grads"][:11520]); print(f"  conv bias grad   {e:.3e} at {i}")
    e, i = rel_err(g[11520:12020], want["grads"][11520:12020]); print(f"  conv filter grad {e:.3e} at {i}")
    e, i = rel_err(g[12020:12120], want["grads"][12020:12120]); print(f"  fc1 bias grad    {e:.3e} at {i}")
    e, i = rel_err(g[12120:300120], want["grads"][12120:300120]); print(f"  fc1 weight grad  {e:.3e} at {i} (n {i // 2880}, k {i % 2880})")
    perm = np.arange(batch, dtype=np.int64)[None, :]
    orc.learn(x, y, batch, 1, lr, ob.CROSS_ENTROPY, reg * batch, threads=1, perms=perm)
    e, i = rel_err(gpu.get_params(), orc.get_params()); print(f"params after step max rel err {e:.3e} at {i}")
    e, i = rel_err(gpu.get_params() - p0, orc.get_params() - p0); print(f"param update      max rel err {e:.3e} at {i}")
    cost = gpu.step_host(x, y, lr, dl.CostFunctionId.CROSS_ENTROPY, reg_factor=reg)
    orc2_cost = None
    outs = orc.eval(x, want_outputs=True)[1]
    orc2_cost = sum(orc.cost(outs[k], y[k]) for k in range(batch))
    print(f"batch cost via step_host {cost:.6f} (oracle on the updated net {orc2_cost:.6f})")


def profile(batch: int, steps: int = 20):
    """DLB_MEGA_PROF=1: per-CTA clock64 stamps at the phase boundaries of the last step."""
    import deeplearning_b200 as dl
    from deeplearning_b200 import _capi
    net = dl.lenet("f32")
    net.init_params(seed=0)
    rng = np.random.default_rng(1)
    n = max(4096, 4 * batch)
    x = rng.random((n, 784), dtype=np.float32); y = np.zeros((n, 10), np.float32); y[np.arange(n), np.arange(n) % 10] = 1
    net.set_dataset(x, y)
    net.reset_random_generator(0)
    ctx = net.ctx
    net.train_steps(batch, 5, 0.01, dl.CostFunctionId.CROSS_ENTROPY, 0.1)
    ctx.synchronize()
    e0, e1 = dl.Event(ctx), dl.Event(ctx)
    e0.record(); net.train_steps(batch, steps, 0.01, dl.CostFunctionId.CROSS_ENTROPY, 0.1); e1.record()
    ms = e0.elapsed_ms(e1)
    print(f"batch {batch}: {1e3 * ms / steps:.2f} us/step over {steps} steps")
    a = np.zeros(148 * 40, np.int64)
    _capi.check(net.lib.dlb_net_read(net._handle(), 107, 0, a.ctypes.data, a.nbytes))
    a = a.reshape(148, 40)
    a = a[a[:, 1] != 0]
    names = ["setup", "conv_fwd", "barrier", "fc1_fwd", "barrier", "head", "barrier", "fc1_bwd(+dW)", "barrier wait", "conv_bwd", "barrier", "update"]
    d = np.diff(a[:, :13], axis=1) / 1965.0
    print(f"  {len(a)} CTAs; per-CTA intervals in us (mean / min / max), SM clock assumed 1965 MHz")
    for i, nm in enumerate(names):
        print(f"  {nm:14s} {d[:, i].mean():7.2f} {d[:, i].min():7.2f} {d[:, i].max():7.2f}")
    print(f"  total          {((a[:, 12] - a[:, 0]) / 1965.0).mean():7.2f}")
    c0 = a[0]
    def rel(slot, base): return (c0[slot] - c0[base]) / 1965.0
    print("  CTA 0 inside conv_fwd (us after phase start): copies landed %.2f, synced %.2f, FFMA loop done %.2f, phase end %.2f" % (rel(13, 1), rel(14, 1), rel(15, 1), rel(2, 1)))
    print("  CTA 0 inside conv_bwd (us after phase start): staged %.2f, landed+synced %.2f, FFMA loop done %.2f, bias sums done %.2f, phase end %.2f" %
          (rel(28, 9), rel(29, 9), rel(30, 9), rel(31, 9), rel(10, 9)))
    print("  CTA 0 inside head (us after phase start): operands staged %.2f, synced %.2f, forward+cost done %.2f, gradient+partials done %.2f, phase end %.2f" %
          (rel(16, 5), rel(17, 5), rel(18, 5), rel(19, 5), rel(6, 5)))
    if c0[35]: print("  head run a second time (warm instruction cache): %.2f us" % ((c0[35] - c0[34]) / 1965.0))
    print("  CTA 0 inside fc1_bwd (us after phase start): start %.2f, operands landed %.2f, mma issue %.2f, dX mma done %.2f, dX stored %.2f, arrived %.2f, dW issued %.2f, dW done %.2f, phase end %.2f" %
          tuple(rel(k, 7) for k in (20, 21, 22, 23, 24, 25, 26, 27, 8)))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "prof":
        os.environ["DLB_MEGA_PROF"] = "1"
        for b in [int(v) for v in sys.argv[2:]] or [256]:
            profile(b)
        sys.exit(0)
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    if os.environ.get("MEGA_DEBUG_CHILD"):
        run(batch)
    else:
        for split in ("1", "0"):
            print(f"===== DLB_MEGA_SPLIT={split} batch {batch}")
            env = dict(os.environ, MEGA_DEBUG_CHILD="1", DLB_MEGA_SPLIT=split)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), str(batch)], env=env, capture_output=True, text=True, timeout=600)
            print(r.stdout[-4000:]); print(r.stderr[-3000:])
