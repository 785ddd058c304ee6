#!/usr/bin/env python
"""Op-level timings against their rooflines (one JSON line per op; CUDA events on the context's stream, warm, several buffer sets
in rotation so that no launch finds its inputs in L2):
  * the pooling kernels (layers.cuh) on the three PLayers of CIFAR-deep at mini-batch 1024 (BASELINE configs[2]), float and double:
    algorithmic bytes = input + output + index map (forward), delta + map + input gradient (backward), against the measured HBM peak;
  * the Real = double contractions of the LeNet NLayer at mini-batch 8192 and 256 (gemm.cuh) against the measured DFMA peak.
usage: bench_ops.py [pool] [gemm64]"""
import ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deeplearning_b200 as dl
from deeplearning_b200 import _capi


def peaks():
    hbm, dfma = 6546.9, 36.3
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    try:
        dfma = json.load(open(os.path.join(ROOT, "profiles", "peaks_r2.json")))["dfma_tflops"]
    except Exception:
        pass
    return float(hbm), float(dfma)


class Dev:
    def __init__(self):
        self.ctx = dl.Context(0); self.lib = self.ctx.lib; self.ptrs = []

    def alloc(self, nbytes, fill=None):
        p = C.c_void_p()
        _capi.check(self.lib.dlb_malloc(self.ctx.h, max(nbytes, 1), C.byref(p)))
        if fill is not None:
            a = np.ascontiguousarray(fill)
            _capi.check(self.lib.dlb_memcpy_h2d(self.ctx.h, p, a.ctypes.data, a.nbytes))
        self.ptrs.append(p)
        return p

    def free_all(self):
        for p in self.ptrs:
            self.lib.dlb_free(self.ctx.h, p)
        self.ptrs = []


def timed(dev, calls, reps):
    reps = int(os.environ.get("BENCH_OPS_REPS", reps))   # 1 under ncu
    """calls: list of zero-argument launches (one per buffer set); returns microseconds per launch"""
    for f in calls:
        f()
    dev.ctx.synchronize()
    e0, e1 = dl.Event(dev.ctx), dl.Event(dev.ctx)
    e0.record()
    for r in range(reps):
        calls[r % len(calls)]()
    e1.record()
    dev.ctx.synchronize()
    return 1e3 * e0.elapsed_ms(e1) / reps


def bench_pool(dev, hbm):
    B = 1024
    for real, dt in (("f32", np.float32), ("f64", np.float64)):
        for (c, h, w) in ((32, 32, 32), (64, 16, 16), (128, 8, 8)):
            spec = dl.PLayer((c, h, w), (2, 2), dl.PoolTypeId.MAX)
            n_in, n_out = B * c * h * w, B * c * (h // 2) * (w // 2)
            es = np.dtype(dt).itemsize
            sets = max(2, int(400e6 // (n_in * es)) + 1)
            rng = np.random.default_rng(0)
            x = rng.random(n_in, dtype=np.float32).astype(dt)
            bufs = [(dev.alloc(n_in * es, x), dev.alloc(n_out * es), dev.alloc(n_out * 4), dev.alloc(n_in * es)) for _ in range(sets)]
            fwd = getattr(dev.lib, f"dlb_pool_forward_{real}"); bwd = getattr(dev.lib, f"dlb_pool_backward_{real}")
            us_f = timed(dev, [lambda b=b: _capi.check(fwd(dev.ctx.h, C.byref(spec), B, b[0], b[1], b[2])) for b in bufs], 40)
            us_b = timed(dev, [lambda b=b: _capi.check(bwd(dev.ctx.h, C.byref(spec), B, b[1], b[2], b[3])) for b in bufs], 40)
            bytes_f = n_in * es + n_out * (es + 4); bytes_b = bytes_f
            for name, us, by in (("pool_forward", us_f, bytes_f), ("pool_backward", us_b, bytes_b)):
                gbps = by / us / 1e3
                print(json.dumps({"op": name, "real": real, "batch": B, "in_size": [c, h, w], "window": [2, 2], "us": round(us, 2),
                                  "algorithmic_bytes": by, "GBps": round(gbps, 1), "hbm_peak_GBps": hbm, "frac": round(gbps / hbm, 3),
                                  "buffer_sets_in_rotation": sets}), flush=True)
            dev.free_all()


def bench_gemm64(dev, dfma):
    rng = np.random.default_rng(1)
    for B in (8192, 256):
        # (name, M, N, K, a_transposed, b_transposed): NLayer 2880 -> 100 forward, weight gradient, input gradient (dlb200.cu)
        for name, M, N, K, at, bt in (("full_fwd", B, 100, 2880, 0, 0), ("full_dW", 100, 2880, B, 1, 1), ("full_dX", B, 2880, 100, 0, 1)):
            A = rng.standard_normal((K, M) if at else (M, K)); Bm = rng.standard_normal((K, N) if bt else (N, K))
            sets = 3
            bufs = [(dev.alloc(A.nbytes, A), dev.alloc(Bm.nbytes, Bm), dev.alloc(M * N * 8)) for _ in range(sets)]
            f = dev.lib.dlb_matmul_f64
            us = timed(dev, [lambda b=b: _capi.check(f(dev.ctx.h, M, N, K, b[0], M if at else K, at, b[1], N if bt else K, bt, b[2])) for b in bufs], 30)
            tf = 2.0 * M * N * K / us / 1e6
            print(json.dumps({"op": "matmul_f64:" + name, "M": M, "N": N, "K": K, "us": round(us, 2), "TFLOPs": round(tf, 2),
                              "dfma_peak_TFLOPs": dfma, "frac": round(tf / dfma, 3)}), flush=True)
            dev.free_all()


if __name__ == "__main__":
    what = sys.argv[1:] or ["pool", "gemm64"]
    hbm, dfma = peaks()
    dev = Dev()
    if "pool" in what:
        bench_pool(dev, hbm)
    if "gemm64" in what:
        bench_gemm64(dev, dfma)
