#!/usr/bin/env python
"""bench.py — training samples/s (forward + backward + update) of the LeNet-style CNN, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--real f32|f64]

A "step" is one mini-batch of Net::learn (Net.inl:226-317): forward, cost derivative, backward, (all-reduce), update.

N = 1   workload = BASELINE.json configs[1]: LeNet (C 5x5x20 ReLU -> P max 2x2 -> N 100 ReLU -> N 10 softmax), B = 256.
N > 1   workload = configs[4]: the same net, global mini-batch 8192 sharded over the ranks, per-step NCCL all-reduce of
        the flat gradient (strong scaling: total work per step is fixed).

`value`     device-timed (CUDA events on the library's stream) over K steps with the data set resident in HBM
            (65 536 items = 205 MB of inputs, larger than the 126 MB L2; each step gathers a fresh slice of the epoch
            permutation), max over ranks.
`e2e`       the same metric through the host-buffer entry point dlb_net_step_host: per step the batch is copied from
            pinned host memory, and the batch cost is read back (8 bytes) - copies inside the timed region.
`roofline`  the kernel with the largest share of the step, timed per launch with CUDA events in a separate
            instrumented pass; algorithmic bytes per launch as defined in DESIGN.md; peak = MEASURED_PEAKS.json hbm_gbs.
`cpu_baseline` / `--impl reference`  the reference's own Net<CpuDC>::learn (oracle/_ref, all host threads) on a bounded
            sample of the same workload.  oracle/ is used here as the baseline being timed, never by the GPU arm.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ITEMS = 65536
LR, LAMBDA = 0.01, 0.1
CROSS_ENTROPY = 2
FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback


def synthetic_data(n, in_n, out_n, dtype, seed=7):
    """MNIST-shaped synthetic items: inputs U[0,1), one-hot labels i mod 10."""
    rng = np.random.default_rng(seed)
    x = rng.random((n, in_n), dtype=np.float32).astype(dtype)
    y = np.zeros((n, out_n), dtype)
    y[np.arange(n), np.arange(n) % out_n] = 1
    return x, y


# ------------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        self.windows = []
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def window(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.f.flush(); self.f.seek(0)
        rows = []
        for line in self.f.read().splitlines():
            p = [t.strip() for t in line.split(",")]
            if len(p) < 9:
                continue
            try:
                ts = time.mktime(time.strptime(p[0].split(".")[0], "%Y/%m/%d %H:%M:%S")) + float("0." + p[0].split(".")[1])
                rows.append((ts, float(p[2]), float(p[3]), float(p[4]), p[5:9]))
            except (ValueError, IndexError):
                continue
        os.unlink(self.f.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        inside = [r for r in rows if any(t0 - 0.05 <= r[0] <= t1 + 0.05 for t0, t1 in self.windows)] or rows
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in inside for i in range(4) if r[4][i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(r[1] for r in inside), "sm_max_mhz": max(r[2] for r in inside),
                "power_w_max": max(r[3] for r in inside), "samples": len(inside), "reasons": reasons}


# ------------------------------------------------------------------------------------------------------------------
# algorithmic bytes per launch of each operator tag (DESIGN.md "Roofline accounting")
# ------------------------------------------------------------------------------------------------------------------
def algorithmic_bytes(net, key: str, batch: int, sz: int) -> float | None:
    layer_s, tag, kernel = key.split(":", 2)
    l = int(layer_s)
    L = net.layers_count()
    P = net.param_count
    if tag == "sgd_update":
        return 3.0 * P * sz
    if tag == "cost_grad":
        return 3.0 * batch * net.layer_out_size(L - 1) * sz
    if tag == "reduce_update":
        return 3.0 * P * sz     # gradient in, parameter in, parameter out (partial sums are overhead of the split)
    if l < 0 or l >= L:
        return None
    in_n, out_n = net.layer_in_size(l), net.layer_out_size(l)
    n_par = (net.layer_param_offset(l + 1) if l + 1 < L else P) - net.layer_param_offset(l)
    spec = net._specs[l]
    bias_n = out_n if spec.type != 3 else 0
    weight_n = n_par - bias_n
    sel = 1 if kernel.startswith("conv28") else 4   # csrc/conv28.cu keeps one-byte window selectors instead of the int32 index map
    if tag == "conv_pool_fwd":   # fused CLayer + PLayer forward: input, pooled output, index map, parameters
        o1 = net.layer_out_size(l + 1)
        return batch * (in_n * sz + o1 * sz + o1 * sel) + n_par * sz
    if tag == "conv_pool_bwd":   # pooled gradient, index map, pooled values, input; bias + filter gradient out
        o1 = net.layer_out_size(l + 1)
        return batch * (o1 * sz + o1 * sel + o1 * sz + in_n * sz) + n_par * sz
    if tag == "head":            # x, reference, g, dx
        return batch * (in_n + 2 * out_n + in_n) * sz + n_par * sz
    if tag == "head_dW":
        return batch * (in_n + out_n) * sz + n_par * sz
    if kernel == "gemm_finalize_kernel":
        return None          # partial-sum traffic is overhead of the split, not algorithmic
    if tag in ("conv_fwd", "full_fwd"):
        return batch * (in_n + 2 * out_n) * sz + n_par * sz
    if tag == "softmax_fwd":
        return batch * 3 * out_n * sz
    if tag == "pool_fwd":
        return batch * (in_n * sz + out_n * sz + out_n * 4)
    if tag == "pool_bwd":
        return batch * (out_n * sz + out_n * 4 + in_n * sz)
    if tag == "act_grad_bias":
        return batch * 3 * out_n * sz + bias_n * sz
    if tag in ("full_dW", "conv_dK"):
        return batch * (out_n + in_n) * sz + weight_n * sz
    if tag in ("full_dX", "conv_dX"):
        return batch * (out_n + in_n) * sz + weight_n * sz
    return None


def step_bytes_per_sample(net, batch: int, sz: int) -> float:
    """SURVEY.md §8(d) accounting for the whole step: activations once per write and once per read across one fused
    kernel per layer per direction, index maps 4 B, parameters 6 P per step."""
    L = net.layers_count()
    act = 0
    for l in range(L):
        in_n, out_n = net.layer_in_size(l), net.layer_out_size(l)
        aux = out_n                                   # derivatives / index map written forward, read backward
        fwd = in_n + out_n + aux
        bwd = out_n + aux + (in_n if l > 0 else 0)    # delta in, aux in, input gradient out (not for layer 0)
        if net._specs[l].type != 3:
            bwd += in_n                               # the layer input is re-read for the weight gradient
        act += fwd + bwd
    act += 3 * net.layer_out_size(L - 1)              # cost: output, reference, derivative
    return act * sz + 6.0 * net.param_count * sz / batch


def measured_traffic(key: str, batch: int, real: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
    (profiles/traffic.json, written by tools/summarize_ncu.py --traffic), or None when that kernel was not captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            table = json.load(f)
        tag = key.split(":", 2)[1]
        return table.get(f"{real}_B{batch}", {}).get(tag)
    except Exception:
        return None


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def measured_peaks():
    """Arithmetic peaks measured on this pool's B200 by tools/peaks.cu (profiles/peaks_r2.json): TF32 tcgen05, FFMA2, DFMA."""
    try:
        with open(os.path.join(ROOT, "profiles", "peaks_r2.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def compute_fraction(net, key: str, batch: int, us: float, real: str, pk: dict) -> dict:
    """Arithmetic side of a contraction / convolution kernel: algorithmic flops per launch (2 per multiply-add, the reference's own
    operation count - NOT the 3x of the 3xTF32 scheme) against the measured peak of the pipe the kernel runs on (profiles/peaks_r2.json:
    3xTF32-equivalent tcgen05, packed FFMA2, DFMA).  Empty for the bandwidth kernels."""
    layer_s, tag, kernel = key.split(":", 2)
    l = int(layer_s)
    if l < 0 or l >= net.layers_count() or kernel == "gemm_finalize_kernel":
        return {}
    in_n, out_n = net.layer_in_size(l), net.layer_out_size(l)
    spec = net._specs[l]
    if tag in ("conv_pool_fwd", "conv_pool_bwd", "conv_fwd", "conv_dK", "conv_dX"):
        flops = 2.0 * batch * out_n * spec.in_c * spec.win_h * spec.win_w
    elif tag in ("full_fwd", "full_dW", "full_dX", "head_dW"):
        flops = 2.0 * batch * in_n * out_n
    elif tag == "head":
        flops = 4.0 * batch * in_n * out_n
    else:
        return {}
    if real == "f64":
        pipe, peak = "dfma", pk.get("dfma_tflops")
    elif "tc_gemm" in kernel:
        pipe, peak = "tcgen05 3xTF32 (fp32-equivalent)", pk.get("tf32_3x_equivalent_fp32_tflops")
    else:
        pipe, peak = "ffma2", pk.get("ffma2_tflops")
    tf = flops / (us * 1e-6) / 1e12
    return {"tflops": tf, "compute_pipe": pipe, "compute_peak_tflops": peak, "frac_of_compute_peak": (tf / peak) if peak else None}


def pool_roofline(dl, ctx, peak_gbs: float) -> dict:
    """The bandwidth kernels of the path measured as bandwidth kernels: PLayer max 2x2 forward / backward (layers.cuh, pool2x2_*) on
    CIFAR-deep's first pooling layer at mini-batch 1024 (32 x 32 x 32 float per sample: 201 MB of algorithmic traffic per launch - input,
    output and int32 index map), CUDA events around 24 launches over three buffer sets in rotation (604 MB: nothing is served from L2)."""
    import ctypes as C
    from deeplearning_b200 import _capi
    lib = ctx.lib
    B, c, h, w = 1024, 32, 32, 32
    spec = dl.PLayer((c, h, w), (2, 2), dl.PoolTypeId.MAX)
    n_in, n_out = B * c * h * w, B * c * (h // 2) * (w // 2)
    ptrs = []

    def alloc(nbytes):
        p = C.c_void_p()
        _capi.check(lib.dlb_malloc(ctx.h, nbytes, C.byref(p)))
        ptrs.append(p)
        return p
    try:
        x_host = np.random.default_rng(3).random(n_in, dtype=np.float32)
        sets = []
        for _ in range(3):
            x, y, m, dx = alloc(n_in * 4), alloc(n_out * 4), alloc(n_out * 4), alloc(n_in * 4)
            _capi.check(lib.dlb_memcpy_h2d(ctx.h, x, x_host.ctypes.data, x_host.nbytes))
            sets.append((x, y, m, dx))
        out = {"workload": "PLayer max 2x2 on {32,32,32} float, mini-batch 1024 (CIFAR-deep layer 1)", "algorithmic_bytes": n_in * 4 + n_out * 8,
               "peak": peak_gbs, "unit": "GB/s"}
        for name, call in (("forward", lambda b: lib.dlb_pool_forward_f32(ctx.h, C.byref(spec), B, b[0], b[1], b[2])),
                           ("backward", lambda b: lib.dlb_pool_backward_f32(ctx.h, C.byref(spec), B, b[1], b[2], b[3]))):
            for b in sets:
                _capi.check(call(b))
            ctx.synchronize()
            e0, e1 = dl.Event(ctx), dl.Event(ctx)
            e0.record()
            for r in range(24):
                _capi.check(call(sets[r % 3]))
            e1.record()
            ctx.synchronize()
            us = 1e3 * e0.elapsed_ms(e1) / 24
            gbs = out["algorithmic_bytes"] / us / 1e3
            out[name] = {"us_per_launch": us, "achieved": gbs, "frac": gbs / peak_gbs}
        return out
    finally:
        for p in ptrs:
            lib.dlb_free(ctx.h, p)


CIFAR_FLOP_PER_SAMPLE = 119_946_240
CIFAR_BYTES_PER_SAMPLE_F32 = 1_826_370


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# ------------------------------------------------------------------------------------------------------------------
def run_reference(real: str, batch: int, steps: int, warmup: int, parallelism: int = -1):
    """Net<CpuDC>::learn of the UNMODIFIED reference (oracle/_ref) with all host threads; falls back to the C port of it
    (one thread) when the prebuilt library is missing.  One step = one mini-batch."""
    from oracle import binding as ob
    specs = ob.lenet_specs()
    dtype = ob.dtype_of(real)
    cores = os.cpu_count() or 1
    n_w, n_t = max(warmup, 1) * batch, steps * batch
    x, y = synthetic_data(n_w + n_t, 784, 10, dtype)
    if ob.ref_available():
        net = ob.RefNet(specs, real, seed=0)
        net.learn(x[:n_w], y[:n_w], batch, 1, LR, CROSS_ENTROPY, LAMBDA, parallelism=parallelism, shuffle_seed=1)
        sec = net.learn(x[n_w:], y[n_w:], batch, 1, LR, CROSS_ENTROPY, LAMBDA, parallelism=parallelism, shuffle_seed=2)
        kind, used = "reference", (min(cores, batch) if parallelism < 0 else parallelism)
    else:
        from tests import helpers as H
        net = ob.OracleNet(specs, real)
        net.set_params(H.synthetic_params(net.param_count, specs, real))
        perm = np.arange(n_t, dtype=np.int64)[None, :]
        t0 = time.perf_counter()
        net.learn(x[n_w:], y[n_w:], batch, 1, LR, CROSS_ENTROPY, LAMBDA, 1, perm)
        sec = time.perf_counter() - t0
        kind, used = "port", 1
    return {"value": n_t / sec, "seconds": sec, "kind": kind, "cores": used,
            "sample": f"{steps} mini-batches of {batch} (= {n_t} items, 1 epoch) after {max(warmup, 1)} warm-up batches"}


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2048)
    ap.add_argument("--warmup", type=int, default=64)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--real", default="f32", choices=["f32", "f64"])
    ap.add_argument("--batch", type=int, default=0, help="override the (global) mini-batch size")
    ap.add_argument("--tc", type=int, default=1, choices=[0, 1, 2], help="fp32 contractions: 0 packed-FMA, 1 library default (tcgen05 from 128 samples per GPU), 2 tcgen05 always")
    ap.add_argument("--collective", default="p2p", choices=["p2p", "nccl"], help="multi-GPU gradient exchange")
    args = ap.parse_args()

    os.environ["NCCL_DEBUG"] = os.environ.get("DLB_NCCL_DEBUG", "WARN")   # keep NCCL's version banner off stdout: one JSON line only
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != max(args.gpus, 1) and world > 1:
        args.gpus = world
    n_gpus = max(args.gpus, 1)
    steps, warmup = max(args.steps, 1), max(args.warmup, 3)
    batch = args.batch or (256 if n_gpus == 1 else 8192)
    workload = ("LeNet-style CNN: CLayer 5x5x20 ReLU -> PLayer max 2x2 -> NLayer 100 ReLU -> NLayer 10 softmax, cross-entropy, "
                + (f"mini-batch {batch}, 1 B200 (BASELINE configs[1])" if n_gpus == 1 else
                   f"data-parallel, global mini-batch {batch} sharded over {n_gpus} B200, per-step gradient all-reduce (BASELINE configs[4])"))
    config = {"workload": workload, "global_batch": batch, "n_items": N_ITEMS, "lr": LR, "lambda": LAMBDA,
              "parallelism": f"dp{n_gpus}", "collective": ("none" if n_gpus == 1 else
                                                            "fused peer-memory reduce+update kernel over NVLink" if args.collective == "p2p" else "ncclAllReduce + update kernel"), "l2_policy": "inputs larger than L2: 65536-item resident data set (205 MB f32), "
              "each step gathers a fresh slice of the epoch permutation; no explicit flush"}
    sz = 4 if args.real == "f32" else 8

    # ---------------- reference arm -----------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        ref_steps = min(steps, 64 if batch <= 256 else 8)
        r = run_reference(args.real, batch, ref_steps, min(warmup, 4))
        line = {"impl": "reference", "metric": "train_samples_per_sec", "value": r["value"], "unit": "samples/s", "n_gpus": n_gpus,
                "steps": ref_steps, "warmup": min(warmup, 4), "ms_per_step": 1e3 * r["seconds"] / ref_steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": args.real, "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": r["kind"],
                                 "sample": r["sample"], "cpu": cpu_model()},
                "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ---------------- our arm -------------------------------------------------------------------------------------
    import deeplearning_b200 as dl

    dist = None
    if world > 1:
        import torch.distributed as dist   # plumbing only: rendezvous, barrier, max-over-ranks
        dist.init_process_group(backend="gloo")
    ctx = dl.Context(device=local_rank)
    net = dl.lenet(real=args.real, context=ctx)
    net.init_params(seed=0)
    if args.real == "f32" and args.tc != 1:
        net.enable_tensor_cores(args.tc)
    if world > 1 and args.collective == "nccl":
        uid = [dl.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        net.set_distributed(rank, world, uid[0])
    elif world > 1:   # default: fused reduce + update over NVLink peer memory (one kernel per rank, no NCCL on the step)
        handles = [None] * world
        dist.all_gather_object(handles, net.p2p_export())
        net.set_distributed_p2p(rank, world, handles)
    dtype = net.dtype
    x, y = synthetic_data(N_ITEMS, 784, 10, dtype)
    net.set_dataset(x, y)
    net.reset_random_generator(0)
    cost_id = dl.CostFunctionId.CROSS_ENTROPY

    def barrier():
        ctx.synchronize()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(v: float) -> float:
        if dist is None:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = dl.Event(ctx), dl.Event(ctx)

    # ---------------- data-parallel parity (N > 1), before anything is timed --------------------------------------------
    # The first PARITY_STEPS global mini-batches through the sharded path must leave bit-identical parameters on every
    # rank, equal (within the float bar of tests/helpers.py times 20, as tests/test_distributed_gpu.py) to the same steps
    # replayed by rank 0 alone on one GPU.  Only a few steps are replayed: float SGD trajectories with different summation
    # orders drift apart chaotically over thousands of steps.  After the timed regions the ranks are compared again, bit
    # for bit, over the whole history.
    dp_parity = None
    scaling_base = None
    PARITY_STEPS = 8

    def ranks_identical() -> bool:
        import hashlib
        digest = hashlib.sha256(net.get_params().tobytes()).hexdigest()
        all_d = [None] * world
        dist.all_gather_object(all_d, digest)
        return all(d == all_d[0] for d in all_d)

    if world > 1:
        from tests import helpers as H
        p_init = net.get_params()
        net.train_steps(batch, PARITY_STEPS, LR, cost_id, LAMBDA)
        same = ranks_identical()
        dp_params = net.get_params()
        if rank == 0:
            solo_ctx = dl.Context(device=local_rank)   # its own stream and scratch
            solo = dl.lenet(real=args.real, context=solo_ctx)
            solo.set_params(p_init)
            if args.real == "f32" and args.tc != 1:
                solo.enable_tensor_cores(args.tc)
            solo.set_dataset(x, y)
            solo.reset_random_generator(0)
            solo.train_steps(batch, PARITY_STEPS, LR, cost_id, LAMBDA)
            want = solo.get_params()
            H.assert_close("data-parallel parameters vs the single-GPU replay of the same global batches", dp_params, want, args.real, rtol_mult=20)
            scale = float(np.max(np.abs(want)))
            dp_parity = {"ranks_identical": bool(same), "replayed_steps": PARITY_STEPS,
                         "max_rel_err": float(np.max(np.abs(dp_params.astype(np.float64) - want) / (np.abs(want) + scale))),
                         "tolerance": H.RTOL[args.real] * 20, "update_norm": float(np.linalg.norm(want.astype(np.float64) - p_init))}
            # the same workload on ONE GPU (the base the scaling series is read against)
            solo.train_steps(batch, 3, LR, cost_id, LAMBDA)
            solo_ctx.synchronize()
            s0, s1 = dl.Event(solo_ctx), dl.Event(solo_ctx)
            s0.record(); it = solo.train_steps(batch, 16, LR, cost_id, LAMBDA); s1.record()
            scaling_base = {"value": it / (s0.elapsed_ms(s1) / 1e3), "unit": "samples/s", "what": f"global mini-batch {batch} on 1 GPU of this box, resident data"}
            solo.close()
        assert same, "data-parallel ranks ended the parity steps with different parameters"
        barrier()

    # resident-data throughput
    net.train_steps(batch, warmup, LR, cost_id, LAMBDA)
    barrier()
    launches0 = ctx.launch_count()
    w0 = time.time()
    e0.record()
    items = net.train_steps(batch, steps, LR, cost_id, LAMBDA)
    e1.record()
    ms = e0.elapsed_ms(e1)
    barrier()
    w1 = time.time()
    launches = ctx.launch_count() - launches0
    ms = max_over_ranks(ms)
    value = items / (ms / 1e3)
    if sampler:
        sampler.window(w0, w1)

    # end to end through host buffers (this rank's share of every batch)
    first, mine = dl.rank_slice(batch, rank, world)
    e2e_steps = min(steps, 512)
    n_host = 8
    hx = dl.pinned_array((n_host, mine, 784), dtype); hy = dl.pinned_array((n_host, mine, 10), dtype)
    for i in range(n_host):
        lo = (i * batch + first) % (N_ITEMS - batch)
        hx[i] = x[lo:lo + mine]; hy[i] = y[lo:lo + mine]
    reg_factor = LAMBDA / N_ITEMS
    costs = dl.pinned_array((e2e_steps,), np.float64)
    for i in range(4):
        net.step_host(hx[i % n_host], hy[i % n_host], LR, cost_id, reg_factor)
    barrier()
    w0 = time.time()
    e0.record()
    for i in range(e2e_steps):   # every step: H2D copy of its batch from pinned memory, the step, D2H copy of its cost
        net.step_host_async(hx[i % n_host], hy[i % n_host], LR, cost_id, reg_factor, costs[i:i + 1])
    ctx.synchronize()   # every step's cost has landed in host memory (the copies ride their own stream) before the clock stops
    e1.record()
    e2e_ms = max_over_ranks(e0.elapsed_ms(e1))
    barrier()
    cost = float(costs[-1])
    assert np.all(np.isfinite(costs)) and np.all(costs > 0), "end-to-end steps did not return their costs"
    w1 = time.time()
    if sampler:
        sampler.window(w0, w1)
    e2e = {"value": e2e_steps * batch / (e2e_ms / 1e3), "unit": "samples/s", "ms_per_step": e2e_ms / e2e_steps,
           "h2d_bytes_per_step": int(mine * (784 + 10) * sz), "d2h_bytes_per_step": 8, "steps": e2e_steps,
           "api": "dlb_net_step_host_async (pinned host batch in, batch cost out, copy of step i+1 overlaps step i)", "last_batch_cost": cost}

    line = {"metric": "train_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": n_gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.real,
            "data": "synthetic", "config": config, "e2e": e2e, "gpu_launches": int(launches)}

    if rank == 0 and n_gpus == 1:
        # per-kernel pass (instrumented with events; not part of `value`)
        dl.timing_enable(True)
        t_steps = min(steps, 64)
        net.train_steps(batch, t_steps, LR, cost_id, LAMBDA)
        rep = dl.timing_report()
        dl.timing_enable(False)
        total_ms = sum(v[1] for v in rep.values()) or 1.0
        kernels = []
        for k, (n, t) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
            ab = algorithmic_bytes(net, k, batch, sz)
            per_launch_us = 1e3 * t / n
            kernels.append({"kernel": k, "launches_per_step": n / t_steps, "us_per_launch": per_launch_us, "share": t / total_ms,
                            "algorithmic_bytes": ab, "gbs": (ab / (per_launch_us * 1e-6) / 1e9) if ab else None,
                            **compute_fraction(net, k, batch, per_launch_us, args.real, measured_peaks())})
        peak, peak_src = hbm_peak()
        top = kernels[0] if kernels else {"kernel": None, "gbs": None, "share": None, "us_per_launch": None}
        line["roofline"] = {"bound": "hbm", "kernel": top["kernel"], "achieved": top["gbs"], "peak": peak, "unit": "GB/s",
                            "frac": (top["gbs"] / peak) if top["gbs"] else None, "traffic": measured_traffic(top["kernel"], batch, args.real) if top["kernel"] else None,
                            "algorithmic_bytes": top.get("algorithmic_bytes"), "peak_source": peak_src,
                            "share_of_step": top["share"], "us_per_launch": top["us_per_launch"]}
        line["kernels"] = kernels[:12]
        bps = step_bytes_per_sample(net, batch, sz)
        line["step_roofline"] = {"bytes_per_sample": bps, "roofline_samples_per_s": peak * 1e9 / bps, "frac": value / (peak * 1e9 / bps),
                                 "note": "bytes = SURVEY 8(d)'s per-kernel (UNFUSED) accounting summed over the step: a fused step moves fewer bytes, "
                                         "so this fraction contains fusion credit and is not an achieved-bandwidth figure (see traffic / kernels)"}

        # the data-parallel workload's per-GPU share on one GPU, for reading the scaling series; with its own kernel table
        net.train_steps(8192, 3, LR, cost_id, LAMBDA)
        ctx.synchronize()
        e0.record(); it = net.train_steps(8192, 16, LR, cost_id, LAMBDA); e1.record()
        big = {"global_batch": 8192, "value": it / (e0.elapsed_ms(e1) / 1e3), "unit": "samples/s"}
        dl.timing_enable(True)
        net.train_steps(8192, 8, LR, cost_id, LAMBDA)
        rep = dl.timing_report()
        dl.timing_enable(False)
        tot = sum(v[1] for v in rep.values()) or 1.0
        big["kernels"] = []
        for k, (n, t) in sorted(rep.items(), key=lambda kv: -kv[1][1])[:10]:
            ab = algorithmic_bytes(net, k, 8192, sz)
            us = 1e3 * t / n
            big["kernels"].append({"kernel": k, "us_per_launch": us, "share": t / tot, "gbs": (ab / (us * 1e-6) / 1e9) if ab else None,
                                   "frac_of_hbm_peak": (ab / (us * 1e-6) / 1e9 / peak) if ab else None,
                                   **compute_fraction(net, k, 8192, us, args.real, measured_peaks())})
        bps8 = step_bytes_per_sample(net, 8192, sz)
        big["step_roofline_frac"] = big["value"] / (peak * 1e9 / bps8)
        big["step_roofline_note"] = "against the UNFUSED per-kernel byte accounting of SURVEY 8(d): fusion credit, not achieved DRAM bandwidth (per-kernel frac_of_hbm_peak above is)"
        line["config5_on_1_gpu"] = big

        pk = measured_peaks()
        # ---- BASELINE configs[2]: CIFAR-deep (3 conv + pool stages, ReLU, L2), mini-batch 1024, every contraction on tcgen05 (3xTF32)
        if args.real == "f32":
            try:
                net3 = dl.cifar_deep("f32", context=ctx)
                net3.init_params(seed=0)
                n3 = 4096
                x3, y3 = synthetic_data(n3, 3 * 32 * 32, 10, np.float32, seed=9)
                net3.set_dataset(x3, y3)
                net3.reset_random_generator(0)
                net3.train_steps(1024, 3, LR, cost_id, LAMBDA)
                ctx.synchronize()
                e0.record(); it3 = net3.train_steps(1024, 12, LR, cost_id, LAMBDA); e1.record()
                ms3 = e0.elapsed_ms(e1)
                v3 = it3 / (ms3 / 1e3)
                c3 = {"workload": "CIFAR-deep: C5x5x32 -> P2 -> C5x5x64 -> P2 -> C3x3x128 -> P2 -> N256 -> softmax 10, ReLU, L2, mini-batch 1024 (BASELINE configs[2])",
                      "value": v3, "unit": "samples/s", "ms_per_step": ms3 / 12, "plan": net3.describe_plan(),
                      "fp32_equivalent_tflops": v3 * CIFAR_FLOP_PER_SAMPLE / 1e12,
                      "hbm_roofline_samples_per_s": peak * 1e9 / CIFAR_BYTES_PER_SAMPLE_F32, "frac_of_hbm_roofline": v3 / (peak * 1e9 / CIFAR_BYTES_PER_SAMPLE_F32)}
                if pk.get("tf32_3x_equivalent_fp32_tflops"):
                    c3["tensor_roofline_tflops_3xtf32"] = pk["tf32_3x_equivalent_fp32_tflops"]
                    c3["frac_of_3xtf32_roofline"] = c3["fp32_equivalent_tflops"] / pk["tf32_3x_equivalent_fp32_tflops"]
                dl.timing_enable(True)
                net3.train_steps(1024, 2, LR, cost_id, LAMBDA)
                rep3 = dl.timing_report()
                dl.timing_enable(False)
                tot3 = sum(v[1] for v in rep3.values()) or 1.0
                c3["kernels"] = [{"kernel": k, "launches_per_step": n / 2, "us_per_launch": 1e3 * t / n, "share": t / tot3}
                                 for k, (n, t) in sorted(rep3.items(), key=lambda kv: -kv[1][1])[:8]]
                line["config3"] = c3
                net3.close()
            except Exception as exc:   # never lose the headline line to a secondary workload
                line["config3"] = {"error": str(exc)[:200]}
            # ---- Real = double (the reference's default Real, defs.h:25-29) on the same LeNet workload
            try:
                net64 = dl.lenet(real="f64", context=ctx)
                net64.init_params(seed=0)
                n64 = 16384
                net64.set_dataset(x[:n64].astype(np.float64), y[:n64].astype(np.float64))
                net64.reset_random_generator(0)
                net64.train_steps(256, 8, LR, cost_id, LAMBDA)
                ctx.synchronize()
                e0.record(); it64 = net64.train_steps(256, 56, LR, cost_id, LAMBDA); e1.record()
                ms64 = e0.elapsed_ms(e1)
                line["f64"] = {"workload": "LeNet, Real = double, mini-batch 256", "value": it64 / (ms64 / 1e3), "unit": "samples/s", "ms_per_step": ms64 / 56,
                               "contractions": "DFMA (tcgen05 has no fp64 kind; the DMMA path mma.sync.m8n8k4.f64 measures the same peak)",
                               "dfma_peak_tflops": pk.get("dfma_tflops"), "dmma_peak_tflops": pk.get("dmma_m8n8k4_tflops")}
                net64.close()
            except Exception as exc:
                line["f64"] = {"error": str(exc)[:200]}
            # ---- the same headline workload through the single-kernel step (dlb_net_enable_step_kernel; not the default)
            try:
                netm = dl.lenet(real="f32", context=ctx)
                netm.init_params(seed=0)
                netm.enable_step_kernel(True)
                netm.set_dataset(x, y)
                netm.reset_random_generator(0)
                netm.train_steps(256, 16, LR, cost_id, LAMBDA)
                ctx.synchronize()
                l0 = ctx.launch_count()
                e0.record(); itm = netm.train_steps(256, 200, LR, cost_id, LAMBDA); e1.record()
                msm = e0.elapsed_ms(e1)
                line["step_kernel"] = {"what": "LeNet B=256 with the whole step as one persistent cooperative kernel (csrc/mega.cu), all steps of the call in one launch",
                                       "value": itm / (msm / 1e3), "unit": "samples/s", "ms_per_step": msm / 200, "launches": int(ctx.launch_count() - l0),
                                       "plan": netm.describe_plan()}
                netm.close()
            except Exception as exc:
                line["step_kernel"] = {"error": str(exc)[:200]}

            # ---- BASELINE configs[3]: recurrent net (MNet of RMLayer), sequences of length 64, hidden size 256, mini-batch 512
            try:
                import time as _time
                depth, dims4 = 64, [64, 256, 16]
                rm = dl.MNet(depth, [dl.RMLayer(dims4[0], dims4[1]), dl.RMLayer(dims4[1], dims4[2])], real="f32", context=ctx)
                p4 = rm.init_params(seed=0, scale=0.05)
                rng4 = np.random.default_rng(5)
                x4 = rng4.standard_normal((512, depth, dims4[0])).astype(np.float32); y4 = (rng4.random((512, depth, dims4[2])) * 2 - 1).astype(np.float32)
                sq = dl.CostFunctionId.SQUARED_ERROR
                rm.set_batch(x4, y4); rm.learn_resident(sq, 0.01, 0.0, 2); ctx.synchronize()
                l0 = ctx.launch_count()
                e0.record(); rm.learn_resident(sq, 0.01, 0.0, 8); e1.record()
                ms4 = e0.elapsed_ms(e1) / 8
                launches4 = (ctx.launch_count() - l0) / 8
                t0 = _time.perf_counter()
                for _ in range(4): rm.learn(x4, y4, sq, 0.01, 0.0)
                ctx.synchronize()
                ms4_host = (_time.perf_counter() - t0) * 1e3 / 4
                flop4 = 0
                for l in range(2):   # multiply-adds x 2, forward + the three backward contractions per weight matrix
                    i_, h_ = dims4[l], dims4[l + 1]
                    flop4 += 2 * depth * (i_ * h_ + h_ * h_) * (3 if l > 0 else 2) + 2 * depth * h_ * h_
                c4 = {"workload": "MNet: RMLayer 64 -> 256 (tanh) -> RMLayer 256 -> 16 (tanh), sequences of length 64, squared error, gradient clipping 5, mini-batch 512 (BASELINE configs[3])",
                      "value": 512 / (ms4 / 1e3), "unit": "sequences/s", "ms_per_step": ms4, "launches_per_step": launches4,
                      "e2e": {"value": 512 / (ms4_host / 1e3), "unit": "sequences/s", "h2d_bytes_per_step": int(x4.nbytes + y4.nbytes), "d2h_bytes_per_step": 0,
                              "api": "dlb_mnet_learn_host (pageable host batch in, timed by the host clock around 4 calls + synchronize)"},
                      "fp32_tflops": 512 * flop4 / (ms4 / 1e3) / 1e12, "contractions": "gemm.cuh FFMA2 (the recurrence is 2 x 63 chained [512 x 256] x [256 x 256] contractions per layer)"}
                try:
                    from oracle import binding as _ob   # cpu_baseline leg: the reference's own MNet<CpuDC>::learn (sequential over the items)
                    if _ob.mnet_ref_available():
                        refm = _ob.RefMNet(depth, dims4, [_ob.TANH, _ob.TANH], [5.0, 5.0], "f32", seed=0)
                        refm.set_params(p4)
                        nref = 8
                        t0 = _time.perf_counter(); refm.learn(x4[:nref], y4[:nref], _ob.SQUARED_ERROR, 0.01, 0.0); dt = _time.perf_counter() - t0
                        c4["cpu_baseline"] = {"value": nref / dt, "unit": "sequences/s", "cores": 1, "kind": "reference",
                                              "sample": f"MNet<CpuDC>::learn on {nref} of the 512 sequences (the reference walks the items sequentially, MNet.inl:106-121)"}
                except Exception as exc:
                    c4["cpu_baseline"] = {"error": str(exc)[:200]}
                line["config4"] = c4
            except Exception as exc:
                line["config4"] = {"error": str(exc)[:200]}

        r = run_reference(args.real, 256, 64, 2)
        line["cpu_baseline"] = {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                                "cpu": cpu_model()}
        r1 = run_reference(args.real, 256, 8, 1, parallelism=1)
        line["cpu_baseline"]["one_thread"] = {"value": r1["value"], "unit": "samples/s", "cores": 1, "sample": r1["sample"]}
        if pk:
            line["measured_peaks"] = {k: pk[k] for k in ("tf32_tcgen05_cta1_m128_n256_tflops", "tf32_3x_equivalent_fp32_tflops", "ffma2_tflops", "dfma_tflops", "dmma_m8n8k4_tflops") if k in pk}
        try:
            line["pool_roofline"] = pool_roofline(dl, ctx, hbm_peak()[0])
        except Exception as exc:  # noqa: BLE001
            line["pool_roofline"] = {"error": str(exc)[:200]}
    if world > 1:
        same_after = ranks_identical()
        # per-kernel table of this rank at N > 1 (instrumented pass on every rank at once: the exchange kernel waits for its peers)
        dl.timing_enable(True)
        t_steps = 16
        net.train_steps(batch, t_steps, LR, cost_id, LAMBDA)
        rep = dl.timing_report()
        dl.timing_enable(False)
        barrier()
        if rank == 0:
            dp_parity["ranks_identical_after_timed_regions"] = bool(same_after)
            line["dp_parity"] = dp_parity
            line["scaling_base"] = scaling_base
            total_ms = sum(v[1] for v in rep.values()) or 1.0
            peak, _ = hbm_peak()
            kernels = []
            for k, (n, t) in sorted(rep.items(), key=lambda kv: -kv[1][1])[:10]:
                ab = algorithmic_bytes(net, k, mine, sz)
                us = 1e3 * t / n
                kernels.append({"kernel": k, "launches_per_step": n / t_steps, "us_per_launch": us, "share": t / total_ms,
                                "gbs": (ab / (us * 1e-6) / 1e9) if ab else None, "frac_of_hbm_peak": (ab / (us * 1e-6) / 1e9 / peak) if ab else None,
                                **compute_fraction(net, k, mine, us, args.real, measured_peaks())})
            line["kernels"] = kernels
            line["serial_kernel_us_per_step"] = 1e3 * total_ms / t_steps
            if kernels:
                line["limiter"] = {"kernel": kernels[0]["kernel"], "us_per_launch": kernels[0]["us_per_launch"], "share": kernels[0]["share"]}
        assert same_after, "data-parallel ranks hold different parameters after the timed regions"
    if sampler:
        line["clocks"] = sampler.stop()
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
