#!/usr/bin/env python
"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built from /root/reference).

Run in the build container only:   python tools/make_golden.py
The reference cannot travel to the GPU box, its outputs can: each fixture stores what the reference's Net<CpuDC>
produced (per-sample traces, summed gradients, parameters after Net::learn, predicted classes) for inputs and initial
parameters that the tests regenerate from numpy seeds (tests/golden_cases.py).  Both the C oracle (CPU tests) and
the CUDA path (GPU tests) are checked against these files.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import binding as ob          # noqa: E402
from tests import golden_cases as gc      # noqa: E402
from tests import helpers as H            # noqa: E402


def main():
    assert ob.build_ref(), "oracle/_ref is not available (needs /root/reference)"
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for case in gc.CASES:
        for real in case["reals"]:
            specs = case["specs"]
            ref = ob.RefNet(specs, real, seed=0)
            params = H.synthetic_params(ref.param_count, specs, real, seed=case["param_seed"])
            ref.set_params(params)
            x, y = gc.case_data(case, ref.in_sizes[0], ref.out_sizes[-1], real)
            store = {}
            nt = case["n_trace"]
            want = H.oracle_batch_trace(ref, x[:nt], y[:nt], case["cost"])
            if case["store_traces"]:
                for k in ("acts", "derivs", "index_map", "in_grads", "cost_grad", "grads"):
                    store["trace_" + k] = want[k]
            else:   # large net: keep a strided sample of the gradient plus exact integer data
                store["trace_grads_strided"] = want["grads"][::gc.STRIDE]
                store["trace_index_map"] = want["index_map"]
                store["trace_out"] = want["acts"][:, -ref.out_sizes[-1]:]
            ln = case["learn"]
            ref.learn(x, y, ln["batch"], ln["epochs"], ln["lr"], case["cost"], ln["lam"], parallelism=1, shuffle_seed=ln["seed"])
            after = ref.get_params()
            store["learn_perms"] = ob.ref_shuffle_sequence(ln["seed"], x.shape[0], ln["epochs"])
            store["learn_params"] = after if case["store_traces"] else after[::gc.STRIDE]
            store["learn_param_sum"] = np.array([after.astype(np.float64).sum(), (after.astype(np.float64) ** 2).sum()])
            cls, outs = ref.eval(x, want_outputs=True)
            store["learn_classes"] = cls
            store["learn_outputs"] = outs[:16]
            path = os.path.join(out_dir, f"{case['name']}_{real}.npz")
            np.savez_compressed(path, **store)
            print(f"{path}: {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
