#!/usr/bin/env python
"""Development: Net::learn through the step kernel at mini-batch 37 - which combination of multi-step launches, tail batch and epochs breaks."""
import os, sys, subprocess
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def run():
    import deeplearning_b200 as dl
    from oracle import binding as ob
    from tests import helpers as H
    specs = ob.lenet_specs()
    for batch, n_items, epochs in [(37, 111, 1), (37, 111, 2), (37, 150, 1), (37, 150, 2), (300, 700, 2), (64, 200, 2)]:
        orc, gpu = H.make_pair(specs, "f32", seed=3, tc_mode=1)
        gpu.enable_step_kernel(os.environ.get('STEPK', '1') == '1')
        x, y = H.synthetic_batch(orc, n_items, "f32", seed=4, lo=0.0, hi=1.0)
        orc.learn(x, y, batch, epochs, 0.05, ob.CROSS_ENTROPY, 0.2, threads=1, perms=dl.shuffle_sequence(6, n_items, epochs))
        gpu.reset_random_generator(6)
        gpu.learn(x, y, batch, epochs, 0.05, dl.CostFunctionId.CROSS_ENTROPY, 0.2)
        got, want = gpu.get_params(), orc.get_params()
        out = []
        for l in range(4):
            a = gpu.layer_param_offset(l); b = gpu.layer_param_offset(l + 1) if l < 3 else gpu.param_count
            if b > a: out.append("L%d %.2e" % (l, np.abs(got[a:b].astype(np.float64) - want[a:b]).max() / (np.abs(want[a:b]).max() + 1e-30)))
        print("stepk=%s multi=%s batch %d items %d epochs %d: %s" % (os.environ.get("STEPK", "1"), os.environ.get("DLB_MEGA_MULTI", "1"), batch, n_items, epochs, " ".join(out)), flush=True)

if __name__ == "__main__":
    if os.environ.get("CHILD"):
        run()
    else:
        for stepk in ("1", "0"):
            r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=dict(os.environ, CHILD="1", STEPK=stepk), capture_output=True, text=True, timeout=900)
            print(r.stdout[-3000:]); print(r.stderr[-1500:])
