"""CPU, world_size 2 over gloo: the data-parallel rule of the path - every rank takes its contiguous
dlb_rank_slice of each mini-batch, sums its per-sample gradients, the flat gradients are all-reduced (sum) and every rank
applies the identical update with 1/B_global - must reproduce the single-process reference semantics (oracle learn) and
leave bit-identical parameters on all ranks.  The per-sample arithmetic comes from the CPU oracle; the partitioning rule
is the product's own host-side helper (dlb_rank_slice), so this is the host logic of the N>1 path without a GPU."""
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, os.environ["DLB_ROOT"])
import torch, torch.distributed as dist
import deeplearning_b200 as dl
from oracle import binding as ob
from tests import golden_cases as gc, helpers as H

dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{os.environ['DLB_PORT']}", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
real, specs, cost = "f64", gc.MINI_LENET, ob.CROSS_ENTROPY
n, batch, epochs, lr, lam, seed = 44, 8, 2, 0.05, 0.3, 3
orc = ob.OracleNet(specs, real)
p0 = H.synthetic_params(orc.param_count, specs, real, seed=1)
x, y = H.synthetic_batch(orc, n, real, seed=2, lo=0.0, hi=1.0)
perms = dl.shuffle_sequence(seed, n, epochs)
# regularised mask in the flat layout: per layer [biases (not regularised), weights (regularised)]
mask = np.zeros(orc.param_count, bool); off = 0
for s, out_n in zip(specs, orc.out_sizes):
    if s.type == ob.PULL: continue
    nb = out_n; nw = (s.out_c * s.in_c * s.in_h * s.in_w) if s.type == ob.FULL else s.out_c * s.in_c * s.win_h * s.win_w
    mask[off + nb: off + nb + nw] = True; off += nb + nw
params = p0.copy()
reg = lam / n
for e in range(epochs):
    start = 0
    while start < n:
        end = n if (n - start) < 1.5 * batch else start + batch
        first, mine = dl.rank_slice(end - start, rank, world)
        orc.set_params(params)
        g = np.zeros(orc.param_count)
        for i in perms[e, start + first: start + first + mine]:
            g += orc.trace(x[i], y[i], cost)["grads"]
        t = torch.from_numpy(g); dist.all_reduce(t, op=dist.ReduceOp.SUM); g = t.numpy()
        l = -lr; gs = l * (1.0 / (end - start))
        params = np.where(mask, params * (1.0 + reg * l) + g * gs, params + g * gs)
        start = end
gathered = [torch.zeros(orc.param_count, dtype=torch.float64) for _ in range(world)]
dist.all_gather(gathered, torch.from_numpy(params))
if rank == 0:
    assert all(torch.equal(gathered[0], t) for t in gathered), "ranks ended with different parameters"
    orc.set_params(p0)
    orc.learn(x, y, batch, epochs, lr, cost, lam, threads=1, perms=perms)
    err = np.max(np.abs(orc.get_params() - params))
    assert err < 1e-12, err
    print("DP-OK", err)
dist.destroy_process_group()
"""


def test_two_rank_data_parallel_matches_single_process(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", DLB_PORT=str(port), DLB_ROOT=ROOT, OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "DP-OK" in outs[0], outs[0]
