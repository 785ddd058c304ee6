"""GPU, 2 ranks (needs >= 2 CUDA devices; skipped otherwise): LeNet data-parallel learn over NCCL must reproduce the
single-GPU learn on the same global batches (the gradient sum is split differently, so: tolerance) and leave identical
parameters on both ranks."""
import ctypes
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def device_count() -> int:
    try:
        cu = ctypes.CDLL("libcuda.so.1")
        if cu.cuInit(0) != 0:
            return 0
        n = ctypes.c_int(0)
        cu.cuDeviceGetCount(ctypes.byref(n))
        return n.value
    except OSError:
        return 0


WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, os.environ["DLB_ROOT"])
import torch.distributed as dist
import deeplearning_b200 as dl
from oracle import binding as ob
from tests import helpers as H

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{os.environ['DLB_PORT']}", rank=rank, world_size=world)
real = os.environ["DLB_REAL"]
specs = ob.lenet_specs()
n, batch, epochs, lr, lam, seed = 1200, 256, 2, 0.05, 0.3, 3
p0 = H.synthetic_params(301130, specs, real, seed=1)
rng = np.random.default_rng(2)
x = rng.random((n, 784)).astype(ob.dtype_of(real)); y = ob.one_hot_labels(real, n)

mode = os.environ["DLB_MODE"]

def run(distributed):
    net = dl.Net([H.to_dl_spec(s) for s in specs], real=real, context=dl.Context(rank if distributed else 0))
    net.set_params(p0)
    if distributed and mode == "nccl":
        uid = [dl.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        net.set_distributed(rank, world, uid[0])
    elif distributed:   # NVLink peer memory: gather every rank's IPC handle
        handles = [None] * world
        dist.all_gather_object(handles, net.p2p_export())
        net.set_distributed_p2p(rank, world, handles)
    net.reset_random_generator(seed)
    net.learn(x, y, batch, epochs, lr, dl.CostFunctionId.CROSS_ENTROPY, lam)
    return net.get_params()

dp = run(True)
import torch
gathered = [torch.zeros(dp.size, dtype=torch.from_numpy(dp).dtype) for _ in range(world)]
dist.all_gather(gathered, torch.from_numpy(dp))
if rank == 0:
    assert all(torch.equal(gathered[0], t) for t in gathered), "ranks ended with different parameters"
    single = run(False)
    H.assert_close("data-parallel vs single-GPU parameters", dp, single, real, rtol_mult=20)
    print("DP-GPU-OK")
dist.barrier()
dist.destroy_process_group()
"""


@pytest.mark.skipif(device_count() < 2, reason="needs at least 2 CUDA devices")
@pytest.mark.parametrize("mode", ["nccl", "p2p"])
@pytest.mark.parametrize("real", ["f32", "f64"])
def test_two_gpu_learn_matches_single_gpu(tmp_path, real, mode):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", DLB_PORT=str(port), DLB_ROOT=ROOT, DLB_REAL=real, DLB_MODE=mode)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "DP-GPU-OK" in outs[0], outs[0]
