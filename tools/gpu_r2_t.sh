#!/bin/bash
# 2-GPU: where the time inside the exchange kernel goes (DLB_P2P_PROF=1), push vs pull
mkdir -p gpurun_out
for M in 1 0; do
  DLB_P2P_PROF=1 DLB_P2P_PUSH=$M timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$M bench.py --gpus 2 --steps 300 --warmup 20 > gpurun_out/r2u_n2_push$M.json 2> gpurun_out/r2u_n2_push$M.err
  echo "push=$M:"; grep '^{' gpurun_out/r2u_n2_push$M.json | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), 'samples/s', round(1e3*d['ms_per_step'],1), 'us/step', d['dp_parity']['ranks_identical'])"
  grep "dlb p2p" gpurun_out/r2u_n2_push$M.err | head -8
done
