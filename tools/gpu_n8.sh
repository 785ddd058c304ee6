#!/bin/bash
mkdir -p gpurun_out
P=${1:-u}; N=${2:-8}
for C in p2p nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 300 --warmup 20 --collective $C > gpurun_out/${P}_bench_n${N}_$C.json 2> gpurun_out/${P}_bench_n${N}_$C.err
  echo "$C:"; grep '^{' gpurun_out/${P}_bench_n${N}_$C.json | tail -1 | cut -c1-220; tail -2 gpurun_out/${P}_bench_n${N}_$C.err | cut -c1-300
done
