#!/bin/bash
# Scaling series on one 8-GPU box: 2-rank distributed tests, then bench.py at N = 2, 4, 8 (fused peer-memory exchange; NCCL at 8).
mkdir -p gpurun_out
P=${1:-s}
echo "== pytest distributed gpu"; timeout 600 python -m pytest tests/test_distributed_gpu.py -x -q 2>&1 | tail -4
for N in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 300 --warmup 20 > gpurun_out/${P}_scale_n${N}_p2p.json 2> gpurun_out/${P}_scale_n${N}_p2p.err
  echo "N=$N p2p:"; grep '^{' gpurun_out/${P}_scale_n${N}_p2p.json | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), 'samples/s', round(1e3*d['ms_per_step'],1), 'us/step  e2e', round(d['e2e']['value']))"; tail -2 gpurun_out/${P}_scale_n${N}_p2p.err | cut -c1-200
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 300 --warmup 20 --collective nccl > gpurun_out/${P}_scale_n8_nccl.json 2> gpurun_out/${P}_scale_n8_nccl.err
echo "N=8 nccl:"; grep '^{' gpurun_out/${P}_scale_n8_nccl.json | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), 'samples/s', round(1e3*d['ms_per_step'],1), 'us/step  e2e', round(d['e2e']['value']))"
