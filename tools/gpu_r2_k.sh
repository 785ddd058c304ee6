#!/bin/bash
# Round 2, 2-GPU call: the distributed tests and the bench at N = 2 (peer-memory exchange, then NCCL).
mkdir -p gpurun_out
P=r2k
echo "== pytest distributed gpu"; timeout 900 python -m pytest tests/test_distributed_gpu.py -x -q 2>&1 | tail -6
for C in p2p nccl; do
  EXTRA=""; [ $C = nccl ] && EXTRA="--collective nccl"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 300 --warmup 20 $EXTRA > gpurun_out/${P}_scale_n2_$C.json 2> gpurun_out/${P}_scale_n2_$C.err
  echo "N=2 $C:"; grep '^{' gpurun_out/${P}_scale_n2_$C.json | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read())
print(round(d['value']), 'samples/s', round(1e3*d['ms_per_step'],1), 'us/step  e2e', round(d['e2e']['value']), 'dp_parity', d.get('dp_parity'), 'scaling_base', d.get('scaling_base'), 'limiter', d.get('limiter'))
for k in d.get('kernels', [])[:10]: print('   ', k)"
  tail -2 gpurun_out/${P}_scale_n2_$C.err | cut -c1-300
done
