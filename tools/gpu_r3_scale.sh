#!/bin/bash
# final scaling series of round 2: usage gpu_r3_scale.sh "<list of N>" [tests]
mkdir -p gpurun_out
nvidia-smi -L | wc -l
if [ -n "$2" ]; then echo "== pytest distributed"; timeout 900 python -m pytest tests/test_distributed_gpu.py -x -q -m gpu 2>&1 | tail -4; fi
for N in $1; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 300 --warmup 20 > gpurun_out/r3_scale_n${N}_p2p.json 2> gpurun_out/r3_scale_n${N}_p2p.err
  echo "N=$N:"; grep '^{' gpurun_out/r3_scale_n${N}_p2p.json | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read())
print(round(d['value']), 'samples/s', round(1e3*d['ms_per_step'],1), 'us/step  e2e', round(d['e2e']['value']), 'parity', d['dp_parity']['ranks_identical'], d['dp_parity']['max_rel_err'], 'base', d.get('scaling_base'), 'limiter', d.get('limiter'))"
  tail -1 gpurun_out/r3_scale_n${N}_p2p.err | cut -c1-200
done
