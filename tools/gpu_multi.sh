#!/bin/bash
# Multi-GPU call: the 2-rank NCCL test, then bench.py at N = 1 .. NGPU through torchrun.
mkdir -p gpurun_out
P=${1:-m}; NG=${2:-2}
nvidia-smi -L | head -8
echo "== pytest distributed gpu"; timeout 900 python -m pytest tests/test_distributed_gpu.py tests/test_cpp_host_gpu.py -x -q 2>&1 | tail -12
for N in 1 2 4 8; do
  if [ $N -gt $NG ]; then break; fi
  echo "== bench N=$N"
  if [ $N -eq 1 ]; then timeout 600 python bench.py --gpus 1 > gpurun_out/${P}_bench_n1.json 2> gpurun_out/${P}_bench_n1.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/${P}_bench_n$N.json 2> gpurun_out/${P}_bench_n$N.err; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 200 --warmup 10 --collective nccl > gpurun_out/${P}_bench_n${N}_nccl.json 2>/dev/null; echo "nccl:"; tail -c 1200 gpurun_out/${P}_bench_n${N}_nccl.json | cut -c1-200; fi
  tail -c 1200 gpurun_out/${P}_bench_n$N.json; echo; tail -3 gpurun_out/${P}_bench_n$N.err
done
