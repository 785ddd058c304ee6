#!/bin/bash
# Round 2, 8-GPU call: distributed tests, then the bench at N = 8 and 4 with the push (default) and the pull exchange, and the mnet test.
mkdir -p gpurun_out
echo "== pytest distributed + mnet"; timeout 900 python -m pytest tests/test_distributed_gpu.py tests/test_mnet_gpu.py -x -q 2>&1 | tail -4
run() {  # N tag extra-env
  env $3 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2952$1 bench.py --gpus $1 --steps 300 --warmup 20 > gpurun_out/r2t_scale_n$1_$2.json 2> gpurun_out/r2t_scale_n$1_$2.err
  echo "N=$1 $2:"; grep '^{' gpurun_out/r2t_scale_n$1_$2.json | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read())
print(round(d['value']), 'samples/s', round(1e3*d['ms_per_step'],1), 'us/step  e2e', round(d['e2e']['value']), 'parity', d['dp_parity']['ranks_identical'], d['dp_parity']['max_rel_err'], 'limiter', d.get('limiter'))"
  tail -1 gpurun_out/r2t_scale_n$1_$2.err | cut -c1-200
}
run 8 push DLB_P2P_PUSH=1
run 8 pull DLB_P2P_PUSH=0
run 4 push DLB_P2P_PUSH=1
run 2 push DLB_P2P_PUSH=1
echo "== config4"; python - <<'PY'
import numpy as np, time, deeplearning_b200 as dl
rm = dl.MNet(64, [dl.RMLayer(64, 256), dl.RMLayer(256, 16)], real="f32"); rm.init_params(seed=0, scale=0.05)
rng = np.random.default_rng(5)
x = rng.standard_normal((512, 64, 64)).astype(np.float32); y = (rng.random((512, 64, 16)) * 2 - 1).astype(np.float32)
rm.set_batch(x, y); rm.learn_resident(dl.CostFunctionId.SQUARED_ERROR, 0.01, 0.0, 2); rm.ctx.synchronize()
e0, e1 = dl.Event(rm.ctx), dl.Event(rm.ctx)
e0.record(); rm.learn_resident(dl.CostFunctionId.SQUARED_ERROR, 0.01, 0.0, 8); e1.record()
ms = e0.elapsed_ms(e1) / 8
print("MNet B=512: %.3f ms/step, %.0f sequences/s" % (ms, 512 / (ms / 1e3)))
PY
