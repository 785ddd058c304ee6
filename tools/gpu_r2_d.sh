#!/bin/bash
# Round 2, GPU call D: the full gpu test suite, smoke, the bench line with config3 / f64 / step-kernel sub-lines.
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench f32"; timeout 900 python bench.py > gpurun_out/r2d_bench_f32.json 2> gpurun_out/r2d_bench_f32.err; tail -3 gpurun_out/r2d_bench_f32.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2d_bench_f32.json"))
print("value %.0f ms/step %.4f e2e %.0f launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
for k in ("config3","f64","step_kernel","cpu_baseline","measured_peaks"):
    v=d.get(k); 
    if isinstance(v,dict): v={a:b for a,b in v.items() if a!="kernels"}
    print(k, v)
print("cfg5@1gpu", d["config5_on_1_gpu"]["value"])
PY
