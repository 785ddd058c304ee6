#!/bin/bash
# Round 2: full gpu suite, smoke, bench line
mkdir -p gpurun_out
echo "== pytest"; timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench f32"; timeout 900 python bench.py > gpurun_out/r2o_bench_f32.json 2> gpurun_out/r2o_bench_f32.err; tail -3 gpurun_out/r2o_bench_f32.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2o_bench_f32.json"))
print("value %.0f ms/step %.4f e2e %.0f launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
print("roofline", d.get("roofline"))
for k in ("config3","config4","f64","step_kernel","cpu_baseline","measured_peaks","step_roofline"):
    v=d.get(k)
    if isinstance(v,dict): v={a:b for a,b in v.items() if a!="kernels"}
    print(k, v)
print("cfg5@1gpu", {a:b for a,b in d["config5_on_1_gpu"].items() if a!="kernels"})
for k in d["config5_on_1_gpu"].get("kernels", [])[:8]: print("   ", k)
for k in d.get("kernels", [])[:10]: print("  ", k)
PY
