#!/usr/bin/env python
"""Pretty-prints a bench.py JSON line."""
import json, sys
d = json.load(open(sys.argv[1]))
print("value %.0f %s  ms/step %.4f  e2e %.0f (%.4f ms/step)  launches/step %.1f  dtype %s  n_gpus %d" % (
    d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"] / d["steps"], d["dtype"], d["n_gpus"]))
if "roofline" in d:
    print(" roofline", {k: d["roofline"][k] for k in ("kernel", "achieved", "frac", "share_of_step", "us_per_launch")})
    print(" step_roofline", d.get("step_roofline"))
    for k in d.get("kernels", []):
        print("  %-52s n/step %.1f %8.2f us share %.3f GB/s %s" % (k["kernel"], k["launches_per_step"], k["us_per_launch"], k["share"],
                                                                   None if k["gbs"] is None else round(k["gbs"])))
    c5 = d.get("config5_on_1_gpu", {})
    for k in c5.get("kernels", []):
        print("  [B=8192] %-44s %9.2f us share %.3f GB/s %s frac %s | %s" % (k["kernel"], k["us_per_launch"], k["share"],
              None if k["gbs"] is None else round(k["gbs"]), None if k["frac_of_hbm_peak"] is None else round(k["frac_of_hbm_peak"], 3),
              "" if not k.get("frac_of_compute_peak") else "%.1f TFLOP/s = %.2f of %s" % (k["tflops"], k["frac_of_compute_peak"], k["compute_pipe"])))
    c5 = {k: v for k, v in c5.items() if k != "kernels"}
    print(" cfg5@1gpu", c5, " cpu", d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("cores"))
print(" clocks", d.get("clocks"))
