#!/usr/bin/env python
"""Summarises ncu output brought back in gpurun_out/ into small text files under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/a_launches.csv profiles/r1_v1_launches.txt
  python tools/summarize_ncu.py full gpurun_out/a_prof_gemm.ncu-rep profiles/r1_v1_gemm_full.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum", "l1tex__t_bytes.sum", "lts__t_bytes.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct")


def launches(src, dst):
    lines = open(src).read().splitlines()
    hi = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    agg = collections.OrderedDict()
    order = []
    for row in csv.DictReader(lines[hi:]):
        if row["Metric Name"] != "gpu__time_duration.sum":
            continue
        full = row["Kernel Name"]
        m = re.match(r"void (\w+)<(.*)>\(", full)
        name = full
        if m:
            args = m.group(2)
            tags = re.findall(r"(Epi\w+|Im2colLoader<\w+, \d>|GradPixelLoader|GradGatherLoader|FilterTLoader|DenseColLoader|DenseRowLoader)", args)
            head = args.split(",")[0]
            name = f"{m.group(1)}<{head}" + (", " + ",".join(tags) if tags else "") + ">"
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(row["Metric Unit"], v)
        key = (name, row["Grid Size"], row["Block Size"])
        a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += v
        order.append(key)
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({src}); per-launch times are cold-cache and serialised:\n")
        f.write(f"# compare SHARES, not absolutes.  {len(order)} launches, {tot:.1f} us total\n")
        f.write("%-110s %-14s %-12s %6s %11s %10s %7s\n" % ("kernel", "grid", "block", "n", "total_us", "avg_us", "share"))
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-110s %-14s %-12s %6d %11.1f %10.2f %7.3f\n" % (k[0][:110], k[1], k[2], a[0], a[1], a[1] / a[0], a[1] / tot))
    print(open(dst).read())


TAGS = (("conv28_forward", "conv_pool_fwd"), ("conv28_backward", "conv_pool_bwd"), ("conv_pool_forward", "conv_pool_fwd"), ("conv_pool_backward", "conv_pool_bwd"), ("full_head_kernel", "head"),
        ("p2p_reduce_update_kernel", "reduce_update"), ("reduce_update_kernel", "reduce_update"), ("batch_outer_partial_kernel", "head_dW"),
        ("tc_gemm_kernel<EpiFullForward", "full_fwd"), ("tc_gemm_kernel<EpiWeightAndBias", "full_dW"),
        ("tc_gemm_kernel<EpiStoreTimesDeriv", "full_dX"))


def traffic(src, key, dst):
    """profiles/traffic.json[key][bench tag] = dram read + write bytes per launch (mean over the captured launches) from a
    raw-page CSV of an ncu --set full capture.  key is e.g. f32_B256."""
    import json, os
    rows = list(csv.reader(open(src).read().splitlines()))
    header, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(header)}
    def num(r, k):
        v = float(r[idx[k]].replace(",", "")); u = units[idx[k]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    acc = collections.defaultdict(list)
    for r in data:
        name = r[idx["Kernel Name"]]
        for pat, tag in TAGS:
            if pat in name:
                acc[tag].append(num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum"))
                break
    table = json.load(open(dst)) if os.path.exists(dst) else {}
    table[key] = {t: sum(v) / len(v) for t, v in sorted(acc.items())}
    table.setdefault("_note", "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none "
                              "--cache-control none (warm L2: small batches are served from the 126 MB L2, so traffic can be below the algorithmic bytes)")
    json.dump(table, open(dst, "w"), indent=1, sort_keys=True)
    print(json.dumps(table[key], indent=1))


def full(src, dst):
    if src.endswith(".csv"):
        out = open(src).read()
    else:
        out = subprocess.check_output(["ncu", "-i", src, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(out.splitlines()))
    header, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(header)}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ({src}); one block per captured launch\n")
        for r in data:
            f.write(f"\n== {r[idx['Kernel Name']][:200]}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}\n")
            for k in KEYS:
                if k in idx:
                    f.write(f"  {k:75s} {r[idx[k]]:>16s} {units[idx[k]]}\n")
            if "dram__bytes_read.sum" in idx and "dram__bytes_write.sum" in idx:
                def num(k):
                    v = float(r[idx[k]].replace(",", "")); u = units[idx[k]]
                    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                f.write(f"  {'traffic = dram read + write (bytes)':75s} {num('dram__bytes_read.sum') + num('dram__bytes_write.sum'):16.0f}\n")
    print(open(dst).read()[:6000])


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
