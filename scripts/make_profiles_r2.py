#!/usr/bin/env python
"""Turn the scratch outputs of scripts/profile_r2.sh (gpurun_out/) into the committed files under profiles/:
  r2_launches.csv        the ncu launch list as captured
  r2_kernels.csv         one row per --set full capture (scripts/summarize_ncu.py's columns), from prof_r2_raw.csv
  r2_index_kernels.csv   the same for the index builder's window-scan kernel
  r2_lines_<kernel>.txt  hot source lines (scripts/ncu_lines.py output, made on the GPU box)
  roofline_traffic.json  DRAM bytes per launch and K1's bytes per segment (bench.py reads it for roofline.traffic)
and print the per-step shares of the launch list (for profiles/README.md)."""
import csv
import json
import os
import re
import shutil
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
SEGMENTS = 400_000  # the reduced workload of profile_r2.sh: 200 k reads x 2 fragments
WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "gpu__time_duration.sum",
        "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def short(name):
    m = re.search(r"(k_[a-z0-9_]+|DeviceScan\w*|DeviceRadixSort\w*)", name)
    return m.group(1) if m else name[:40]


def summarise(raw, out):
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
    col = {h: i for i, h in enumerate(hdr)}
    res = {}
    for r in rows[2:]:
        k = short(r[col["Kernel Name"]])
        unit_t = units[col["gpu__time_duration.sum"]]
        t = float(r[col["gpu__time_duration.sum"]]) * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(unit_t, 1.0)
        def gb(c):
            v = float(r[col[c]]); u = units[col[c]]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
        res.setdefault(k, {"ms": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
        res[k]["ms"] += t; res[k]["dram_read_bytes"] += gb("dram__bytes_read.sum"); res[k]["dram_write_bytes"] += gb("dram__bytes_write.sum")
    return res


os.makedirs(P, exist_ok=True)
per = summarise(os.path.join(G, "prof_r2_raw.csv"), os.path.join(P, "r2_kernels.csv"))
if os.path.exists(os.path.join(G, "prof_r2_index_raw.csv")):
    summarise(os.path.join(G, "prof_r2_index_raw.csv"), os.path.join(P, "r2_index_kernels.csv"))
shutil.copy(os.path.join(G, "launches_r2.csv"), os.path.join(P, "r2_launches.csv"))
for f in os.listdir(G):
    if f.startswith("lines_r2_"):
        shutil.copy(os.path.join(G, f), os.path.join(P, "r2_" + f[len("lines_r2_"):].replace(".txt", "_lines.txt")))
ks = per.get("k_sketch", {})
json.dump({"source": "profiles/r2_kernels.csv (ncu --set full --clock-control none, scripts/profile_r2.sh: 200000 reads x 10 kb = 400000 segments vs 300 Mbp, s=220)",
           "segments_in_capture": SEGMENTS,
           "k_sketch_dram_bytes_per_segment": round((ks.get("dram_read_bytes", 0) + ks.get("dram_write_bytes", 0)) / SEGMENTS, 2),
           "per_kernel": per}, open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)

# launch list: the launches of the LAST resident (`value`) step: the index build and the warm-up step come first, the
# e2e parts after it
rows = [r for r in csv.reader(l for l in open(os.path.join(G, "launches_r2.csv")) if l.startswith('"'))]
hdr = rows[0]; c = {h: i for i, h in enumerate(hdr)}
ev = [(short(r[c["Kernel Name"]]), float(r[c["Metric Value"]]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6}.get(r[c["Metric Unit"]], 1e-6))
      for r in rows[1:]]
last_pack = max(i for i, (k, _) in enumerate(ev) if k == "k_pack_bases")
end = next(i for i in range(last_pack, len(ev)) if ev[i][0] == "k_l2_scan")  # the e2e steps' launches follow
step = defaultdict(float)
for k, t in ev[last_pack:end + 1]:
    step[k] += t
groups = {"K0": ["k_pack_bases"], "K1": ["k_sketch", "k_sketch_table"], "K2": ["k_l1_probe", "k_l1_warp", "k_l1_cta"],
          "K3": [k for k in step if k.startswith("k_l2") or k.startswith("Device")]}
tot = sum(sum(step[k] for k in v) for v in groups.values())
print("launch list, last step (ms):", {k: round(v, 3) for k, v in sorted(step.items(), key=lambda x: -x[1])})
for g, v in groups.items():
    s = sum(step[k] for k in v)
    print(f"  {g}: {s:.3f} ms = {100 * s / tot:.1f} %")
print("index build kernels (ms):", {k: round(t, 2) for k, t in ev[:last_pack] if t > 5})
