#!/usr/bin/env python
"""Turn the full-scale capture of scripts/final_r2.sh (gpurun_out/prof_r2_full_raw.csv: ncu --set full of every kernel of
a map step at BASELINE config 2 -- 2 M fragments against the 3 Gbp / 16 GB index) into profiles/r2_kernels_full.csv and
fold its DRAM traffic into profiles/roofline_traffic.json (bench.py's roofline.traffic then refers to the measured
workload itself; the reduced-workload figures stay under "reduced_workload")."""
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
SEGMENTS = 2_000_000

raw = os.path.join(G, "prof_r2_full_raw.csv")
rows = list(csv.reader(open(raw)))
hdr, units, body = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "gpu__time_duration.sum",
        "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
idx = [col[w] for w in WANT if w in col]


def short(name):
    import re
    m = re.search(r"(k_[a-z0-9_]+)", name)
    return m.group(1) if m else name[:40]


def val(r, c, scale):
    return float(r[col[c]].replace(",", "")) * scale.get(units[col[c]], 1.0)


T = {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3}
B = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
# the capture holds the warm-up step and the timed step: keep the LAST launch of every kernel
last = {}
for r in body:
    last[short(r[col["Kernel Name"]])] = r
with open(os.path.join(P, "r2_kernels_full.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
    for k, r in last.items():
        w.writerow([r[i] for i in idx])
def num(r, c, scale):
    try:
        v = val(r, c, scale)
        return v if v == v else 0.0  # a launch too short to sample reports nan
    except ValueError:
        return 0.0


per = {k: {"ms": num(r, "gpu__time_duration.sum", T), "dram_read_bytes": num(r, "dram__bytes_read.sum", B),
           "dram_write_bytes": num(r, "dram__bytes_write.sum", B)} for k, r in last.items()}
tp = os.path.join(P, "roofline_traffic.json")
old = json.load(open(tp)) if os.path.exists(tp) else {}
reduced = old.get("reduced_workload") or {k: old[k] for k in ("source", "segments_in_capture", "k_sketch_dram_bytes_per_segment", "per_kernel") if k in old}
ks = per.get("k_sketch", {})
json.dump({"source": "profiles/r2_kernels_full.csv (ncu --set full --clock-control none, scripts/final_r2.sh: BASELINE config 2 itself, "
                     "1 M reads x 10 kb = 2 M segments vs the 3 Gbp index, s=220; last launch of every kernel)",
           "segments_in_capture": SEGMENTS,
           "k_sketch_dram_bytes_per_segment": round((ks.get("dram_read_bytes", 0) + ks.get("dram_write_bytes", 0)) / SEGMENTS, 2),
           "per_kernel": per, "reduced_workload": reduced}, open(tp, "w"), indent=1)
for f in os.listdir(G):
    if f.startswith("lines_r2_full_"):
        shutil.copy(os.path.join(G, f), os.path.join(P, "r2_full_" + f[len("lines_r2_full_"):].replace(".txt", "_lines.txt")))
tot = sum(v["ms"] for v in per.values())
for k, v in sorted(per.items(), key=lambda x: -x[1]["ms"]):
    print(f"{k:16s} {v['ms']:8.2f} ms {100 * v['ms'] / tot:5.1f} %  dram R {v['dram_read_bytes'] / 1e9:7.2f} GB  W {v['dram_write_bytes'] / 1e9:6.2f} GB  "
          f"= {(v['dram_read_bytes'] + v['dram_write_bytes']) / SEGMENTS:9.1f} B/segment")
