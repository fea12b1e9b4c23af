#!/usr/bin/env python
"""Hot source lines of one kernel from an ncu report captured with --import-source on (read here, no GPU needed):
share of warp-stall samples and of executed instructions per CUDA source line, plus the summed stall reasons.
usage: ncu_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv
import subprocess
import sys
from collections import defaultdict

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur, hdr, out = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif len(r) > 7 and r[0].isdigit():
        try:
            out.append((cur, int(r[0]), r[1].strip()[:110], int(r[4]), int(r[7]), r))
        except ValueError:
            pass
if not out:
    sys.exit("no source-correlated rows (was the report captured with --import-source on and -lineinfo?)")
ts, ti = sum(o[3] for o in out), sum(o[4] for o in out)
print(f"{kern}: {ts} samples, {ti} warp instructions")
by_file = defaultdict(lambda: [0, 0])
for f, _, _, s, i, _ in out:
    by_file[f][0] += s
    by_file[f][1] += i
for f, (s, i) in sorted(by_file.items(), key=lambda x: -x[1][0]):
    print(f"  {f:34s} samples {100 * s / ts:5.1f}%  instructions {100 * i / ti:5.1f}%")
print("hot lines:")
for f, l, src, s, i, _ in sorted(out, key=lambda x: -x[3])[:top]:
    print(f"  {f}:{l:<5d} {100 * s / ts:5.1f}% {100 * i / ti:5.1f}%  {src}")
st = [k for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = defaultdict(int)
for o in out:
    for k in st:
        try:
            agg[hdr[k]] += int(o[5][k])
        except (ValueError, IndexError):
            pass
print("stall reasons:", ", ".join(f"{k[6:]} {100 * v / max(ts, 1):.0f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:6]))
