#!/usr/bin/env python
"""List the loops (backward branches) of one kernel in `cuobjdump -sass` output with their static instruction mix.
usage: sass_loops.py file.o kernel_substring"""
import collections
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
for f in funcs[1:]:
    name = f.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = []
    for line in f.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr_idx = {a: i for i, (a, _) in enumerate(ins)}
    print(name[:100], "instructions:", len(ins))
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_idx:
                body = ins[addr_idx[tgt]: i + 1]
                if len(body) < 40:
                    continue
                c = collections.Counter()
                for _, b in body:
                    op = re.sub(r"^@!?U?P\d+\s+", "", b).split()[0]
                    k = op.split(".")[0]
                    if k == "IMAD":
                        k = "IMAD." + (op.split(".")[1] if "." in op and op.split(".")[1] in ("MOV", "SHL", "WIDE", "IADD") else "mul")
                    c[k] += 1
                print(f"  loop {tgt:#x}..{a:#x}: {len(body)} instr:", ", ".join(f"{k} {v}" for k, v in c.most_common(12)))
