#!/usr/bin/env python
"""End-to-end throughput of the mashmap-b200 program from FASTA files to PAF (reads parsed from disk cache):
writes a synthetic reference + reads as FASTA under a scratch directory, runs the CLI with the bulk (memory-mapped,
multi-threaded) reader and with the line reader (MM_SERIAL_INPUT=1), checks the two PAF files are identical.
usage: cli_throughput.py [n_reads] [ref_bp] [contigs]"""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from mashmap_b200 import hostlib, synth_gpu  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
ref_bp = int(sys.argv[2]) if len(sys.argv) > 2 else 300_000_000
contigs = int(sys.argv[3]) if len(sys.argv) > 3 else 32
READ_LEN = 10000
dev = torch.device("cuda:0")
wd = tempfile.mkdtemp(prefix="mmcli_")
ref = synth_gpu.random_reference(contigs, ref_bp // contigs, seed=1, device=dev)
reads, _ = synth_gpu.simulate_reads(ref, n_reads, READ_LEN, 0.02, 0.14, seed=2, chunk=8192)
ref_h, reads_h = ref.cpu().numpy(), reads.cpu().numpy().reshape(n_reads, READ_LEN)
t0 = time.time()
with open(os.path.join(wd, "ref.fa"), "wb") as f:
    for c in range(contigs):
        f.write(b">ctg%d\n" % c)
        f.write(ref_h[c].tobytes())
        f.write(b"\n")
nl = np.full((n_reads, 1), 10, dtype=np.uint8)
with open(os.path.join(wd, "reads.fa"), "wb") as f:  # header lines of equal width so that the file is one numpy block
    hdr = np.frombuffer(b"".join(b">read%09d\n" % i for i in range(n_reads)), dtype=np.uint8).reshape(n_reads, -1)
    f.write(np.concatenate([hdr, reads_h, nl], axis=1).tobytes())
print(f"wrote {os.path.getsize(os.path.join(wd, 'reads.fa')) / 1e9:.2f} GB of reads + reference in {time.time() - t0:.1f} s", flush=True)
outs = []
for tag, env in (("bulk reader + overlapped mapping", {}), ("line reader, synchronous", {"MM_SERIAL_INPUT": "1"})):
    out = os.path.join(wd, f"out_{len(outs)}.paf")
    e = dict(os.environ, **env)
    p = subprocess.run([hostlib.CLI_PATH, "-r", os.path.join(wd, "ref.fa"), "-q", os.path.join(wd, "reads.fa"), "-s", "5000", "--pi", "85",
                        "-t", str(os.cpu_count() or 8), "-o", out], env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [l for l in p.stderr.splitlines() if "time spent mapping the query" in l][-1]
    print(f"{tag}: {line.split('] ')[-1]}", flush=True)
    for l in p.stderr.splitlines():
        if "skch::Map]" in l or "input read and handed" in l or "reference index:" in l or "[trace] mapBatch" in l:
            print("   ", l, flush=True)
    outs.append(open(out).read())
print("PAF identical:", outs[0] == outs[1], "lines:", outs[0].count("\n"))
