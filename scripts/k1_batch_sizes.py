"""K1 time per segment as a function of the number of segments per launch (C ABI events), text and packed input."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mashmap_b200 import capi

S, L, K = 220, 5000, 19
rng = np.random.default_rng(1)
nmax = 800_000
bases = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=nmax * L, dtype=np.uint8)]
for n_seg in (50_000, 134_218, 400_000, 800_000, 134_218):
    segs = np.zeros(n_seg, dtype=capi.segment_dtype)
    segs["offset"] = np.arange(n_seg, dtype=np.uint64) * L
    segs["length"] = L
    segs["seq_counter"] = np.arange(n_seg)
    segs["name_id"] = -1
    segs["ref_group"] = -1
    ctx = capi.Context(kmer_size=K, seg_length=L, sketch_size=S)
    t = []
    for it in range(4):
        ctx.sketch_segments(bases[: n_seg * L], segs)
        t.append(ctx.stage_ms()[0])
    print(f"{n_seg} segments per launch: K1 {min(t[1:]):.3f} ms = {min(t[1:]) * 1e6 / n_seg:.1f} ns per segment (runs {['%.3f' % x for x in t]})", flush=True)
    ctx.close()
