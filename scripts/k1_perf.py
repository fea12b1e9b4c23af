"""K1 (and K0) alone on random segments: kernel time from the C ABI's own CUDA events. For quick A/B runs and ncu captures:
  python scripts/k1_perf.py [n_segments] [sketch] [seg_length] [k]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mashmap_b200 import capi

n_seg = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 220
L = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
K = int(sys.argv[4]) if len(sys.argv) > 4 else 19
rng = np.random.default_rng(1)
bases = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n_seg * L, dtype=np.uint8)]
segs = np.zeros(n_seg, dtype=capi.segment_dtype)
segs["offset"] = np.arange(n_seg, dtype=np.uint64) * L
segs["length"] = L
segs["seq_counter"] = np.arange(n_seg)
segs["name_id"] = -1
segs["ref_group"] = -1
ctx = capi.Context(kmer_size=K, seg_length=L, sketch_size=S)
for it in range(3):
    out, cnt = ctx.sketch_segments(bases, segs)
    ms = ctx.stage_ms()[0]
    print(f"iter {it}: K1 {ms:.3f} ms for {n_seg} segments of {L} bp (k={K}, s={S}): {n_seg * (L - K + 1) / ms / 1e6:.2f} G positions/s, "
          f"{n_seg * L / ms / 1e6:.2f} Gbp/s; full sketches {float((cnt == S).mean()):.4f}", flush=True)
ctx.close()
