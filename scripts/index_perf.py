"""Device index builder alone on a random reference: python scripts/index_perf.py [ref_bp] [contigs] [sketch] [seg]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from mashmap_b200 import capi, synth_gpu

ref_bp = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000_000
contigs = int(sys.argv[2]) if len(sys.argv) > 2 else 256
S = int(sys.argv[3]) if len(sys.argv) > 3 else 220
L = int(sys.argv[4]) if len(sys.argv) > 4 else 5000
dev = torch.device("cuda:0")
ref = synth_gpu.random_reference(contigs, ref_bp // contigs, seed=1, device=dev)
torch.cuda.synchronize()
ctx = capi.Context(kmer_size=19, seg_length=L, sketch_size=S)
offs = np.arange(contigs + 1, dtype=np.uint64) * np.uint64(ref_bp // contigs)
t0 = time.time()
st = ctx.index_build(None, offs, device_ptr=ref.data_ptr())
print(f"TPSM={os.environ.get('MM_INDEX_TPSM')} CHUNK={os.environ.get('MM_INDEX_CHUNK')}: {time.time() - t0:.2f} s; scan {st['ms_scan'] / 1e3:.2f} s, "
      f"records {st['ms_post'] / 1e3:.2f} s, lookup {st['ms_lookup'] / 1e3:.2f} s; {st['n_minmers']} minmers, {st['n_keys']} keys, {st['n_chunks']} chunks, "
      f"{st['n_fixed_chunks']} re-scanned", flush=True)
ctx.close()
