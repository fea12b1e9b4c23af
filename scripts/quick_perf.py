"""Early stage timing on a scaled model of config 2 (index built by the reference harness)."""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import refh
from mashmap_b200 import capi, synth

n_contigs, clen, n_reads = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
sk = sys.argv[4] if len(sys.argv) > 4 else "220"
wd = tempfile.mkdtemp()
t = time.time()
genome = synth.random_genome(n_contigs, clen, seed=1)
names = [f"c{i}" for i in range(n_contigs)]
reads, truth = synth.simulate_reads(genome, n_reads, 10000, 0.02, 0.14, seed=2)
ref = os.path.join(wd, "ref.fa"); synth.write_fasta(ref, names, genome)
print("gen", time.time() - t, flush=True)
t = time.time()
R = refh.RefSession(["-r", ref, "-q", ref, "-s", "5000", "--pi", "85", "-J", sk, "-t", "64"])
print("ref index", time.time() - t, flush=True)
ctx = capi.Context(kmer_size=19, seg_length=5000, sketch_size=R.p.sketchSize)
t = time.time()
idx = R.index(); keys, offs, pts, fr = R.lookup()
ctx.index_upload(idx, keys, offs, pts, fr, R.contig_len)
ctx.tables_upload(R.cutoffs(), R.min_hits_table())
print("upload", time.time() - t, len(idx), len(keys), len(pts), flush=True)
lens = [len(r) for r in reads]
ridx, start, length = synth.split_segments(lens, 5000, 19)
offsr = np.zeros(len(lens) + 1, dtype=np.int64); offsr[1:] = np.cumsum(lens)
bases = np.concatenate(reads)
segs = np.zeros(len(ridx), dtype=capi.segment_dtype)
segs["offset"] = offsr[ridx] + start; segs["length"] = length; segs["seq_counter"] = ridx; segs["name_id"] = -1; segs["ref_group"] = -1
ctx.batch_upload(bases, segs)
for it in range(4):
    t = time.time(); nc, nl = ctx.map_resident(); dt = time.time() - t
    ms = ctx.stage_ms()
    print(f"iter {it}: wall {dt*1e3:.2f} ms  K1 {ms[0]:.3f}  K2 {ms[1]:.3f}  K3 {ms[2]:.3f} ms  cands {nc} loci {nl}  "
          f"K1 Gbp/s {bases.size/ms[0]/1e6:.2f} total-kernel Gbp/s {bases.size/(ms[0]+ms[1]+ms[2])/1e6:.2f}", flush=True)
seg_res, cands, loci = ctx.batch_fetch()
print("avg points", seg_res["n_points"].mean(), "avg cands", seg_res["n_candidates"].mean())
