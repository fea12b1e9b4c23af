"""Host tail micro-benchmark on the CPU: device-format records of simulated reads (made by the reference's own stage
functions through oracle/_ref), then the product's tail (mapRead + PAF formatting) timed in a C loop on one thread."""
import ctypes as C, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import refh, datasets
from mashmap_b200 import capi, hostlib, synth
from test_host_cpu import _records_from_reference_stages, _tail_params

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 400
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
wd = tempfile.mkdtemp()
genome = synth.random_genome(4, 500_000, seed=1)
names = [f"c{i}" for i in range(4)]
reads, truth = synth.simulate_reads(genome, n_reads, 10000, 0.02, 0.14, seed=2)
ref = os.path.join(wd, "ref.fa"); synth.write_fasta(ref, names, genome)
R = refh.RefSession(["-r", ref, "-q", ref, "-s", "5000", "--pi", "85", "-J", "220", "-t", "4"])
d = {"reads": reads, "rnames": [f"read{i}" for i in range(n_reads)]}
S, SR, CA, LO, first, lens = [], [], [], [], [0], []
nc = nl = 0
for ri in range(n_reads):
    segs, seg_res, cands, loci = _records_from_reference_stages(R, d, ri, 5000, 19)
    seg_res = seg_res.copy(); cands = cands.copy()
    seg_res["first_candidate"] += nc
    cands["first_locus"] += nl
    cands["segment"] += first[-1]
    S.append(segs); SR.append(seg_res); CA.append(cands); LO.append(loci)
    nc += len(cands); nl += len(loci)
    first.append(first[-1] + len(segs)); lens.append(len(reads[ri]))
S, SR, CA, LO = (np.concatenate(x) for x in (S, SR, CA, LO))
first = np.array(first, dtype=np.uint64); lens = np.array(lens, dtype=np.int32)
tail = hostlib.HostTail(_tail_params(R), R.contig_names, R.contig_len)
L = hostlib.lib()
L.skch_tail_bench.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 6 + [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
L.skch_tail_bench.restype = None
tm, tf, nm = C.c_double(), C.c_double(), C.c_uint64()
L.skch_tail_bench(tail.h, n_reads, lens.ctypes.data, first.ctypes.data, S.ctypes.data, SR.ctypes.data, CA.ctypes.data, LO.ctypes.data, iters,
                  C.byref(tm), C.byref(tf), C.byref(nm))
print(f"{n_reads} reads x {iters}: mapRead {tm.value / n_reads / iters * 1e6:.3f} us/read, format {tf.value / n_reads / iters * 1e6:.3f} us/read, "
      f"{nm.value / iters:.0f} mappings per pass; {len(CA)} candidates, {len(LO)} loci")
