"""Generates tests/golden/fragments.json from the UNMODIFIED reference (oracle/_ref harness built from
/root/reference by oracle/Makefile). Run in the build container: python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refh  # noqa: E402
from mashmap_b200 import synth  # noqa: E402

rng = np.random.default_rng(2024)
cases = []
seqs = {
    "random_5000": synth.random_sequence(5000, rng),
    "random_300": synth.random_sequence(300, rng),
    "with_n": None, "lower_iupac": None, "tandem": np.tile(synth.random_sequence(37, rng), 60),
    "homopolymer": np.full(400, ord("A"), np.uint8), "short_19": synth.random_sequence(19, rng),
    "short_18": synth.random_sequence(18, rng),
}
w = synth.random_sequence(3000, rng)
w[100:130] = ord("N"); w[2990:] = ord("N")
seqs["with_n"] = w
seqs["lower_iupac"] = np.frombuffer((synth.random_sequence(1500, rng).tobytes().lower() + b"RYKMacgtnACGT" * 20), np.uint8).copy()
for name, q in seqs.items():
    for k, s in ((19, 20), (16, 7), (21, 64)):
        sk = refh.sketch_sequence(q, k, s, seq_id=3)
        cases.append(dict(name=f"{name}_k{k}_s{s}", seq=q.tobytes().decode("latin1"), k=k, s=s, seq_id=3,
                          sketch=[[int(m["hash"]), int(m["wpos"]), int(m["wpos_end"]), int(m["strand"])] for m in sk]))
L = refh.lib()
min_hits = [dict(s=s, k=k, pi=pi, value=L.refh_min_hits(s, k, pi)) for k in (19, 16) for pi in (0.85, 0.9, 0.95)
            for s in (1, 5, 20, 70, 130, 199, 220, 400)]
import tempfile

wd = tempfile.mkdtemp()
ref = os.path.join(wd, "r.fa")
synth.write_fasta(ref, ["c0"], synth.random_genome(1, 30_000, seed=3))
cutoffs = []
for s, pi in ((20, "95"), (70, "90"), (130, "85")):
    R = refh.RefSession(["-r", ref, "-q", ref, "-J", str(s), "--pi", pi])
    cutoffs.append(dict(s=s, k=19, pi=R.p.percentageIdentity, value=R.cutoffs().tolist()))
    R.close()
json.dump(dict(sketch_cases=cases, min_hits=min_hits, cutoffs=cutoffs), open(os.path.join(HERE, "fragments.json"), "w"))
print("wrote", len(cases), "sketch cases")
