"""The oracle (CPU restatement, oracle/libmm_oracle.so) pinned against the UNMODIFIED reference
(oracle/_ref harness) stage by stage, and against the committed golden fixtures generated from it.
No GPU."""
import json
import os

import numpy as np
import pytest

import datasets
import oracle_py
import refh
from mashmap_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
needs_ref = pytest.mark.skipif(not refh.available(), reason="oracle/_ref not built")
pytestmark = pytest.mark.skipif(not oracle_py.available(), reason="oracle/libmm_oracle.so not built")


def test_known_answer_hashes():
    """vectors produced by the reference's getHash"""
    L = oracle_py.lib()
    assert L.orc_hash(b"ACGTACGTACGTACGTACG", 19) == 0x272053CD152323BC
    assert L.orc_hash(b"AAAAAAAAAAAAAAAAAAA", 19) == 0xDBEF19B067885992
    assert L.orc_hash(b"GATTACAGATTACAGATTA", 19) == 0x12C6558960FB8EDC
    assert L.orc_hash(b"CGTACGTACGTACGTACGT", 19) == 0xE967113624A4C7C9  # reverse complement of the first


def test_known_answer_sketch():
    x, seq = 12345, []
    for _ in range(300):
        x = (x * 1664525 + 1013904223) & 0xFFFFFFFF
        seq.append("ACGT"[(x >> 24) & 3])
    sk = oracle_py.sketch_sequence("".join(seq), 19, 5, seq_id=7)
    exp = [(0x001ACB8DBE9AE686, 198, 198, -1), (0x009D802CB0003D60, 80, 80, 1), (0x00D46E04C7546695, 137, 137, 1),
           (0x0132A6788398CF41, 62, 62, 1), (0x022EE27ABF2D4617, 19, 19, 1)]
    assert [(int(m["hash"]), int(m["wpos"]), int(m["wpos_end"]), int(m["strand"])) for m in sk] == exp


def test_golden_fixtures():
    """fixtures written by tests/golden/make_golden.py from the reference harness"""
    path = os.path.join(HERE, "golden", "fragments.json")
    if not os.path.exists(path):
        pytest.skip("no golden fixtures")
    G = json.load(open(path))
    for case in G["sketch_cases"]:
        sk = oracle_py.sketch_sequence(case["seq"], case["k"], case["s"], seq_id=case["seq_id"])
        got = [[int(m["hash"]), int(m["wpos"]), int(m["wpos_end"]), int(m["strand"])] for m in sk]
        assert got == case["sketch"], case["name"]
    for case in G["min_hits"]:
        assert oracle_py.lib().orc_min_hits(case["s"], case["k"], case["pi"]) == case["value"]
    for case in G["cutoffs"]:
        O = oracle_py.Oracle(case["k"], 5000, case["s"], case["pi"])
        assert O.cutoffs().tolist() == case["value"]
        O.close()


def _attach_index(O, R):
    idx = R.index()
    keys, offs, pts, fr = R.lookup()
    O.set_index(idx, keys, offs, pts, fr, R.contig_len, R.contig_names)


@needs_ref
@pytest.mark.parametrize("which,args", [("random", ["-s", "5000", "--pi", "85"]),
                                        ("random", ["-s", "5000", "--pi", "95", "--dense"]),
                                        ("panel", ["-s", "5000", "--pi", "85"]),
                                        ("panel", ["-s", "2000", "--pi", "90", "-J", "25", "--noHgFilter"]),
                                        ("panel", ["-s", "5000", "--pi", "90", "-n", "3", "-f", "none"])])
def test_oracle_stages_match_reference(workdir, which, args):
    d = datasets.make_random_set(workdir, tag="orc") if which == "random" else datasets.make_panel_set(workdir, tag="orcp", n_strains=3,
                                                                                                       chrom_len=60_000)
    R = refh.RefSession(["-r", d["ref"], "-q", d["qry"], "-t", "2"] + args)
    try:
        O = oracle_py.Oracle(params=R.p)
        assert np.array_equal(O.cutoffs(), R.cutoffs())
        _attach_index(O, R)
        lens = [len(r) for r in d["reads"]]
        ridx, start, length = synth.split_segments(lens, R.p.segLength, R.p.kmerSize)
        n_l2 = 0
        for i in range(len(ridx)):
            frag = d["reads"][ridx[i]][start[i] : start[i] + length[i]]
            a = R.map_fragment(d["rnames"][ridx[i]], frag, full_len=lens[ridx[i]], seq_counter=int(ridx[i]))
            b = O.map_fragment(frag, seq_counter=int(ridx[i]), full_len=lens[ridx[i]])
            for f in ("hash", "wpos", "wpos_end", "strand"):
                assert np.array_equal(a["sketch"][f], b["sketch"][f]), (i, f)
            assert np.float32(a["kmerComplexity"]).tobytes() == np.float32(b["kmerComplexity"]).tobytes() or len(a["sketch"]) == 0
            assert a["n_points"] == b["n_points"]
            for f in ("pos", "hash", "seqId", "side"):
                assert np.array_equal(a["points"][f], b["points"][f]), (i, f)
            assert a["minimumHits"] == b["minimumHits"]
            assert a["l1"].tolist() == b["l1"].tolist(), i
            assert a["l2"].tolist() == b["l2"].tolist() and a["l2_cand"].tolist() == b["l2_cand"].tolist(), i
            n_l2 += len(a["l2"])
            for f in ("refStartPos", "refEndPos", "refSeqId", "conservedSketches", "strand", "blockLength", "approxMatches", "sketchSize"):
                assert np.array_equal(a["mappings"][f], b["mappings"][f]), (i, f)
            assert a["mappings"]["nucIdentity"].tobytes() == b["mappings"]["nucIdentity"].tobytes()
            assert a["mappings"]["nucIdentityUpperBound"].tobytes() == b["mappings"]["nucIdentityUpperBound"].tobytes()
        assert n_l2 > 0
        # whole reads (mapModule)
        for ri in range(len(d["reads"])):
            if lens[ri] < R.p.kmerSize:
                continue
            a = R.map_read(d["rnames"][ri], d["reads"][ri], ri)
            b = O.map_read(d["reads"][ri], ri)
            if len(a) == 0 and refh.is_uninitialised_n_merged_case(b, R.p.segLength, lens[ri]):
                continue  # reference UB (computeMap.hpp:1227/:1584/:429): outcome depends on stack garbage
            assert len(a) == len(b), ri
            for f in ("queryLen", "queryStartPos", "queryEndPos", "refSeqId", "refStartPos", "refEndPos", "strand",
                      "conservedSketches", "blockLength"):
                assert np.array_equal(a[f], b[f]), (ri, f)
            assert np.allclose(a["nucIdentity"], b["nucIdentity"], atol=1e-6, rtol=0)
        O.close()
    finally:
        R.close()
