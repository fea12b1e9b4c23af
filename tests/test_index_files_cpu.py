"""Index persistence: the product reads what the reference's --saveIndex wrote and the reference reads
what the product wrote (reference src/map/include/winSketch.hpp:266-374: PREFIX.index, PREFIX.map, TSV). CPU only."""
import os
import subprocess

import numpy as np
import pytest

import datasets
import refh
from mashmap_b200 import capi, hostlib

pytestmark = pytest.mark.skipif(not os.path.exists(refh.REF_BIN), reason="oracle/_ref not built")

ARGS = ["-s", "5000", "--pi", "85", "-t", "3"]


def run(cmd):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, (cmd, p.stderr[-1500:])


def read_index_file(path):
    raw = open(path, "rb").read()
    n = int(np.frombuffer(raw[:8], dtype=np.uint64)[0])
    assert len(raw) == 8 + 24 * n
    return np.frombuffer(raw[8:], dtype=capi.minmer_dtype, count=n)


def read_map_file(path):
    """{hash: [(pos, seqId, side), ...]}; the key order in the file is unspecified (std::unordered_map iteration)"""
    raw = open(path, "rb").read()
    n = int(np.frombuffer(raw[:8], dtype=np.uint64)[0])
    at, out = 8, {}
    for _ in range(n):
        key, cnt = (int(x) for x in np.frombuffer(raw[at : at + 16], dtype=np.uint64))
        at += 16
        pts = np.frombuffer(raw[at : at + 24 * cnt], dtype=capi.ipoint_dtype, count=cnt)
        at += 24 * cnt
        assert np.all(pts["hash"] == key)
        out[key] = list(zip(pts["pos"].tolist(), pts["seqId"].tolist(), pts["side"].tolist()))
    assert at == len(raw)
    return out


def same_index(a, b):
    for x, y in zip(a.arrays(), b.arrays()):
        if x.dtype.names:
            for f in x.dtype.names:
                if not f.startswith("_") and not np.array_equal(x[f], y[f]):
                    return False
        elif not np.array_equal(x, y):
            return False
    return a.freq_threshold == b.freq_threshold


@pytest.fixture(scope="module")
def d(workdir):
    return datasets.make_panel_set(workdir, tag="idxf", n_strains=3, chrom_len=60_000)


def test_binary_index_files_interoperate(workdir, d):
    base = ["-r", d["ref"], "-q", d["qry"]] + ARGS
    ref_prefix, our_prefix = os.path.join(workdir, "ref_saved"), os.path.join(workdir, "our_saved")
    paf1, paf2 = os.path.join(workdir, "idx1.paf"), os.path.join(workdir, "idx2.paf")
    run([refh.REF_BIN] + base + ["--saveIndex", ref_prefix, "-o", paf1])
    built = hostlib.HostIndex.from_cli(base)
    # (1) the product loads the reference's files and ends up with the same Sketch as when it builds it
    loaded = hostlib.HostIndex.from_cli(base + ["--loadIndex", ref_prefix])
    assert built.n_minmers > 0 and same_index(built, loaded)
    # (2) the product's files hold what the reference's hold (saved before the frequent-seed filter, winSketch.hpp:127-134)
    saver = hostlib.HostIndex.from_cli(base + ["--saveIndex", our_prefix])
    assert same_index(built, saver)
    a, b = read_index_file(ref_prefix + ".index"), read_index_file(our_prefix + ".index")
    assert len(a) == len(b)
    for f in ("hash", "wpos", "wpos_end", "seqId", "strand"):
        assert np.array_equal(a[f], b[f]), f
    assert read_map_file(ref_prefix + ".map") == read_map_file(our_prefix + ".map")
    # (3) the reference maps from the product's files exactly as from its own build
    run([refh.REF_BIN] + base + ["--loadIndex", our_prefix, "-o", paf2])
    assert open(paf1).read() == open(paf2).read() and os.path.getsize(paf1) > 0
    for h in (built, loaded, saver):
        h.close()


def test_tsv_index_files_interoperate(workdir, d):
    base = ["-r", d["ref"], "-q", d["qry"]] + ARGS
    ref_tsv, our_tsv = os.path.join(workdir, "ref_saved.tsv"), os.path.join(workdir, "our_saved.tsv")
    run([refh.REF_BIN] + base + ["--saveIndex", ref_tsv, "-o", os.path.join(workdir, "idx3.paf")])
    saver = hostlib.HostIndex.from_cli(base + ["--saveIndex", our_tsv])
    assert open(ref_tsv).read() == open(our_tsv).read()
    assert read_map_file(ref_tsv + ".map") == read_map_file(our_tsv + ".map")
    loaded = hostlib.HostIndex.from_cli(base + ["--loadIndex", ref_tsv])
    built = hostlib.HostIndex.from_cli(base)
    assert same_index(built, loaded)
    for h in (built, loaded, saver):
        h.close()
