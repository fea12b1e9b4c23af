"""Randomised differential tests on the CPU (seeded, a few seconds each): random parameter sets and deliberately awkward inputs
-- tandem repeats, N runs, IUPAC and lower-case letters, low-complexity and palindromic sequence, contigs that share a mutated
prefix, reads shorter than k, shorter than a segment and several segments long, noisy and reverse-complemented -- through
  (1) the window machine of the index builder (one restatement compiled for the host builder AND the device kernel,
      csrc/mm_winmachine.h), whole-contig and chunked + stitched, against the reference's addMinmers record for record;
  (2) the oracle port against the unmodified reference stage by stage (sketch, interval points, L1 candidates, L2 loci,
      fragment mappings, read mappings);
  (3) the product's host index (skch::Sketch: build, index, frequency filter) and host tail (mapModule on the device's record
      formats -> PAF fields) against the reference's.
The same generators were run over 15,000 window-machine cases, 180 stage configurations and 440 host configurations while this
file was written: no difference found."""
import os

import numpy as np
import pytest

import oracle_py
import refh
from mashmap_b200 import hostlib, synth

needs_ref = pytest.mark.skipif(not refh.available(), reason="oracle/_ref/libmm_ref.so not built")
FIELDS = ("hash", "wpos", "wpos_end", "seqId", "strand")


def awkward_sequence(rng):
    kind = int(rng.integers(0, 7))
    n = int(rng.integers(30, 20000))
    g = synth.random_sequence(n, rng)
    if kind == 1:  # tandem repeat, lightly mutated
        unit = synth.random_sequence(int(rng.integers(1, 300)), rng)
        g = np.tile(unit, n // len(unit) + 1)[:n].copy()
        m = rng.random(n) < rng.choice([0, 0.001, 0.01])
        g[m] = synth.random_sequence(int(m.sum()), rng)
    elif kind == 2:  # N runs, IUPAC letters, lower case
        for _ in range(int(rng.integers(1, 8))):
            a, ln = int(rng.integers(0, n)), int(rng.integers(1, 200))
            g[a : a + ln] = ord(rng.choice(list("NnRYK")))
        lo = rng.random(n) < 0.1
        g[lo] |= 0x20
    elif kind == 3:  # two- or three-letter alphabet
        g = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, int(rng.choice([2, 3])), n)].copy()
    elif kind == 4:  # palindrome
        h = g[: n // 2]
        g = np.concatenate([h, synth.revcomp(h)])
    elif kind == 5:  # homopolymer stretches
        for _ in range(int(rng.integers(1, 6))):
            a, ln = int(rng.integers(0, n)), int(rng.integers(10, 400))
            g[a : a + ln] = ord(rng.choice(list("ACGT")))
    elif kind == 6:  # a duplicated segment
        a, ln = int(rng.integers(0, max(1, n // 2))), int(rng.integers(20, max(21, n // 3)))
        b = int(rng.integers(0, max(1, n - ln)))
        g[b : b + ln] = g[a : a + ln][: len(g[b : b + ln])]
    return g


def awkward_genome(rng, n_contigs, length):
    cs = []
    for _ in range(n_contigs):
        g = synth.random_sequence(int(length * rng.uniform(0.3, 1.5)), rng)
        kind, n = int(rng.integers(0, 5)), len(g)
        if kind == 1:
            unit = synth.random_sequence(int(rng.integers(20, 2000)), rng)
            a = int(rng.integers(0, n // 2))
            rep = np.tile(unit, int(rng.integers(2, 30)))[: n - a]
            g[a : a + len(rep)] = rep
        elif kind == 2:
            for _ in range(3):
                a = int(rng.integers(0, n))
                g[a : a + int(rng.integers(1, 500))] = ord("N")
        elif kind == 3 and cs:  # shares a mutated prefix with an earlier contig
            src = cs[int(rng.integers(0, len(cs)))]
            ln = min(len(src), n) // 2
            g[:ln] = src[:ln]
            m = rng.random(ln) < 0.02
            g[:ln][m] = synth.random_sequence(int(m.sum()), rng)
        cs.append(g)
    return cs


def awkward_reads(rng, genome, n_reads, seg):
    reads = []
    for _ in range(n_reads):
        c = genome[int(rng.integers(0, len(genome)))]
        ln = int(rng.choice([rng.integers(5, 60), rng.integers(60, seg), seg, rng.integers(seg, 3 * seg)]))
        ln = min(ln, len(c))
        a = int(rng.integers(0, len(c) - ln + 1))
        r = c[a : a + ln].copy()
        m = rng.random(ln) < rng.choice([0, 0.01, 0.05, 0.12])
        r[m] = synth.random_sequence(int(m.sum()), rng)
        if rng.random() < 0.5:
            r = synth.revcomp(r)
        if rng.random() < 0.2 and ln > 10:
            b = int(rng.integers(0, ln))
            r[b : b + int(rng.integers(1, 40))] = ord("N")
        reads.append(r)
    return reads


def random_arguments(rng):
    seg = int(rng.choice([300, 500, 1000, 2000, 5000]))
    args = ["-s", str(seg), "--pi", str(int(rng.choice([80, 85, 90, 95, 99]))), "-k", str(int(rng.choice([12, 15, 16, 19, 21, 27])))]
    pct = 0.001
    if rng.random() < 0.5:
        args += ["-J", str(int(rng.integers(1, 80)))]
    if rng.random() < 0.3:
        args += ["--noHgFilter"]
    if rng.random() < 0.3 and "-J" not in args:
        args += ["--dense"]
    if rng.random() < 0.4:
        pct = float(rng.choice([0.5, 5, 20]))
        args += ["--kmerThreshold", str(pct)]
    if rng.random() < 0.3:
        args += ["-n", str(int(rng.integers(1, 5)))]
    if rng.random() < 0.2:
        args += ["-f", "none"]
    if rng.random() < 0.2:
        args += ["--noMerge"]
    if rng.random() < 0.2:
        args += ["--kmerComplexity", str(rng.choice([0.1, 0.5, 0.9]))]
    if rng.random() < 0.2:
        args += ["--filterLengthMismatches"]
    if rng.random() < 0.2:
        args += ["--hgFilterAniDiff", str(rng.choice([0.5, 2])), "--hgFilterConf", "99"]
    return seg, pct, args


def same_records(a, b):
    return len(a) == len(b) and all(np.array_equal(a[f], b[f]) for f in FIELDS)


@needs_ref
def test_window_machine_on_random_parameters_and_awkward_sequences():
    rescanned = 0
    for case in range(400):
        rng = np.random.default_rng(7_000_000 + case)
        seq = awkward_sequence(rng)
        k = int(rng.choice([8, 11, 15, 16, 19, 21, 27, 32]))
        w = int(rng.choice([50, 100, 333, 500, 1000, 2000, 5000]))
        s = int(rng.integers(1, max(2, min(w - k, 60))))
        ref = refh.add_minmers(seq, k, w, s, seq_id=1)
        assert same_records(ref, hostlib.add_minmers(seq, k, w, s, seq_id=1)), (case, k, w, s, len(seq))
        chunk, warm = int(rng.choice([w + 7, 2 * w, 3 * w + 1, 8 * w])), int(rng.choice([w, 2 * w, w + 13]))
        got, r = hostlib.add_minmers_chunked(seq, k, w, s, chunk, warm, seq_id=1)
        rescanned += r
        assert same_records(ref, got), (case, k, w, s, len(seq), chunk, warm)
    print("chunks re-scanned from exact state:", rescanned)


def _session(workdir, tag, case):
    rng = np.random.default_rng(9_000_000 + case)
    genome = awkward_genome(rng, int(rng.integers(1, 6)), int(rng.choice([5000, 30000, 80000])))
    ref = os.path.join(workdir, f"fz_{tag}_{case}.fa")
    synth.write_fasta(ref, [f"c{i}" for i in range(len(genome))], genome)
    seg, pct, args = random_arguments(rng)
    reads = awkward_reads(rng, genome, 25, seg)
    return rng, genome, reads, pct, args, refh.RefSession(["-r", ref, "-q", ref, "-t", "2"] + args)


@needs_ref
def test_oracle_port_equals_the_reference_on_random_configurations(workdir):
    from test_oracle import _attach_index

    n_frag = 0
    for case in range(14):
        rng, genome, reads, pct, args, R = _session(workdir, "st", case)
        try:
            if refh.lib().refh_index_size(R.h) == 0:
                continue
            O = oracle_py.Oracle(params=R.p)
            assert np.array_equal(O.cutoffs(), R.cutoffs()), args
            _attach_index(O, R)
            for ri, read in enumerate(reads):
                if len(read) < R.p.kmerSize:
                    continue
                _, start, length = synth.split_segments([len(read)], R.p.segLength, R.p.kmerSize)
                for i in range(len(start)):
                    frag = read[start[i] : start[i] + length[i]]
                    a = R.map_fragment(f"q{ri}", frag, full_len=len(read), seq_counter=ri)
                    b = O.map_fragment(frag, seq_counter=ri, full_len=len(read))
                    n_frag += 1
                    where = (case, args, ri, i)
                    for f in ("hash", "wpos", "wpos_end", "strand"):
                        assert np.array_equal(a["sketch"][f], b["sketch"][f]), where
                    assert a["n_points"] == b["n_points"], where
                    assert a["l1"].tolist() == b["l1"].tolist(), where
                    assert a["l2"].tolist() == b["l2"].tolist() and a["l2_cand"].tolist() == b["l2_cand"].tolist(), where
                    for f in ("refStartPos", "refEndPos", "refSeqId", "conservedSketches", "strand", "blockLength"):
                        assert np.array_equal(a["mappings"][f], b["mappings"][f]), where
                    assert a["mappings"]["nucIdentity"].tobytes() == b["mappings"]["nucIdentity"].tobytes(), where
                a, b = R.map_read(f"q{ri}", read, ri), O.map_read(read, ri)
                if len(a) == 0 and refh.is_uninitialised_n_merged_case(b, R.p.segLength, len(read)):
                    continue
                assert len(a) == len(b), (case, args, ri)
                for f in ("queryLen", "queryStartPos", "queryEndPos", "refSeqId", "refStartPos", "refEndPos", "strand", "conservedSketches", "blockLength"):
                    assert np.array_equal(a[f], b[f]), (case, args, ri, f)
            O.close()
        finally:
            R.close()
    assert n_frag > 200


@needs_ref
def test_host_index_and_host_tail_equal_the_reference_on_random_configurations(workdir):
    from test_host_cpu import _paf_fields, _records_from_reference_stages, _tail_params

    n_reads = 0
    for case in range(14):
        rng, genome, reads, pct, args, R = _session(workdir, "ht", case)
        try:
            seqs = np.concatenate(genome)
            offs = np.zeros(len(genome) + 1, dtype=np.uint64)
            offs[1:] = np.cumsum([len(g) for g in genome])
            hi = hostlib.HostIndex.build(seqs, offs, R.p.kmerSize, R.p.segLength, R.p.sketchSize, threads=int(rng.integers(1, 5)), kmer_pct_threshold=pct)
            if refh.lib().refh_index_size(R.h) == 0:
                assert hi.n_minmers == 0, args
                hi.close()
                continue
            mi, keys, offs2, pts, fr = hi.arrays()
            rkeys, roffs, rpts, rfr = R.lookup()
            assert hi.freq_threshold == R.freq_threshold(), args
            assert same_records(mi, R.index()), args
            assert np.array_equal(keys, rkeys) and np.array_equal(offs2, roffs) and np.array_equal(fr, rfr), args
            for f in ("pos", "hash", "seqId", "side"):
                assert np.array_equal(pts[f], rpts[f]), (args, f)
            hi.close()
            tail = hostlib.HostTail(_tail_params(R), R.contig_names, R.contig_len)
            d = {"reads": reads, "rnames": [f"q{i}" for i in range(len(reads))]}
            for ri in range(len(reads)):
                if len(reads[ri]) < R.p.kmerSize:
                    continue
                segs, seg_res, cands, loci = _records_from_reference_stages(R, d, ri, R.p.segLength, R.p.kmerSize)
                text, _ = tail.map_read(d["rnames"][ri], len(reads[ri]), ri, segs, seg_res, cands, loci)
                want = R.map_read(d["rnames"][ri], reads[ri], ri)
                got = [tuple(line.split("\t")) for line in text.splitlines()]
                n_reads += 1
                if len(want) == 0 and len(got) == 1 and len(reads[ri]) > R.p.segLength and int(got[0][3]) - int(got[0][2]) == R.p.segLength:
                    continue  # reference UB on n_merged (refh.is_uninitialised_n_merged_case)
                assert len(got) == len(want), (case, args, ri)
                for g, m in zip(got, want):
                    exp = _paf_fields(m, d["rnames"][ri], R.contig_names, R.contig_len)
                    assert tuple(str(x) for x in exp) == g[:11], (case, args, ri)
                    assert abs(float(g[12].split(":")[2]) - float(m["nucIdentity"])) <= 1e-4
            tail.close()
        finally:
            R.close()
    assert n_reads > 200


@needs_ref
@pytest.mark.parametrize("n,threads,span_every", [(3000, 1, 40), (50_000, 4, 0), (120_000, 8, 60)])
def test_run_wide_one_to_one_step_equals_the_reference_itself(workdir, n, threads, span_every):
    """-f one-to-one's run-wide step (computeMap.hpp:358-405) on the SAME random mappings through the reference's own
    filterByGroup / Filter::ref::filterMappings / std::sort calls (oracle/_ref, refh_one_to_one) and through the product's
    MapTail::finalizeOneToOne on several threads: same mappings in the same order. The mappings are full of ties -- equal
    sort keys, equal identities, equal start positions (the sweep refuses an equivalent mapping, and which one arrives first
    is decided by std::sort's treatment of equal keys) -- and some cover a whole contig or end on its last base (the
    position+1 wrap into the next contig of filter.hpp:311-324)."""
    import ctypes as C

    rng = np.random.default_rng(n + threads)
    lens = [int(x) for x in rng.integers(200_000, 400_000, 24)]
    unit = synth.random_sequence(1000, rng)  # only the contig LENGTHS matter to the step: one repeated unit keeps the session cheap
    ref = os.path.join(workdir, f"o2o_{n}.fa")
    synth.write_fasta(ref, [f"c{i}" for i in range(len(lens))], [np.resize(unit, ln) for ln in lens])
    R = refh.RefSession(["-r", ref, "-q", ref, "-s", "5000", "--pi", "90", "-f", "one-to-one", "-t", "2"])
    try:
        assert [int(x) for x in R.contig_len] == lens
        m = np.zeros(n, dtype=refh.mapping_dtype)
        n_q = max(10, n // 2)
        m["querySeqId"] = np.sort(rng.integers(0, n_q, n))  # read order, as mapQuery collects them
        m["queryLen"] = 20000
        m["queryStartPos"] = rng.integers(0, 4, n) * 5000
        m["queryEndPos"] = m["queryStartPos"] + 5000
        m["refSeqId"] = rng.integers(0, len(lens), n)
        clen = np.array(lens)[m["refSeqId"]]
        m["refStartPos"] = rng.integers(0, 800, n) * 250  # a coarse grid: many equal starts
        m["refEndPos"] = np.minimum(m["refStartPos"] + 4999 + rng.integers(0, 3, n) * 2500, clen - 1)
        if span_every:
            whole = rng.integers(0, span_every, n) == 0
            m["refStartPos"][whole] = 0
            m["refEndPos"][whole] = clen[whole] - 1
        tail_end = rng.integers(0, 50, n) == 0  # ends on the last base of its contig
        m["refEndPos"][tail_end] = clen[tail_end] - 1
        m["refStartPos"] = np.minimum(m["refStartPos"], m["refEndPos"])
        m["nucIdentity"] = rng.choice(np.float32([0.95, 0.96, 0.97, 0.9712, 0.99, 1.0]), n)
        m["nucIdentityUpperBound"] = m["nucIdentity"]
        m["blockLength"], m["sketchSize"], m["conservedSketches"], m["n_merged"] = 5000, 20, 15, 1
        m["strand"] = rng.choice([1, -1], n)
        m["kmerComplexity"] = 0.9
        want = np.zeros(n, dtype=refh.mapping_dtype)
        L = refh.lib()
        L.refh_one_to_one.restype = C.c_int64
        L.refh_one_to_one.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        n_want = L.refh_one_to_one(R.h, m.ctypes.data, n, want.ctypes.data)
        from test_host_cpu import _tail_params

        tail = hostlib.HostTail(_tail_params(R), R.contig_names, R.contig_len)
        H = hostlib.lib()
        H.skch_tail_one_to_one.restype = C.c_int64
        H.skch_tail_one_to_one.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int]
        got = np.zeros(n, dtype=refh.mapping_dtype)
        n_got = H.skch_tail_one_to_one(tail.h, m.ctypes.data, n, got.ctypes.data, n_q, threads)
        tail.close()
        assert 0 < n_want < n and n_got == n_want
        for f in ("querySeqId", "queryStartPos", "queryEndPos", "refSeqId", "refStartPos", "refEndPos", "strand", "nucIdentity"):
            assert np.array_equal(got[f][:n_got], want[f][:n_want]), f
        print(f"n={n}: {n_want} mappings kept by both")
    finally:
        R.close()


@needs_ref
@pytest.mark.parametrize("extra", [[], ["--legacy"], ["--reportPercentage"], ["--noMerge"], ["--noMerge", "--reportPercentage"]])
def test_paf_text_equals_the_reference_s_own_report(workdir, extra):
    """reportReadMappings (computeMap.hpp:1758-1805) of the reference ITSELF, through its std::ofstream, against the product's
    stream-free formatter on the same 20,000 random mappings: same bytes. Identities and complexities of every kind (floats,
    means of two to four floats computed in double, dyadic values, 0, 1, tiny values), in every output mode."""
    import ctypes as C

    from test_host_cpu import _tail_params

    rng = np.random.default_rng(len(extra) * 7 + 3)
    genome = [synth.random_sequence(int(n), rng) for n in (30_000, 45_000, 31_000)]
    ref = os.path.join(workdir, "paf_text.fa")
    synth.write_fasta(ref, ["chrA", "chrB some description", "c"], genome)
    R = refh.RefSession(["-r", ref, "-q", ref, "-s", "2000", "--pi", "90", "-t", "1"] + extra)
    try:
        n = 20_000
        m = np.zeros(n, dtype=refh.mapping_dtype)
        m["queryLen"] = rng.integers(1, 300_000, n)
        m["queryStartPos"] = rng.integers(0, 100_000, n)
        m["queryEndPos"] = m["queryStartPos"] + rng.integers(0, 100_000, n)
        m["refSeqId"] = rng.integers(0, 3, n)
        m["refStartPos"] = rng.integers(0, 30_000, n)
        m["refEndPos"] = m["refStartPos"] + rng.integers(0, 10_000, n)
        m["strand"] = rng.choice([1, -1], n)
        m["sketchSize"] = rng.integers(1, 1000, n)
        m["conservedSketches"] = (rng.random(n) * (m["sketchSize"] + 1)).astype(np.int32)
        m["blockLength"] = rng.integers(0, 100_000, n)
        f32 = rng.random(n).astype(np.float32)
        kind = rng.integers(0, 6, n)
        ident = np.where(kind == 0, (rng.integers(0, 129, n) / 128).astype(np.float32),
                 np.where(kind == 1, np.float32(1.0), np.where(kind == 2, f32 * np.float32(1e-4), np.where(kind == 3, np.float32(0.0), f32))))
        m["nucIdentity"] = ident
        a, b, c, d = (rng.random(n).astype(np.float32).astype(np.float64) for _ in range(4))
        kk = rng.integers(0, 5, n)
        m["kmerComplexity"] = np.where(kk == 0, a, np.where(kk == 1, (a + b) / 2, np.where(kk == 2, ((a + b) + c) / 3, np.where(kk == 3, (((a + b) + c) + d) / 4, 1.0))))
        path = os.path.join(workdir, "paf_text.out")
        L = refh.lib()
        L.refh_report_mappings.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_char_p, C.c_char_p]
        assert L.refh_report_mappings(R.h, m.ctypes.data, n, b"read_17 x", path.encode()) == 0
        want = open(path, "rb").read()
        tail = hostlib.HostTail(_tail_params(R), R.contig_names, R.contig_len)
        H = hostlib.lib()
        H.skch_tail_format.restype = C.c_void_p
        H.skch_tail_format.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_char_p, C.POINTER(C.c_uint64)]
        nb = C.c_uint64()
        p = H.skch_tail_format(tail.h, m.ctypes.data, n, b"read_17 x", C.byref(nb))
        got = C.string_at(p, nb.value)
        tail.close()
        if got != want:
            gl, wl = got.split(b"\n"), want.split(b"\n")
            bad = [(g, w) for g, w in zip(gl, wl) if g != w][:3]
            raise AssertionError(f"{len(gl)} vs {len(wl)} lines; first differences: {bad}")
        assert want.count(b"\n") == n
    finally:
        R.close()
