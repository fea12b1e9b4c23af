"""Test-side wrapper of the reference harness oracle/_ref/libmm_ref.so (the UNMODIFIED reference,
compiled by oracle/Makefile from /root/reference where it lies). Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libmm_ref.so")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "mashmap_ref")

minmer_dtype = np.dtype(
    [("hash", "<u8"), ("wpos", "<i4"), ("wpos_end", "<i4"), ("seqId", "<i4"), ("strand", "<i2"), ("_pad", "<i2")]
)
ipoint_dtype = np.dtype(
    [("pos", "<i4"), ("_pad0", "<i4"), ("hash", "<u8"), ("seqId", "<i4"), ("side", "i1"), ("_pad1", "i1", (3,))]
)
l1_dtype = np.dtype([("seqId", "<i4"), ("rangeStartPos", "<i4"), ("rangeEndPos", "<i4"), ("intersectionSize", "<i4")])
l2_dtype = np.dtype(
    [("seqId", "<i4"), ("meanOptimalPos", "<i4"), ("optimalStart", "<i4"), ("optimalEnd", "<i4"),
     ("sharedSketchSize", "<i4"), ("strand", "<i4")]
)
mapping_dtype = np.dtype(
    [("queryLen", "<i4"), ("refStartPos", "<i4"), ("refEndPos", "<i4"), ("queryStartPos", "<i4"),
     ("queryEndPos", "<i4"), ("refSeqId", "<i4"), ("querySeqId", "<i4"), ("blockLength", "<i4"),
     ("nucIdentity", "<f4"), ("nucIdentityUpperBound", "<f4"), ("sketchSize", "<i4"), ("conservedSketches", "<i4"),
     ("strand", "<i4"), ("approxMatches", "<i4"), ("n_merged", "<i4"), ("splitMappingId", "<i4"),
     ("discard", "<i4"), ("selfMapFilter", "<i4"), ("kmerComplexity", "<f8")]
)


class OrcParams(C.Structure):
    _fields_ = [
        ("kmerSize", C.c_int32), ("segLength", C.c_int32), ("sketchSize", C.c_int32), ("alphabetSize", C.c_int32),
        ("percentageIdentity", C.c_float), ("filterMode", C.c_int32), ("numMappingsForSegment", C.c_int32),
        ("numMappingsForShortSequence", C.c_int32), ("block_length", C.c_int32), ("chain_gap", C.c_int32),
        ("split", C.c_int32), ("mergeMappings", C.c_int32), ("stage1_topANI_filter", C.c_int32),
        ("ANIDiff", C.c_float), ("ANIDiffConf", C.c_float), ("stage2_full_scan", C.c_int32),
        ("keep_low_pct_id", C.c_int32), ("kmer_pct_threshold", C.c_float), ("kmerComplexityThreshold", C.c_float),
        ("skip_self", C.c_int32), ("skip_prefix", C.c_int32), ("prefix_delim", C.c_int32),
        ("lower_triangular", C.c_int32), ("filterLengthMismatches", C.c_int32), ("legacy_output", C.c_int32),
        ("report_ANI_percentage", C.c_int32), ("sparsity_hash_threshold", C.c_uint64), ("referenceSize", C.c_uint64),
    ]


def is_uninitialised_n_merged_case(mappings, seg_length, read_len):
    """A split read whose fragments produced exactly ONE mapping: the reference then reads
    MappingResult::n_merged uninitialised (computeMap.hpp:1227 declares it, mergeMappingsInRange returns early
    at :1584, filterWeakMappings reads it at :429-430). Whether the mapping survives depends on stack garbage
    (observed: the CLI build drops it unless it came from fragment 0; the harness build varies). This repo and
    the oracle define n_merged = 1 (kept)."""
    if read_len <= seg_length or len(mappings) != 1:
        return False
    m = mappings[0]
    return int(m["queryEndPos"]) - int(m["queryStartPos"]) == seg_length and int(m["n_merged"]) == 1


def available():
    return os.path.exists(REF_LIB)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(REF_LIB)
        vp = C.c_void_p
        L.refh_open.restype = vp
        L.refh_open.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
        L.refh_close.argtypes = [vp]
        L.refh_params.argtypes = [vp, C.POINTER(OrcParams)]
        L.refh_n_contigs.argtypes = [vp]
        L.refh_contig_name.argtypes = [vp, C.c_int]
        L.refh_contig_name.restype = C.c_char_p
        L.refh_contig_len.argtypes = [vp, C.c_int]
        L.refh_index_size.argtypes = [vp]
        L.refh_index_size.restype = C.c_int64
        L.refh_index_data.argtypes = [vp]
        L.refh_index_data.restype = vp
        L.refh_lookup_build.argtypes = [vp]
        L.refh_lookup_build.restype = C.c_int64
        for f in ("refh_lookup_keys", "refh_lookup_offs", "refh_lookup_pts", "refh_lookup_isfreq"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = vp
        L.refh_freq_threshold.argtypes = [vp]
        L.refh_is_freq.argtypes = [vp, C.c_uint64]
        L.refh_cutoffs.argtypes = [vp, vp, C.c_int]
        L.refh_hash.argtypes = [C.c_char_p, C.c_int]
        L.refh_hash.restype = C.c_uint64
        L.refh_min_hits.argtypes = [C.c_int, C.c_int, C.c_float]
        L.refh_j2md.argtypes = [C.c_float, C.c_int]
        L.refh_j2md.restype = C.c_float
        L.refh_md2j.argtypes = [C.c_float, C.c_int]
        L.refh_md2j.restype = C.c_float
        L.refh_md_lower_bound.argtypes = [C.c_float, C.c_int, C.c_int]
        L.refh_md_lower_bound.restype = C.c_float
        L.refh_recommended_sketch_size.argtypes = [C.c_int, C.c_float, C.c_int64, C.c_uint64]
        L.refh_recommended_sketch_size.restype = C.c_int64
        L.refh_sketch_sequence.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int]
        L.refh_add_minmers.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int64]
        L.refh_add_minmers.restype = C.c_int64
        L.refh_map_fragment.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int,
                                        vp, C.POINTER(C.c_int), C.POINTER(C.c_float),
                                        vp, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int),
                                        vp, C.c_int, C.POINTER(C.c_int),
                                        vp, vp, C.c_int, C.POINTER(C.c_int),
                                        vp, C.c_int, C.POINTER(C.c_int)]
        L.refh_map_read.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_int, C.c_int, vp, C.c_int]
        _lib = L
    return _lib


def _bytes(seq):
    if isinstance(seq, np.ndarray):
        return seq.tobytes()
    if isinstance(seq, str):
        return seq.encode()
    return bytes(seq)


def ref_hash(seq, k):
    return int(lib().refh_hash(_bytes(seq), k))


def sketch_sequence(seq, k, s, seq_id=0):
    b = _bytes(seq)
    out = np.zeros(s + 1, dtype=minmer_dtype)
    n = lib().refh_sketch_sequence(b, len(b), k, s, seq_id, out.ctypes.data, len(out))
    assert n >= 0
    return out[:n].copy()


def add_minmers(seq, k, w, s, seq_id=0):
    b = _bytes(seq)
    cap = max(1024, 4 * (len(b) // max(1, w) + 2) * (s + 2) + 4 * len(b) // 10)
    while True:
        out = np.zeros(cap, dtype=minmer_dtype)
        n = lib().refh_add_minmers(b, len(b), k, w, s, seq_id, out.ctypes.data, cap)
        if n >= 0:
            return out[:n].copy()
        cap = -n + 16


def parse_only(args):
    """skch::Parameters the reference's own parseandSave produces for a command line (no Sketch, the files are not read)"""
    L = lib()
    L.refh_parse.restype = C.c_void_p
    L.refh_parse.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
    argv = (C.c_char_p * len(args))(*[a.encode() for a in args])
    h = L.refh_parse(len(args), argv)
    p = OrcParams()
    L.refh_params(C.c_void_p(h), C.byref(p))
    L.refh_close(C.c_void_p(h))
    return p


class RefSession:
    """The reference's Sketch + Map built from a FASTA, with stage-level access."""

    def __init__(self, args):
        L = lib()
        argv = (C.c_char_p * len(args))(*[a.encode() for a in args])
        self.h = L.refh_open(len(args), argv)
        self.p = OrcParams()
        L.refh_params(self.h, C.byref(self.p))
        self.n_contigs = L.refh_n_contigs(self.h)
        self.contig_names = [L.refh_contig_name(self.h, i).decode() for i in range(self.n_contigs)]
        self.contig_len = np.array([L.refh_contig_len(self.h, i) for i in range(self.n_contigs)], dtype=np.int32)

    def close(self):
        if self.h:
            lib().refh_close(self.h)
            self.h = None

    def index(self):
        L = lib()
        n = L.refh_index_size(self.h)
        if n == 0:
            return np.zeros(0, dtype=minmer_dtype)
        buf = (C.c_char * (n * 24)).from_address(L.refh_index_data(self.h))
        return np.frombuffer(buf, dtype=minmer_dtype, count=n).copy()

    def lookup(self):
        """(keys ascending, offsets[n+1], points, is_freq)"""
        L = lib()
        n = L.refh_lookup_build(self.h)
        keys = np.frombuffer((C.c_char * (n * 8)).from_address(L.refh_lookup_keys(self.h)), dtype=np.uint64).copy()
        offs = np.frombuffer((C.c_char * ((n + 1) * 8)).from_address(L.refh_lookup_offs(self.h)), dtype=np.uint64).copy()
        npts = int(offs[-1])
        pts = np.frombuffer((C.c_char * (npts * 24)).from_address(L.refh_lookup_pts(self.h)), dtype=ipoint_dtype).copy()
        fr = np.frombuffer((C.c_char * (n * 4)).from_address(L.refh_lookup_isfreq(self.h)), dtype=np.int32).copy()
        return keys, offs, pts, fr.astype(np.uint8)

    def freq_threshold(self):
        return lib().refh_freq_threshold(self.h)

    def cutoffs(self):
        out = np.zeros(1002, dtype=np.int32)
        n = lib().refh_cutoffs(self.h, out.ctypes.data, len(out))
        return out[:n].copy()

    def min_hits_table(self, smax=None):
        smax = self.p.sketchSize if smax is None else smax
        L = lib()
        return np.array([0] + [L.refh_min_hits(s, self.p.kmerSize, self.p.percentageIdentity) for s in range(1, smax + 1)],
                        dtype=np.int32)

    def map_fragment(self, name, seq, full_len=None, seq_counter=0, ip_cap=1 << 20):
        b = _bytes(seq)
        S = self.p.sketchSize
        sk = np.zeros(S + 1, dtype=minmer_dtype)
        n_sk, kc = C.c_int(), C.c_float()
        ip = np.zeros(ip_cap, dtype=ipoint_dtype)
        n_ip, mh = C.c_int64(), C.c_int()
        l1 = np.zeros(4096, dtype=l1_dtype)
        n_l1 = C.c_int()
        l2 = np.zeros(8192, dtype=l2_dtype)
        l2c = np.zeros(8192, dtype=np.int32)
        n_l2 = C.c_int()
        mp = np.zeros(4096, dtype=mapping_dtype)
        n_mp = C.c_int()
        rc = lib().refh_map_fragment(self.h, name.encode(), b, len(b), len(b) if full_len is None else full_len,
                                     seq_counter, sk.ctypes.data, C.byref(n_sk), C.byref(kc), ip.ctypes.data, ip_cap,
                                     C.byref(n_ip), C.byref(mh), l1.ctypes.data, len(l1), C.byref(n_l1),
                                     l2.ctypes.data, l2c.ctypes.data, len(l2), C.byref(n_l2), mp.ctypes.data, len(mp),
                                     C.byref(n_mp))
        return dict(rc=rc, sketch=sk[: n_sk.value].copy(), kmerComplexity=kc.value,
                    points=ip[: min(n_ip.value, ip_cap)].copy(), n_points=n_ip.value, minimumHits=mh.value,
                    l1=l1[: n_l1.value].copy(), l2=l2[: n_l2.value].copy(), l2_cand=l2c[: n_l2.value].copy(),
                    mappings=mp[: n_mp.value].copy())

    def map_read(self, name, seq, seq_counter=0):
        b = _bytes(seq)
        out = np.zeros(8192, dtype=mapping_dtype)
        n = lib().refh_map_read(self.h, name.encode(), b, len(b), seq_counter, out.ctypes.data, len(out))
        return out[:n].copy()
