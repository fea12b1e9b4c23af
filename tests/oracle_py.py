"""Test-side wrapper of oracle/libmm_oracle.so (the CPU restatement of the path). Test infrastructure only:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, never by the product."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

import refh  # record dtypes + OrcParams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "libmm_oracle.so")

_lib = None


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        vp = C.c_void_p
        L.orc_hash.argtypes = [C.c_char_p, C.c_int]
        L.orc_hash.restype = C.c_uint64
        L.orc_sketch_sequence.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int]
        L.orc_min_hits.argtypes = [C.c_int, C.c_int, C.c_float]
        L.orc_create.argtypes = [C.POINTER(refh.OrcParams)]
        L.orc_create.restype = vp
        L.orc_destroy.argtypes = [vp]
        L.orc_cutoffs.argtypes = [vp, vp, C.c_int]
        L.orc_set_index.argtypes = [vp, vp, C.c_uint64, vp, vp, C.c_uint64, vp, vp, vp, C.POINTER(C.c_char_p), vp, C.c_int]
        L.orc_map_fragment.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       vp, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_uint64),
                                       vp, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int),
                                       vp, C.c_int, C.POINTER(C.c_int), vp, vp, C.c_int, C.POINTER(C.c_int),
                                       vp, C.c_int, C.POINTER(C.c_int)]
        L.orc_map_read.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int]
        _lib = L
    return _lib


def _bytes(seq):
    return seq.tobytes() if isinstance(seq, np.ndarray) else (seq.encode() if isinstance(seq, str) else bytes(seq))


def sketch_sequence(seq, k, s, seq_id=0):
    b = _bytes(seq)
    out = np.zeros(s + 1, dtype=refh.minmer_dtype)
    n = lib().orc_sketch_sequence(b, len(b), k, s, seq_id, out.ctypes.data, len(out))
    assert n >= 0
    return out[:n].copy()


def default_params(k=19, seg_length=5000, sketch_size=220, pi=0.85, **kw):
    p = refh.OrcParams()
    p.kmerSize, p.segLength, p.sketchSize, p.alphabetSize = k, seg_length, sketch_size, 4
    p.percentageIdentity = pi
    p.filterMode, p.numMappingsForSegment, p.numMappingsForShortSequence = 1, 1, 1
    p.block_length, p.chain_gap, p.split, p.mergeMappings = seg_length, seg_length, 1, 1
    p.stage1_topANI_filter, p.ANIDiff, p.ANIDiffConf, p.stage2_full_scan = 1, 0.0, 0.999, 1
    p.keep_low_pct_id, p.kmer_pct_threshold, p.kmerComplexityThreshold = 1, 0.001, 0.0
    p.sparsity_hash_threshold = (1 << 64) - 1
    for name, v in kw.items():
        setattr(p, name, v)
    return p


class Oracle:
    def __init__(self, k=19, seg_length=5000, sketch_size=220, pi=0.85, params=None, **kw):
        self.p = params if params is not None else default_params(k, seg_length, sketch_size, pi, **kw)
        self.h = lib().orc_create(C.byref(self.p))

    def close(self):
        if self.h:
            lib().orc_destroy(self.h)
            self.h = None

    def cutoffs(self):
        out = np.zeros(1002, dtype=np.int32)
        n = lib().orc_cutoffs(self.h, out.ctypes.data, len(out))
        return out[:n].copy()

    def set_index(self, minmers, keys, offs, pts, is_freq, contig_len, contig_names=None, contig_group=None):
        self._keep = [np.ascontiguousarray(minmers, dtype=refh.minmer_dtype), np.ascontiguousarray(keys, dtype=np.uint64),
                      np.ascontiguousarray(offs, dtype=np.uint64), np.ascontiguousarray(pts, dtype=refh.ipoint_dtype),
                      np.ascontiguousarray(is_freq, dtype=np.uint8), np.ascontiguousarray(contig_len, dtype=np.int32)]
        m, k, o, p, f, cl = self._keep
        names = None
        if contig_names is not None:
            names = (C.c_char_p * len(contig_names))(*[n.encode() for n in contig_names])
        grp = None if contig_group is None else np.ascontiguousarray(contig_group, dtype=np.int32)
        lib().orc_set_index(self.h, m.ctypes.data, len(m), k.ctypes.data, o.ctypes.data, len(k), p.ctypes.data, f.ctypes.data,
                            cl.ctypes.data, names, None if grp is None else grp.ctypes.data, len(cl))

    def map_fragment(self, seq, seq_counter=0, full_len=None, name_id=-1, ref_group=-1, ip_cap=1 << 20):
        b = _bytes(seq)
        S = self.p.sketchSize
        sk = np.zeros(S + 1, dtype=refh.minmer_dtype)
        n_sk, kc, raw_n, raw_max = C.c_int(), C.c_float(), C.c_int(), C.c_uint64()
        ip = np.zeros(ip_cap, dtype=refh.ipoint_dtype)
        n_ip, mh = C.c_int64(), C.c_int()
        l1 = np.zeros(4096, dtype=refh.l1_dtype)
        n_l1 = C.c_int()
        l2 = np.zeros(8192, dtype=refh.l2_dtype)
        l2c = np.zeros(8192, dtype=np.int32)
        n_l2 = C.c_int()
        mp = np.zeros(4096, dtype=refh.mapping_dtype)
        n_mp = C.c_int()
        rc = lib().orc_map_fragment(self.h, b, len(b), len(b) if full_len is None else full_len, seq_counter, name_id, ref_group,
                                    sk.ctypes.data, C.byref(n_sk), C.byref(kc), C.byref(raw_n), C.byref(raw_max),
                                    ip.ctypes.data, ip_cap, C.byref(n_ip), C.byref(mh), l1.ctypes.data, len(l1), C.byref(n_l1),
                                    l2.ctypes.data, l2c.ctypes.data, len(l2), C.byref(n_l2), mp.ctypes.data, len(mp), C.byref(n_mp))
        return dict(rc=rc, sketch=sk[: n_sk.value].copy(), sketch_size=n_sk.value, kmerComplexity=kc.value,
                    raw_count=raw_n.value, raw_max_hash=raw_max.value, points=ip[: min(n_ip.value, ip_cap)].copy(),
                    n_points=n_ip.value, minimumHits=mh.value, l1=l1[: n_l1.value].copy(), l2=l2[: n_l2.value].copy(),
                    l2_cand=l2c[: n_l2.value].copy(), mappings=mp[: n_mp.value].copy())

    def map_read(self, seq, seq_counter=0, name_id=-1, ref_group=-1):
        b = _bytes(seq)
        out = np.zeros(8192, dtype=refh.mapping_dtype)
        n = lib().orc_map_read(self.h, b, len(b), seq_counter, name_id, ref_group, out.ctypes.data, len(out))
        return out[:n].copy()
