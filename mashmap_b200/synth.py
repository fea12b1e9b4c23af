"""Deterministic synthetic inputs: random reference genomes, ONT/HiFi-like reads,
a repeat-rich multi-genome panel. numpy only; bench.py has a torch/GPU variant of the read simulator
for the full-size configuration.
"""
from __future__ import annotations

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[:] = ord("N")
for a, b in zip(b"ACGTacgt", b"TGCAtgca"):
    _COMP[a] = b


def random_sequence(n, rng):
    return ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]


def random_genome(n_contigs, contig_len, seed=1):
    """uniform i.i.d. ACGT contigs (list of uint8 arrays)."""
    rng = np.random.default_rng(seed)
    return [random_sequence(contig_len, rng) for _ in range(n_contigs)]


def revcomp(seq):
    return _COMP[seq[::-1]]


def mutate(seq, err, rng, ratio=(4, 3, 3)):
    """Apply substitutions / insertions / deletions at total rate `err` (sub:ins:del = ratio)."""
    n = len(seq)
    tot = float(sum(ratio))
    r = rng.random(n)
    p_sub, p_ins, p_del = (err * ratio[0] / tot, err * ratio[1] / tot, err * ratio[2] / tot)
    is_sub = r < p_sub
    is_ins = (r >= p_sub) & (r < p_sub + p_ins)
    is_del = (r >= p_sub + p_ins) & (r < p_sub + p_ins + p_del)
    out = seq.copy()
    # substitution: a different base
    shift = rng.integers(1, 4, size=n, dtype=np.uint8)
    code = np.searchsorted(ACGT, seq)  # ACGT sorted ascending in ASCII: A C G T
    code = np.clip(code, 0, 3)
    out[is_sub] = ACGT[(code[is_sub] + shift[is_sub]) % 4]
    counts = np.ones(n, dtype=np.int64)
    counts[is_del] = 0
    counts[is_ins] = 2
    idx = np.repeat(np.arange(n), counts)
    res = out[idx]
    # second copy of an inserted position becomes a random base
    first = np.ones(len(idx), dtype=bool)
    first[1:] = idx[1:] != idx[:-1]
    ins_pos = ~first
    res[ins_pos] = random_sequence(int(ins_pos.sum()), rng)
    return res


def simulate_reads(genome, n_reads, read_len, err_lo, err_hi, seed=2):
    """Reads of exactly read_len bases drawn uniformly from the genome (either strand), per-read error
    rate ~ U[err_lo, err_hi]. Returns (list of uint8 arrays, truth list of (contig, start, strand, err))."""
    rng = np.random.default_rng(seed)
    lens = np.array([len(c) for c in genome], dtype=np.int64)
    span = int(read_len * 1.25) + 64
    ok = lens >= span
    w = np.where(ok, lens - span + 1, 0).astype(np.float64)
    w /= w.sum()
    reads, truth = [], []
    for _ in range(n_reads):
        c = int(rng.choice(len(genome), p=w))
        start = int(rng.integers(0, lens[c] - span + 1))
        err = float(rng.uniform(err_lo, err_hi))
        src = genome[c][start : start + span]
        strand = 1
        if rng.random() < 0.5:
            src = revcomp(src)
            strand = -1
        m = mutate(src, err, rng)
        while len(m) < read_len:  # extremely deletion-heavy draw: pad from fresh sequence
            m = np.concatenate([m, random_sequence(read_len - len(m), rng)])
        reads.append(np.ascontiguousarray(m[:read_len]))
        truth.append((c, start, strand, err))
    return reads, truth


def panel_genome(n_strains, n_chrom, chrom_len, divergence=0.01, seed=5, repeat_len=6000, n_repeats=6):
    """A repeat-rich panel imitating the 8-yeast-genome fixture: one ancestral genome with interspersed
    repeat copies, `n_strains` mutated copies of every chromosome. Returns (names, contigs)."""
    rng = np.random.default_rng(seed)
    ancestor = []
    rep = random_sequence(repeat_len, rng)
    for c in range(n_chrom):
        s = random_sequence(chrom_len, rng)
        for _ in range(n_repeats):  # paste diverged repeat copies
            at = int(rng.integers(0, chrom_len - repeat_len))
            cp = mutate(rep, 0.02, rng, ratio=(1, 0, 0))
            s[at : at + repeat_len] = cp[:repeat_len]
        ancestor.append(s)
    names, contigs = [], []
    for st in range(n_strains):
        for c in range(n_chrom):
            names.append(f"strain{st}#1#chr{c + 1}")
            if st == 0:
                contigs.append(ancestor[c].copy())
            else:
                contigs.append(mutate(ancestor[c], divergence, rng, ratio=(8, 1, 1)))
    return names, contigs


def write_fasta(path, names, seqs, width=80):
    with open(path, "wb") as f:
        for name, s in zip(names, seqs):
            f.write(b">" + name.encode() + b"\n")
            s = np.ascontiguousarray(s, dtype=np.uint8)
            n = len(s)
            full = (n // width) * width
            if full:
                block = np.empty((n // width, width + 1), dtype=np.uint8)
                block[:, :width] = s[:full].reshape(-1, width)
                block[:, width] = 10
                f.write(block.tobytes())
            if n > full:
                f.write(s[full:].tobytes() + b"\n")


def write_fai(path_fasta, names, seqs, width=80):
    """A .fai next to the FASTA so the reference skips its length pre-pass (computeMap.hpp:281-304)."""
    off = 0
    with open(path_fasta + ".fai", "w") as f:
        for name, s in zip(names, seqs):
            off += len(name) + 2
            n = len(s)
            f.write(f"{name}\t{n}\t{off}\t{width}\t{width + 1}\n")
            off += n + (n + width - 1) // width


def split_segments(read_lens, seg_length, kmer_size):
    """Fragmenting rule of Map::mapModule (computeMap.hpp:587-671): reads shorter than k are skipped,
    reads <= seg_length map whole, longer reads give floor(len/L) disjoint fragments plus one
    overlapping tail fragment [len-L, len) when len % L != 0.
    Returns (read_index, start) arrays and per-fragment length."""
    ridx, start, length = [], [], []
    for i, n in enumerate(read_lens):
        n = int(n)
        if n < kmer_size:
            continue
        if n <= seg_length:
            ridx.append(i); start.append(0); length.append(n)
            continue
        k = n // seg_length
        for j in range(k):
            ridx.append(i); start.append(j * seg_length); length.append(seg_length)
        if n % seg_length:
            ridx.append(i); start.append(n - seg_length); length.append(seg_length)
    return np.array(ridx, dtype=np.int64), np.array(start, dtype=np.int64), np.array(length, dtype=np.int32)
