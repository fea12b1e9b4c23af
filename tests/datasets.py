"""Small deterministic data sets shared by the tests (generated, never read from /root/reference)."""
from __future__ import annotations

import os

import numpy as np

from mashmap_b200 import synth


def make_random_set(workdir, tag="rnd", n_contigs=3, contig_len=400_000, n_reads=40, read_len=10_000, seed=11):
    """noisy reads vs a random genome (the config-2 regime: few seed hits, wide L1 ranges)"""
    ref = os.path.join(workdir, f"{tag}_ref.fa")
    qry = os.path.join(workdir, f"{tag}_reads.fa")
    genome = synth.random_genome(n_contigs, contig_len, seed=seed)
    names = [f"ctg{i}" for i in range(n_contigs)]
    reads, truth = synth.simulate_reads(genome, n_reads, read_len, 0.02, 0.14, seed=seed + 1)
    rng = np.random.default_rng(seed + 2)
    # edge cases: a short read (< segLength), a read with N runs, lower-case, a non-multiple length, all-N
    extra = [genome[0][1000:4200].copy(), genome[1][50_000:62_345].copy(), genome[2][10_000:21_000].copy(),
             np.full(6000, ord("N"), dtype=np.uint8), genome[0][300_000:300_018].copy()]
    extra[1][3000:3400] = ord("N")
    extra[1][7000] = ord("n")
    extra[2] = np.frombuffer(extra[2].tobytes().lower(), dtype=np.uint8).copy()
    extra[2][::997] = ord("R")
    reads = reads + extra
    rnames = [f"read{i}" for i in range(len(reads))]
    synth.write_fasta(ref, names, genome)
    synth.write_fasta(qry, rnames, reads)
    synth.write_fai(qry, rnames, reads)
    return dict(ref=ref, qry=qry, genome=genome, names=names, reads=reads, rnames=rnames, truth=truth)


def make_panel_set(workdir, tag="panel", n_strains=4, n_chrom=2, chrom_len=150_000, seed=21):
    """repeat-rich all-vs-all panel (the yeast regime: many seed hits, narrow L1 ranges)"""
    ref = os.path.join(workdir, f"{tag}.fa")
    names, contigs = synth.panel_genome(n_strains, n_chrom, chrom_len, divergence=0.01, seed=seed)
    synth.write_fasta(ref, names, contigs)
    synth.write_fai(ref, names, contigs)
    return dict(ref=ref, qry=ref, genome=contigs, names=names, reads=contigs, rnames=names)


def make_assembly_set(workdir, tag="asm", n_contigs=3, contig_len=700_000, seed=31):
    """assembly-vs-assembly (BASELINE config 5 shape): a second genome = the first with 3 % SNPs, 0.3 % short indels,
    one inversion and one translocation; contigs of the query are named differently"""
    rng = np.random.default_rng(seed)
    genome = synth.random_genome(n_contigs, contig_len, seed=seed)
    other = [synth.mutate(c, 0.033, rng, ratio=(30, 2, 1)) for c in genome]
    a, b = 150_000, 270_000  # inversion on contig 0
    other[0] = np.concatenate([other[0][:a], synth.revcomp(other[0][a:b]), other[0][b:]])
    x, y = other[1][400_000:520_000].copy(), other[2][100_000:220_000].copy()  # translocation 1 <-> 2
    other[1] = np.concatenate([other[1][:400_000], y, other[1][520_000:]])
    other[2] = np.concatenate([other[2][:100_000], x, other[2][220_000:]])
    ref = os.path.join(workdir, f"{tag}_a.fa")
    qry = os.path.join(workdir, f"{tag}_b.fa")
    names = [f"a_ctg{i}" for i in range(n_contigs)]
    qnames = [f"b_ctg{i}" for i in range(n_contigs)]
    synth.write_fasta(ref, names, genome)
    synth.write_fasta(qry, qnames, other)
    synth.write_fai(qry, qnames, other)
    return dict(ref=ref, qry=qry, genome=genome, names=names, reads=other, rnames=qnames)


def make_hifi_set(workdir, tag="hifi", n_contigs=3, contig_len=500_000, n_reads=36, read_len=20_000, seed=41):
    """HiFi-like reads (BASELINE config 4 shape): 20 kb, 0.5 % error"""
    genome = synth.random_genome(n_contigs, contig_len, seed=seed)
    reads, truth = synth.simulate_reads(genome, n_reads, read_len, 0.004, 0.006, seed=seed + 1)
    ref = os.path.join(workdir, f"{tag}_ref.fa")
    qry = os.path.join(workdir, f"{tag}_reads.fa")
    names = [f"ctg{i}" for i in range(n_contigs)]
    rnames = [f"hifi{i}" for i in range(len(reads))]
    synth.write_fasta(ref, names, genome)
    synth.write_fasta(qry, rnames, reads)
    synth.write_fai(qry, rnames, reads)
    return dict(ref=ref, qry=qry, genome=genome, names=names, reads=reads, rnames=rnames, truth=truth)


def make_big_random_set(workdir, tag="big", n_contigs=8, contig_len=4_000_000, n_reads=60, read_len=10_000, seed=51):
    """32 Mbp random reference + noisy 10 kb reads (BASELINE config 3 shape: --dense --pi 95, s = 199, real point
    densities of a reference two orders of magnitude larger than the other fixtures)"""
    genome = synth.random_genome(n_contigs, contig_len, seed=seed)
    reads, truth = synth.simulate_reads(genome, n_reads, read_len, 0.01, 0.08, seed=seed + 1)
    ref = os.path.join(workdir, f"{tag}_ref.fa")
    qry = os.path.join(workdir, f"{tag}_reads.fa")
    names = [f"ctg{i}" for i in range(n_contigs)]
    rnames = [f"read{i}" for i in range(len(reads))]
    synth.write_fasta(ref, names, genome)
    synth.write_fasta(qry, rnames, reads)
    synth.write_fai(qry, rnames, reads)
    return dict(ref=ref, qry=qry, genome=genome, names=names, reads=reads, rnames=rnames, truth=truth)


def make_repeat_set(workdir, tag="rep", seed=61, n_copies=420, element_len=4000, spacer=1000, tandem_period=5200,
                    tandem_copies=6):
    """a repeat-dense reference: `n_copies` interspersed copies (1 % diverged) of one 4 kb element, each followed by a
    short unique spacer -- a fragment that covers a copy gathers more interval points than any fixed-size buffer of the
    L1 kernel (CTA path, its global slices, the bump-allocated pool) -- and an exact tandem array of period 5.2 kb: one
    L1 candidate spans all its copies, L2 finds `tandem_copies` equally good loci more than a fragment apart (more than
    the two locus slots of the stream kernel -> general L2 kernel)."""
    rng = np.random.default_rng(seed)
    element = synth.random_sequence(element_len, rng)
    parts = []
    for _ in range(n_copies):
        parts.append(synth.mutate(element, 0.01, rng, ratio=(1, 0, 0))[:element_len])
        parts.append(synth.random_sequence(spacer, rng))
    c0 = np.concatenate(parts)
    unit = synth.random_sequence(tandem_period, rng)
    c1 = np.concatenate([synth.random_sequence(60_000, rng)] + [unit] * tandem_copies + [synth.random_sequence(60_000, rng)])
    genome = [c0, c1]
    names = ["repeats", "tandem"]
    reads = []
    for i in (3, 57, 200, 411):  # fragments starting inside interspersed copies
        at = i * (element_len + spacer) + 200
        reads.append(c0[at : at + 10_000].copy())
    reads.append(unit[100:5100].copy())                   # one fragment inside the tandem unit
    reads.append(np.concatenate([unit, unit])[2600:12600].copy())
    reads.append(synth.revcomp(unit[50:5050]))
    rnames = [f"rep{i}" for i in range(len(reads))]
    ref = os.path.join(workdir, f"{tag}_ref.fa")
    qry = os.path.join(workdir, f"{tag}_reads.fa")
    synth.write_fasta(ref, names, genome)
    synth.write_fasta(qry, rnames, reads)
    synth.write_fai(qry, rnames, reads)
    return dict(ref=ref, qry=qry, genome=genome, names=names, reads=reads, rnames=rnames)
