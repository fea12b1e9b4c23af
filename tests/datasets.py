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
