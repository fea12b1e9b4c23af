"""torch/CUDA versions of the synthetic-input generators in synth.py for the full-size benchmark
configuration (3 Gbp reference, 1 M x 10 kb reads): same model (uniform i.i.d. ACGT reference; reads drawn
uniformly from either strand, per-read error rate ~ U[lo, hi], sub:ins:del = 4:3:3, exactly read_len bases),
generated on the device so that setting up 10 Gbp of reads takes seconds. Not bit-identical to synth.py."""
from __future__ import annotations

import torch

_ASCII = (65, 67, 71, 84)  # A C G T


def random_reference(n_contigs, contig_len, seed, device):
    """[n_contigs, contig_len] uint8 ASCII on `device`"""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lut = torch.tensor(_ASCII, dtype=torch.uint8, device=device)
    out = torch.empty((n_contigs, contig_len), dtype=torch.uint8, device=device)
    step = max(1, (256 << 20) // contig_len)
    for a in range(0, n_contigs, step):
        b = min(n_contigs, a + step)
        codes = torch.randint(0, 4, (b - a, contig_len), generator=g, device=device, dtype=torch.uint8)
        out[a:b] = lut[codes.long()]
    return out


def simulate_reads(ref, n_reads, read_len, err_lo, err_hi, seed, chunk=16384):
    """ref: [n_contigs, contig_len] uint8 ASCII on the device. Returns (reads [n_reads, read_len] uint8 on the
    device, truth dict of tensors: contig, start, strand, err)."""
    device = ref.device
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    n_contigs, contig_len = ref.shape
    span = int(read_len * 1.25) + 64
    lut = torch.tensor(_ASCII, dtype=torch.uint8, device=device)
    comp = torch.zeros(256, dtype=torch.uint8, device=device)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    out = torch.empty((n_reads, read_len), dtype=torch.uint8, device=device)
    t_contig = torch.empty(n_reads, dtype=torch.int32, device=device)
    t_start = torch.empty(n_reads, dtype=torch.int64, device=device)
    t_strand = torch.empty(n_reads, dtype=torch.int8, device=device)
    t_err = torch.empty(n_reads, dtype=torch.float32, device=device)
    ar = torch.arange(span, device=device)
    for a in range(0, n_reads, chunk):
        n = min(chunk, n_reads - a)
        contig = torch.randint(0, n_contigs, (n,), generator=g, device=device)
        start = torch.randint(0, contig_len - span + 1, (n,), generator=g, device=device)
        err = torch.rand(n, generator=g, device=device) * (err_hi - err_lo) + err_lo
        rev = torch.rand(n, generator=g, device=device) < 0.5
        src = ref[contig[:, None], start[:, None] + ar[None, :]]  # [n, span]
        src_rc = comp[src.flip(1).long()]
        src = torch.where(rev[:, None], src_rc, src)
        r = torch.rand((n, span), generator=g, device=device)
        p_sub, p_ins = err * 0.4, err * 0.3
        is_sub = r < p_sub[:, None]
        is_ins = (r >= p_sub[:, None]) & (r < (p_sub + p_ins)[:, None])
        is_del = (r >= (p_sub + p_ins)[:, None]) & (r < err[:, None])
        # substitution: one of the three other bases
        code = ((src >> 1) & 3).long()  # A0 C1 T2 G3
        order = torch.tensor([0, 1, 3, 2], device=device)  # code -> index in ACGT
        idx = order[code]
        shift = torch.randint(1, 4, (n, span), generator=g, device=device)
        sub_base = lut[(idx + shift) % 4]
        base = torch.where(is_sub, sub_base, src)
        counts = torch.ones((n, span), dtype=torch.int64, device=device)
        counts[is_del] = 0
        counts[is_ins] = 2
        pos = torch.cumsum(counts, dim=1) - counts  # output position of each source base
        buf = torch.zeros((n, read_len + 2), dtype=torch.uint8, device=device)
        keep = (counts > 0) & (pos < read_len)
        rows = torch.arange(n, device=device)[:, None].expand(n, span)
        buf[rows[keep], pos[keep]] = base[keep]
        ins_ok = is_ins & (pos + 1 < read_len)
        ins_base = lut[torch.randint(0, 4, (n, span), generator=g, device=device)]
        buf[rows[ins_ok], (pos + 1)[ins_ok]] = ins_base[ins_ok]
        # (a read shortened below read_len by deletions cannot happen: span = 1.25 * read_len + 64)
        out[a : a + n] = buf[:, :read_len]
        t_contig[a : a + n] = contig.int()
        t_start[a : a + n] = start
        t_strand[a : a + n] = torch.where(rev, -1, 1).to(torch.int8)
        t_err[a : a + n] = err
    return out, dict(contig=t_contig, start=t_start, strand=t_strand, err=t_err)


def mutated_genome(ref, out_len, snp, indel, n_inversions, n_translocations, seed, rows=4):
    """A second assembly of the same genome (BASELINE config 5): every contig of `ref` with substitutions at rate `snp`
    and short insertions / deletions at rate `indel` (half each), cut to exactly out_len bases, then `n_inversions` blocks
    of 50-200 kb reverse-complemented in place and `n_translocations` pairs of equally long blocks swapped between
    contigs. Returns [n_contigs, out_len] uint8 on the device."""
    device = ref.device
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    n_contigs, contig_len = ref.shape
    assert out_len <= contig_len - 64
    lut = torch.tensor(_ASCII, dtype=torch.uint8, device=device)
    comp = torch.zeros(256, dtype=torch.uint8, device=device)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    order = torch.tensor([0, 1, 3, 2], device=device)
    out = torch.empty((n_contigs, out_len), dtype=torch.uint8, device=device)
    for a in range(0, n_contigs, rows):
        src = ref[a : a + rows]
        n, span = src.shape
        r = torch.rand((n, span), generator=g, device=device)
        is_sub = r < snp
        is_ins = (r >= snp) & (r < snp + indel / 2)
        is_del = (r >= snp + indel / 2) & (r < snp + indel)
        idx = order[((src >> 1) & 3).long()]
        shift = torch.randint(1, 4, (n, span), generator=g, device=device)
        base = torch.where(is_sub, lut[(idx + shift) % 4], src)
        counts = torch.ones((n, span), dtype=torch.int32, device=device)
        counts[is_del] = 0
        counts[is_ins] = 2
        pos = torch.cumsum(counts, dim=1, dtype=torch.int64) - counts
        buf = torch.zeros((n, out_len + 2), dtype=torch.uint8, device=device)
        rowsx = torch.arange(n, device=device)[:, None].expand(n, span)
        keep = (counts > 0) & (pos < out_len)
        buf[rowsx[keep], pos[keep]] = base[keep]
        ins_ok = is_ins & (pos + 1 < out_len)
        ins_base = lut[torch.randint(0, 4, (n, span), generator=g, device=device)]
        buf[rowsx[ins_ok], (pos + 1)[ins_ok]] = ins_base[ins_ok]
        out[a : a + n] = buf[:, :out_len]
        del r, is_sub, is_ins, is_del, idx, shift, base, counts, pos, buf, rowsx, keep, ins_ok, ins_base
    cpu = torch.Generator()
    cpu.manual_seed(seed + 1)
    for _ in range(n_inversions):
        c = int(torch.randint(0, n_contigs, (1,), generator=cpu))
        ln = int(torch.randint(50_000, 200_001, (1,), generator=cpu))
        if ln + 2 >= out_len:
            continue
        st = int(torch.randint(0, out_len - ln, (1,), generator=cpu))
        out[c, st : st + ln] = comp[out[c, st : st + ln].flip(0).long()]
    for _ in range(n_translocations):
        c1, c2 = (int(x) for x in torch.randint(0, n_contigs, (2,), generator=cpu))
        ln = int(torch.randint(50_000, 200_001, (1,), generator=cpu))
        if ln + 2 >= out_len or c1 == c2:
            continue
        s1, s2 = (int(x) for x in torch.randint(0, out_len - ln, (2,), generator=cpu))
        tmp = out[c1, s1 : s1 + ln].clone()
        out[c1, s1 : s1 + ln] = out[c2, s2 : s2 + ln]
        out[c2, s2 : s2 + ln] = tmp
    return out
