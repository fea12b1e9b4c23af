"""The bulk (memory-mapped, multi-threaded) FASTA reader of skch::Map gives exactly what the line reader gives
(reference seqiter.hpp:20-111 semantics), and declines what it cannot map (gzip, FASTQ). CPU only."""
import gzip
import os

import numpy as np
import pytest

from mashmap_b200 import hostlib


def write(path, data):
    with open(path, "wb") as f:
        f.write(data)
    return path


def random_bases(rng, n):
    return bytes(np.frombuffer(b"ACGTacgtNn", dtype=np.uint8)[rng.integers(0, 10, n)])


@pytest.mark.parametrize("threads", [1, 3, 16])
def test_bulk_reader_equals_line_reader(tmp_path, threads):
    rng = np.random.default_rng(3)
    recs = []
    for i in range(400):
        n = int(rng.choice([0, 1, 17, 59, 60, 61, 1000, 5000, 20011]))
        recs.append((f"read{i} some description\twith tab", random_bases(rng, n)))
    # single-line records, trailing newline
    a = b"".join(b">" + n.encode() + b"\n" + s + b"\n" for n, s in recs)
    # 60-column lines, an empty line here and there, '>' in the middle of a header, no newline at the end of the file
    b = b""
    for n, s in recs:
        b += b">" + n.encode() + b" >not a record\n"
        for o in range(0, len(s), 60):
            b += s[o : o + 60] + b"\n"
        if len(s) % 7 == 0:
            b += b"\n"
    b = b.rstrip(b"\n")
    # CRLF line ends (the reference keeps the '\r': so do both readers)
    c = a.replace(b"\n", b"\r\n")
    # headers without any space, one record only, record without sequence lines at the very end
    d = b">x\nACGT\n>y"
    for name, data in (("a.fa", a), ("b.fa", b), ("c.fa", c), ("d.fa", d)):
        p = write(str(tmp_path / name), data)
        diff, n_rec, n_bases = hostlib.fasta_readers_diff(p, threads)
        assert diff == 0, (name, diff)
        assert n_rec > 0
    assert hostlib.fasta_readers_diff(write(str(tmp_path / "a2.fa"), a), threads)[1] == len(recs)


def test_bulk_reader_declines_gzip_and_fastq(tmp_path):
    fa = b">r1\nACGTACGT\n>r2\nTTTT\n"
    gz = str(tmp_path / "x.fa.gz")
    with gzip.open(gz, "wb") as f:
        f.write(fa)
    fq = write(str(tmp_path / "x.fq"), b"@r1\nACGT\n+\nIIII\n@r2\nGG\n+\nII\n")
    assert hostlib.fasta_readers_diff(gz)[0] == -1 and hostlib.fasta_readers_diff(gz)[1] == 2
    assert hostlib.fasta_readers_diff(fq)[0] == -1 and hostlib.fasta_readers_diff(fq)[1] == 2


def test_host_packer_matches_the_format_statement():
    """seqio::pack_bases (AVX2 + scalar tail) == the numpy statement of the nibble format (capi.pack_bases): every byte
    value, lower case, IUPAC codes, odd lengths, lengths around the 32-byte vector width"""
    import numpy as np

    from mashmap_b200 import capi

    rng = np.random.default_rng(3)
    allbytes = np.arange(256, dtype=np.uint8)
    dna = np.frombuffer(b"ACGTacgtNnRYKM", dtype=np.uint8)[rng.integers(0, 14, size=5000)]
    for seq in (allbytes, dna, dna[:31], dna[:32], dna[:33], dna[:63], dna[:64], dna[:65], dna[:1], dna[:4999]):
        want = capi.pack_bases(seq)
        got = hostlib.pack_bases(seq)
        if len(seq) & 1:  # the unused high nibble of the last byte is an N in both
            assert want[-1] >> 4 == 8 and got[-1] >> 4 == 8
        assert np.array_equal(want, got), len(seq)
