"""Size-independent properties at a scale the CPU oracle cannot check read by read (60,000 reads x 10 kb = 120,000
segments vs a 100 Mbp reference): the sketches are sorted distinct bottom-s sets, candidate / locus records are
well-formed, reads land where they were simulated from, and the result does not depend on how the batch is cut into
parts (upload chunks, lanes, sub-batch size) nor on re-running it."""
import os

import numpy as np
import pytest

from conftest import have_gpu

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_gpu(), reason="no GPU")]

K, SEG, S, PI = 19, 5000, 220, 0.85
N_READS, READ_LEN, N_CONTIGS, CONTIG_LEN = 60_000, 10_000, 16, 6_250_000


@pytest.fixture(scope="module")
def world():
    import torch

    from mashmap_b200 import hostlib, synth_gpu

    dev = torch.device("cuda:0")
    ref = synth_gpu.random_reference(N_CONTIGS, CONTIG_LEN, seed=5, device=dev)
    reads, truth = synth_gpu.simulate_reads(ref, N_READS, READ_LEN, 0.02, 0.14, seed=6, chunk=8192)
    offs = np.arange(N_CONTIGS + 1, dtype=np.uint64) * np.uint64(CONTIG_LEN)
    hi = hostlib.HostIndex.build(ref.cpu().numpy().reshape(-1), offs, K, SEG, S, threads=os.cpu_count() or 8)
    t = tuple(truth[k].cpu().numpy() for k in ("contig", "start", "strand"))
    reads_h = reads.cpu().numpy()
    del ref, reads
    torch.cuda.empty_cache()
    yield dict(hi=hi, reads=reads_h, truth=t)
    hi.close()


def map_all(world, env):
    from mashmap_b200 import hostlib

    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        bm = hostlib.BatchMapper(world["hi"], pi=PI, device=0, threads=min(32, os.cpu_count() or 8))
        batch = bm.make_batch(N_READS, READ_LEN)
        batch.fill(world["reads"].reshape(-1), threads=min(32, os.cpu_count() or 8))
        info = bm.map(batch)
        res, paf = bm.results(), bm.paf()
        info2 = bm.map(batch)  # the same batch again on the warm pipeline
        assert bm.paf() == paf and info2["mappings"] == info["mappings"]
        batch.close()
        bm.close()
        return res, paf
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_reads_land_on_their_true_locus_and_partitioning_is_invisible(world):
    res, paf = map_all(world, {"MM_SUB_BATCH_BASES": str(64 << 20)})     # 10 parts, three lanes
    res2, paf2 = map_all(world, {"MM_SUB_BATCH_BASES": str(1 << 40)})    # one part, one lane, no pipeline
    res3, paf3 = map_all(world, {"MM_SUB_BATCH_BASES": str(7_000_000)})  # 86 small parts
    assert paf == paf2 == paf3 and len(paf) > 0
    contig_of, start_of, strand_of = world["truth"]
    q = res[:, 0]
    assert len(np.unique(q)) >= 0.999 * N_READS
    ok = (res[:, 3] == contig_of[q]) & (np.abs(res[:, 4].astype(np.int64) - start_of[q]) < 20_000) & (res[:, 6] == strand_of[q])
    assert ok.mean() >= 0.999
    assert np.all(res[:, 1] >= 0) and np.all(res[:, 2] <= READ_LEN) and np.all(res[:, 1] < res[:, 2])
    assert np.all(res[:, 4] >= 0) and np.all(res[:, 5] >= res[:, 4]) and np.all((res[:, 9] >= int(PI * 1e6) - 1_000_000) & (res[:, 9] <= 1_000_000))


def test_stage_records_are_well_formed(world):
    from mashmap_b200 import capi, synth

    n = 4000
    reads = world["reads"][:n]
    ridx, start, length = synth.split_segments([READ_LEN] * n, SEG, K)
    segs = np.zeros(len(ridx), dtype=capi.segment_dtype)
    segs["offset"] = ridx.astype(np.int64) * READ_LEN + start
    segs["length"] = length
    segs["seq_counter"] = ridx
    segs["name_id"] = -1
    segs["ref_group"] = -1
    ctx = capi.Context(kmer_size=K, seg_length=SEG, sketch_size=S)
    from mashmap_b200 import hostlib

    world["hi"].upload(ctx)
    ctx.tables_upload(hostlib.sketch_cutoffs(S, K), hostlib.min_hits_table(S, K, PI))
    seg_res, cands, loci = ctx.map_segments(reads.reshape(-1), segs)
    sk, cnt = ctx.batch_fetch_sketch()
    assert np.all(cnt <= S) and np.all(cnt == seg_res["sketch_size"])
    full = cnt == S  # every segment of a random sequence has far more than S distinct k-mers
    assert full.mean() > 0.999
    h = sk["hash"].reshape(len(segs), S)
    assert np.all(h[full][:, 1:] > h[full][:, :-1])                      # strictly ascending: sorted and distinct
    assert np.all(seg_res["sketch_max_hash"][full] == h[full][:, -1])
    w0, w1 = sk["wpos"].reshape(len(segs), S)[full], sk["wpos_end"].reshape(len(segs), S)[full]
    assert np.all(w0 >= 0) and np.all(w0 <= w1) and np.all(w1 <= SEG - K)
    assert np.all(np.isin(sk["strand"].reshape(len(segs), S)[full], (-1, 0, 1)))
    # candidates: one contiguous slice per segment, ranges inside the contig, overlap bounded by the sketch size
    assert int(seg_res["n_candidates"].sum()) == len(cands)
    assert np.all(cands["rangeStartPos"] <= cands["rangeEndPos"]) and np.all(cands["rangeStartPos"] >= 0)
    assert np.all(cands["rangeEndPos"] < CONTIG_LEN) and np.all(cands["seqId"] < N_CONTIGS)
    assert np.all(cands["intersectionSize"] <= seg_res["sketch_size"][cands["segment"]])
    assert np.all(cands["intersectionSize"] >= seg_res["minimum_hits"][cands["segment"]])
    # loci: inside the candidate's contig, ordered, shared count bounded by the sketch size
    for c in cands[:: max(1, len(cands) // 2000)]:
        l = loci[c["first_locus"] : c["first_locus"] + c["n_loci"]]
        assert np.all(l["seqId"] == c["seqId"]) and np.all(l["optimalStart"] <= l["optimalEnd"])
        assert np.all(l["sharedSketchSize"] <= seg_res["sketch_size"][c["segment"]])
        assert np.all(np.diff(l["optimalStart"]) > 0)
    ctx.close()
