"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: read sharding, the record gather, and the broadcast
helper on a CPU byte buffer. The NCCL / device-memory path itself runs in bench.py --gpus N on the GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mashmap_b200 import dist as mdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = mdist.shard_reads(1001, rank, world)
        # records: row r = (global read id, rank) for this rank's reads that "mapped" (every third one)
        ids = np.arange(lo, hi)[::3]
        rec = torch.from_numpy(np.stack([ids, np.full_like(ids, rank)] + [ids * 0] * 8, axis=1).astype(np.int32))
        parts, total = mdist.gather_records(dist, rec, world)
        blob = torch.arange(257, dtype=torch.uint8) if rank == 0 else torch.zeros(257, dtype=torch.uint8)
        dist.broadcast(blob, 0)
        q.put((rank, lo, hi, total, [p[:, 0].tolist() for p in parts], [p[:, 1].tolist() for p in parts], blob.tolist()))
    finally:
        dist.destroy_process_group()


def test_shard_reads_covers_everything():
    for n in (0, 1, 7, 1000, 1001):
        for w in (1, 2, 3, 8):
            blocks = [mdist.shard_reads(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_gather_and_broadcast_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    results.sort()
    (r0, lo0, hi0, tot0, ids0, rk0, blob0), (r1, lo1, hi1, tot1, ids1, rk1, blob1) = results
    assert (lo0, hi0, lo1, hi1) == (0, 501, 501, 1001)
    expect0, expect1 = list(range(0, 501, 3)), list(range(501, 1001, 3))
    assert tot0 == tot1 == len(expect0) + len(expect1)
    assert ids0 == ids1 == [expect0, expect1]          # every rank sees every rank's records, untruncated, in rank order
    assert rk0[0] == [0] * len(expect0) and rk0[1] == [1] * len(expect1)
    assert blob0 == blob1 == list(range(256)) + [0]


def test_host_threads_of_a_rank_follow_its_numa_node_and_the_quota(monkeypatch):
    """bench.py sizes a rank's host threads to the CPUs it can really run on: the CPUs of its GPU's NUMA node shared by the
    ranks bound to that node, and never more than the cgroup quota's share"""
    import importlib.util
    import sys

    spec = importlib.util.spec_from_file_location("bench_for_test", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    sys.modules["bench_for_test"] = bench
    spec.loader.exec_module(bench)
    monkeypatch.setattr(bench, "host_cpu_info", lambda: {"visible": 128, "affinity": 64, "cgroup_quota_cpus": 96.0})
    assert bench.host_threads_for_rank(8, (0, 64, 4)) == 12   # two nodes x 4 GPUs, 96-CPU quota
    assert bench.host_threads_for_rank(4, (0, 64, 4)) == 16   # 4 GPUs on one node
    assert bench.host_threads_for_rank(2, (0, 64, 2)) == 32
    monkeypatch.setattr(bench, "host_cpu_info", lambda: {"visible": 128, "affinity": 128, "cgroup_quota_cpus": 16.0})
    assert bench.host_threads_for_rank(8, (1, 64, 4)) == 2    # a 16-CPU quota: 2 per rank
    assert bench.host_threads_for_rank(8, None) == 2          # no binding: usable CPUs / ranks
    assert bench.host_threads_for_rank(1, None) == 16


def test_algorithmic_bytes_of_l1_and_l2_are_counted_on_the_records():
    """bench.py's K2 / K3 roofline entries: SURVEY 8(d)'s B2 and B3 counted on a step's own records and the index records
    (n_scan = index entries from lower_bound((seqId, rangeStart - L - 1)) to the last one with wpos <= rangeEnd)"""
    import importlib.util
    import sys

    from mashmap_b200 import capi

    spec = importlib.util.spec_from_file_location("bench_for_test2", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    sys.modules["bench_for_test2"] = bench
    spec.loader.exec_module(bench)
    # index: contig 0 has entries at wpos 0, 10, ..., 990; contig 1 at 5, 15, ..., 495
    idx_seq = np.concatenate([np.zeros(100, np.int32), np.ones(50, np.int32)])
    idx_wpos = np.concatenate([np.arange(100, dtype=np.int32) * 10, np.arange(50, dtype=np.int32) * 10 + 5])
    seg_res = np.zeros(3, dtype=capi.segres_dtype)
    seg_res["sketch_size"] = [20, 20, 7]
    seg_res["n_points"] = [40, 12, 0]
    cands = np.zeros(3, dtype=capi.l1_dtype)
    cands["seqId"] = [0, 1, 0]
    cands["rangeStartPos"] = [300, 20, 50]
    cands["rangeEndPos"] = [400, 60, 55]
    cands["segment"] = [0, 0, 1]
    loci = np.zeros(4, dtype=capi.l2_dtype)
    r = bench.algorithmic_bytes_l1_l2(seg_res, cands, loci, idx_seq, idx_wpos, seg_len=100)
    # candidate 0: wpos in [199, 400] on contig 0 -> 200..400 = 21 entries; candidate 1: [-81 -> 0, 60] on contig 1 -> 5..55 = 6;
    # candidate 2: [-51 -> 0, 55] on contig 0 -> 0..50 = 6
    assert r["index_entries_scanned"] == 21 + 6 + 6
    assert r["B2_bytes"] == 32 * 47 + 24 * 52 + 16 * 3
    assert r["B3_bytes"] == 24 * 33 + 24 * (20 + 20 + 20) + 24 * 4
