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
