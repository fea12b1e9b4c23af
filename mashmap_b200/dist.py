"""Multi-GPU plumbing: one process per GPU, torch.distributed for the collectives.
  * the reference index image goes from rank 0 to every rank with ONE broadcast (NCCL over NVLink on GPUs);
  * reads are sharded by contiguous blocks (a read's fragments stay on one rank, output order is preserved);
  * mapping records come back with one all_gather of counts + one all_gather of padded fixed-size records.
The functions take the process-group module so the same code runs on `gloo` with CPU tensors in the tests."""
from __future__ import annotations

import numpy as np
import torch


def shard_reads(n_reads: int, rank: int, world: int):
    """contiguous block [lo, hi) of reads for this rank, sizes differing by at most one"""
    base, rem = divmod(n_reads, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def wrap_device_memory(ptr: int, nbytes: int, device):
    """a torch uint8 tensor over raw device memory owned by the C ABI context (the index blob)"""

    class _Arr:
        pass

    a = _Arr()
    a.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
    return torch.as_tensor(a, device=device)


def broadcast_index(dist, ctx, rank: int, device):
    """rank 0 owns the index image (mm_index_blob); every other rank allocates it (mm_index_blob_alloc), receives it
    with one broadcast and adopts it (mm_index_adopt_blob). Returns the number of bytes moved."""
    nbytes = torch.zeros(1, dtype=torch.int64, device=device)
    ptr = 0
    if rank == 0:
        ptr, n = ctx.index_blob()
        nbytes[0] = n
    dist.broadcast(nbytes, 0)
    n = int(nbytes.item())
    if rank != 0:
        ptr = ctx.index_blob_alloc(n)
    blob = wrap_device_memory(ptr, n, device)
    dist.broadcast(blob, 0)
    # the collective runs on the process group's own stream and the C ABI context reads the image on ITS stream: wait for
    # the device before adopting (without this, ranks far down the broadcast tree read a stale header -- the round-1
    # 8-GPU failure "blob header mismatch")
    if getattr(device, "type", str(device)) == "cuda":
        torch.cuda.synchronize(device)
    if rank != 0:
        ctx.index_adopt_blob()
    return n


def gather_records(dist, records: torch.Tensor, world: int):
    """records: [n, w] int32 on this rank (n differs per rank). Returns (list of per-rank tensors trimmed to their true
    length, total count) on every rank."""
    device = records.device
    n = torch.tensor([records.shape[0]], dtype=torch.int64, device=device)
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    m = max(counts + [1])
    pad = torch.zeros((m, records.shape[1]), dtype=records.dtype, device=device)
    pad[: records.shape[0]] = records
    out = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return [o[:c] for o, c in zip(out, counts)], sum(counts)
