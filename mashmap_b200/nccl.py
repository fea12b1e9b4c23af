"""ctypes binding of libmashmap_nccl.so (include/mashmap_b200_nccl.h): the product's own multi-GPU entry points --
one NCCL broadcast of the index image, one all-gather of mapping records -- for a host program with one process per GPU.
The 128-byte NCCL unique id that rank 0 creates travels over whatever channel the host program has; bench.py and the
tests use torch.distributed (gloo / nccl object broadcast) for that bootstrap and for nothing on the data path."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmashmap_nccl.so")
ID_BYTES = 128
_lib = None

EXPORTED_SYMBOLS = ["mm_comm_unique_id", "mm_comm_create", "mm_comm_destroy", "mm_comm_last_error", "mm_index_broadcast",
                    "mm_records_allgather", "mm_index_replicate"]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        capi.lib()  # libmashmap_b200.so first (the add-on is written on its C ABI)
        L = C.CDLL(LIB_PATH)
        vp, u64 = C.c_void_p, C.c_uint64
        L.mm_comm_unique_id.argtypes = [vp]
        L.mm_comm_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
        L.mm_comm_destroy.argtypes = [vp]
        L.mm_comm_last_error.argtypes = [vp]
        L.mm_comm_last_error.restype = C.c_char_p
        L.mm_index_broadcast.argtypes = [vp, vp, C.c_int, C.POINTER(u64)]
        L.mm_records_allgather.argtypes = [vp, vp, u64, C.c_uint32, vp, u64, vp]
        L.mm_index_replicate.argtypes = [vp, C.POINTER(vp), C.c_int]
        _lib = L
    return _lib


def unique_id() -> bytes:
    buf = (C.c_uint8 * ID_BYTES)()
    rc = lib().mm_comm_unique_id(buf)
    if rc != capi.MM_OK:
        raise capi.MashmapError(rc, lib().mm_comm_last_error(None).decode())
    return bytes(buf)


class Comm:
    """one rank of the communicator (ncclCommInitRank on `device`)"""

    def __init__(self, uid: bytes, n_ranks: int, rank: int, device: int):
        assert len(uid) == ID_BYTES
        self._L = lib()
        self.n_ranks, self.rank, self.device = n_ranks, rank, device
        h = C.c_void_p()
        buf = (C.c_uint8 * ID_BYTES).from_buffer_copy(uid)
        rc = self._L.mm_comm_create(buf, n_ranks, rank, device, C.byref(h))
        if rc != capi.MM_OK:
            raise capi.MashmapError(rc, self._L.mm_comm_last_error(None).decode())
        self._h = h

    def _check(self, rc):
        if rc != capi.MM_OK:
            raise capi.MashmapError(rc, self._L.mm_comm_last_error(self._h).decode())

    def index_broadcast(self, ctx_handle, root=0) -> int:
        """ctx_handle: a capi.Context or a raw mm_ctx* (int / c_void_p). Returns the image size in bytes."""
        h = ctx_handle._h if hasattr(ctx_handle, "_h") else C.c_void_p(ctx_handle if isinstance(ctx_handle, int) else ctx_handle.value)
        n = C.c_uint64()
        self._check(self._L.mm_index_broadcast(h, self._h, root, C.byref(n)))
        return int(n.value)

    def records_allgather(self, records: np.ndarray):
        """records: [n, ...] array of fixed-size rows on this rank. Returns (all ranks' rows in rank order, counts[n_ranks])."""
        r = np.ascontiguousarray(records)
        row_bytes = int(r.dtype.itemsize * (np.prod(r.shape[1:]) if r.ndim > 1 else 1))
        counts = np.zeros(self.n_ranks, dtype=np.uint64)
        cap = max(int(len(r)) * self.n_ranks + int(len(r)) // 2, 1024)
        while True:
            # the gathered rows land in a pinned buffer kept between calls (a fresh pageable array per call costs page
            # faults on hundreds of MB and a staged device->host copy at a fraction of the PCIe rate)
            need = cap * row_bytes
            if getattr(self, "_pin", None) is None or self._pin.nbytes < need:
                self._pin = capi.PinnedBuffer(need)
            out = self._pin.array[:need].view(r.dtype).reshape((cap,) + r.shape[1:])
            rc = self._L.mm_records_allgather(self._h, r.ctypes.data, len(r), row_bytes, out.ctypes.data, cap, counts.ctypes.data)
            if rc == capi.MM_ECAPACITY:
                cap = int(counts.sum()) + 16
                continue
            self._check(rc)
            return out[: int(counts.sum())], counts

    def close(self):
        if self._h:
            self._L.mm_comm_destroy(self._h)
            self._h = None


def create_with_torch(dist, rank: int, world: int, device_index: int) -> Comm:
    """bootstrap over an initialised torch.distributed process group: rank 0 makes the id, everyone joins"""
    obj = [unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    return Comm(obj[0], world, rank, device_index)


def index_replicate(src_ctx, dst_ctxs):
    """one process, several devices: copy the index image of src_ctx to every context of dst_ctxs (grouped ncclBroadcast)"""
    arr = (C.c_void_p * len(dst_ctxs))(*[c._h for c in dst_ctxs])
    rc = lib().mm_index_replicate(src_ctx._h, arr, len(dst_ctxs))
    if rc != capi.MM_OK:
        raise capi.MashmapError(rc, lib().mm_comm_last_error(None).decode())
