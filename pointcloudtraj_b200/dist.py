"""Multi-GPU plumbing: one process per GPU, cloud index replicated once, query batches split per rank.

The path has exactly one exchange step -- replicating the built index (``pc_index_broadcast``: one ncclBroadcast of
the tree array over NVLink/NVSwitch).  Queries are independent given the read-only index, so every rank answers a
contiguous slice of the batch (``pc_shard_range``) and writes its slice of the outputs; there is no other
collective on the data path (results are gathered only when a caller asks, e.g. for verification).

``torch.distributed`` is used only for the rendezvous (shipping the 128-byte NCCL unique id, barriers) -- with the
``gloo`` backend on CPU in the tests, ``nccl`` on the GPU box.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib as L
from .index import PointCloudIndex, shard_range


def env_rank():
    """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0")))


def exchange_unique_id(rank: int, world: int, make_id):
    """Rank 0 creates the NCCL unique id (make_id() -> 128 bytes); everybody receives it through torch.distributed."""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(L.PC_NCCL_UNIQUE_ID_BYTES, dtype=torch.uint8)
    if rank == 0:
        buf = torch.frombuffer(bytearray(make_id()), dtype=torch.uint8).clone()
    if world > 1:
        if dist.get_backend() == "nccl":
            dev = torch.device("cuda", torch.cuda.current_device())
            t = buf.to(dev)
            dist.broadcast(t, 0)
            buf = t.cpu()
        else:
            dist.broadcast(buf, 0)
    return bytes(buf.numpy().tobytes())


class Replicator:
    """Owns a pc_comm and replicates built indexes from a root rank to every rank's handle."""

    def __init__(self, rank: int, world: int, device: int):
        self._L = L.load()
        self.rank, self.world, self.device = rank, world, device

        def make_id():
            b = C.create_string_buffer(L.PC_NCCL_UNIQUE_ID_BYTES)
            rc = self._L.pc_comm_unique_id(b)
            if rc != L.PC_OK:
                raise L.PcError(rc, self._L.pc_last_error(None).decode())
            return b.raw

        uid = exchange_unique_id(rank, world, make_id)
        h = C.c_void_p()
        rc = self._L.pc_comm_init(C.byref(h), rank, world, uid, device)
        if rc != L.PC_OK:
            raise L.PcError(rc, self._L.pc_last_error(None).decode())
        self._h = h

    def broadcast(self, index: PointCloudIndex, root: int = 0):
        rc = self._L.pc_index_broadcast(index._h, self._h, root)
        if rc != L.PC_OK:
            raise L.PcError(rc, self._L.pc_last_error(index._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.pc_comm_destroy(self._h)
            self._h = None

    __del__ = close


def sharded_call(fn, q, rank: int, world: int):
    """Apply fn (e.g. ``lambda part: index.radius(part, params)``) to this rank's contiguous slice of q.
    Returns (begin, end, result)."""
    b, e = shard_range(len(q), rank, world)
    return b, e, fn(q[b:e])


def gather_slices(local: np.ndarray, m: int, rank: int, world: int):
    """Verification helper: assemble the per-rank output slices into the full array on every rank."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local
    parts = [None] * world
    dist.all_gather_object(parts, (rank, local))
    out = np.empty((m,) + local.shape[1:], dtype=local.dtype)
    for r, a in parts:
        b, e = shard_range(m, r, world)
        assert e - b == len(a)
        out[b:e] = a
    return out
