"""Multi-GPU contract of the hot path: every query point is independent given the cloud, so
rank r owns the contiguous point-index range ``shard_range(n, r, world)`` of the replicated cloud and the only
collective is an all-gather of the outputs into file order.

``GatherBuffers`` removes every copy around that collective: the full-size output arrays are allocated once, each rank's
pipeline writes its range straight into its slot of them (``local_views``), and ``all_gather`` gathers IN PLACE
(``all_gather_into_tensor`` with the input a view of the output, one call per dtype: normals stay float32 [.,3] = 12
B/point, curvatures float64 [.,2] = 16 B/point -- nothing is widened, zero-filled, concatenated or re-cast).  Slots are
``capacity = ceil(n / world)`` rows apart; when ``n`` is not a multiple of ``world`` the last rows of the short slots are
padding and ``file_order`` compacts them (the only case that copies).
"""
from __future__ import annotations

from typing import Optional

import torch


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced point-index range [begin, end) of ``rank`` (first n % world ranks get one more)."""
    base, rem = divmod(int(n), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_capacity(n: int, world: int) -> int:
    return (int(n) + int(world) - 1) // int(world)


class GatherBuffers:
    def __init__(self, n: int, world: int, device, want_curv: bool = True):
        self.n, self.world = int(n), int(world)
        self.cap = shard_capacity(n, world)
        self.normals = torch.empty((self.world * self.cap, 3), dtype=torch.float32, device=device)
        self.curv = torch.empty((self.world * self.cap, 2), dtype=torch.float64, device=device) if want_curv else None
        self.ragged = self.world * self.cap != self.n
        if self.ragged:       # padding rows are gathered too: keep them finite
            self.normals.zero_()
            if self.curv is not None:
                self.curv.zero_()

    def local_views(self, rank: int):
        """(normals view [count,3], curv view [count,2] or None) of this rank's slot: pass them to the pipeline as its
        output buffers."""
        b, e = shard_range(self.n, rank, self.world)
        o = rank * self.cap
        return self.normals[o:o + (e - b)], (self.curv[o:o + (e - b)] if self.curv is not None else None)

    def all_gather(self, rank: int):
        """In-place all-gather of every slot (NCCL over NVLink on GPUs, gloo in the CPU tests).  Asynchronous w.r.t. the
        host on CUDA tensors (ordered on the current stream by torch.distributed)."""
        if self.world == 1:
            return
        import torch.distributed as dist
        o = rank * self.cap
        dist.all_gather_into_tensor(self.normals, self.normals[o:o + self.cap])
        if self.curv is not None:
            dist.all_gather_into_tensor(self.curv, self.curv[o:o + self.cap])

    def file_order(self):
        """(normals [n,3], curv [n,2] or None) in file order: views when n is a multiple of world, else compacted."""
        if not self.ragged:
            return self.normals, self.curv
        pieces_n, pieces_c = [], []
        for r in range(self.world):
            b, e = shard_range(self.n, r, self.world)
            o = r * self.cap
            pieces_n.append(self.normals[o:o + (e - b)])
            if self.curv is not None:
                pieces_c.append(self.curv[o:o + (e - b)])
        return torch.cat(pieces_n, 0), (torch.cat(pieces_c, 0) if self.curv is not None else None)

    def bytes_per_rank(self) -> int:
        return self.cap * (12 + (16 if self.curv is not None else 0))


def gather_outputs(local_normals: torch.Tensor, local_curv: Optional[torch.Tensor], n: int, rank: int, world: int):
    """Convenience form for callers that already hold their shard in separate tensors: copies it into a fresh
    ``GatherBuffers`` slot, gathers, returns file-order (normals, curv).  The hot path uses ``GatherBuffers`` directly
    and skips the copy."""
    bufs = GatherBuffers(n, world, local_normals.device, want_curv=local_curv is not None)
    vn, vc = bufs.local_views(rank)
    vn.copy_(local_normals)
    if vc is not None:
        vc.copy_(local_curv)
    bufs.all_gather(rank)
    return bufs.file_order()
