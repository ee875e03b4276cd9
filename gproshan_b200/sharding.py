"""Multi-GPU plumbing for batched solves: independent source sets are block-partitioned across ranks (one
process per GPU, mesh replicated) and the distance rows are gathered once at the end — the only collective on
the path (NCCL over NVLink on GPUs; the same code runs on gloo/CPU tensors for the host-logic tests)."""
from __future__ import annotations

import numpy as np


def shard_bounds(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block partition; the first n_items % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sources(sources, rank: int, world: int, per_rank: int | None = None) -> np.ndarray:
    """This rank's sources. per_rank fixes the shard size (weak scaling: rank r takes [r*per_rank, (r+1)*per_rank))."""
    src = np.ascontiguousarray(sources, dtype=np.uint32)
    if per_rank is not None:
        if per_rank * world > src.size:
            raise ValueError("not enough sources for per_rank * world")
        return src[rank * per_rank:(rank + 1) * per_rank]
    lo, hi = shard_bounds(src.size, rank, world)
    return src[lo:hi]


def gather_rows(rows, world: int, out=None):
    """all_gather of equally sized (B_local, V) row blocks into (world*B_local, V), rank-major = source order."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return rows
    if out is None:
        out = torch.empty((rows.shape[0] * world,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    dist.all_gather_into_tensor(out, rows.contiguous())
    return out


def gather_rows_ragged(rows, n_items: int, rank: int, world: int):
    """Gather when n_items is not a multiple of world: pad shards to the largest, gather, drop the padding."""
    import torch
    if world == 1:
        return rows
    sizes = [shard_bounds(n_items, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    pad[:rows.shape[0]] = rows
    full = gather_rows(pad, world)
    keep = torch.cat([torch.arange(r * mx, r * mx + (hi - lo)) for r, (lo, hi) in enumerate(sizes)]).to(full.device)
    return full.index_select(0, keep)
