"""Multi-GPU sharding of a batch of independent streams.

Streams carry no cross-stream state, so a batch shards by stream index with no data-path collective.
This module holds the host-side plumbing: the partition, and — for callers whose data starts on one
rank — a scatter of input rows and a gather of (variable-length) results over `torch.distributed`
point-to-point calls (NCCL over NVLink on GPUs, gloo in the CPU tests).  The per-shard compute is passed
in as a callable; the product passes `tamp_b200.batch.compress_batch` / `decompress_batch`.
"""
from __future__ import annotations

from typing import Callable

import torch
import torch.distributed as dist


def partition(n_streams: int, world: int) -> list[tuple[int, int]]:
    """Contiguous [start, end) ranges, sizes differing by at most one stream."""
    base, extra = divmod(n_streams, world)
    out, start = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((start, start + n))
        start += n
    return out


def scatter_rows(rows: torch.Tensor | None, n_streams: int, stride: int, *, src: int = 0, device=None,
                 dtype=torch.uint8) -> torch.Tensor:
    """Rank `src` holds `rows` (n_streams, stride); every rank returns its shard (grouped P2P sends)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    parts = partition(n_streams, world)
    lo, hi = parts[rank]
    if rank == src:
        ops = [dist.P2POp(dist.isend, rows[a:b].contiguous(), r) for r, (a, b) in enumerate(parts)
               if r != src and b > a]
        mine = rows[lo:hi].clone()
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return mine
    mine = torch.empty((hi - lo, stride), dtype=dtype, device=device)
    if hi > lo:
        for w in dist.batch_isend_irecv([dist.P2POp(dist.irecv, mine, src)]):
            w.wait()
    return mine


def gather_rows(mine: torch.Tensor, n_streams: int, *, dst: int = 0) -> torch.Tensor | None:
    """Inverse of scatter_rows for fixed-stride per-stream results (sizes, status, output rows)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    parts = partition(n_streams, world)
    if rank == dst:
        full = torch.empty((n_streams,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
        lo, hi = parts[rank]
        full[lo:hi] = mine
        ops = [dist.P2POp(dist.irecv, full[a:b], r) for r, (a, b) in enumerate(parts) if r != dst and b > a]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return full
    if mine.shape[0]:
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, mine.contiguous(), dst)]):
            w.wait()
    return None


def run_sharded(fn: Callable[[torch.Tensor], tuple[torch.Tensor, torch.Tensor, torch.Tensor]],
                rows: torch.Tensor | None, n_streams: int, stride: int, *, root: int = 0, device=None):
    """scatter -> fn(shard) -> gather.  `fn` returns (out_rows, sizes, status) for its shard.
    Returns (out_rows, sizes, status) for the whole batch on `root`, None elsewhere."""
    shard = scatter_rows(rows, n_streams, stride, src=root, device=device)
    out, sizes, status = fn(shard)
    full_out = gather_rows(out, n_streams, dst=root)
    full_sizes = gather_rows(sizes, n_streams, dst=root)
    full_status = gather_rows(status, n_streams, dst=root)
    if dist.get_rank() == root:
        return full_out, full_sizes, full_status
    return None
