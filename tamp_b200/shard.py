"""Multi-GPU sharding of a batch of independent streams (BASELINE.json config 5).

Streams carry no cross-stream state, so a batch shards by stream index with no data-path collective inside the codec.
This module is the plumbing for callers whose data starts (and ends) on ONE rank: the root scatters input rows to the
ranks, every rank compresses its shard, and the root gathers the COMPRESSED BYTES — compacted on the device, so that
only payload crosses NVLink (gather-v: sizes first, then the packed frames at prefix-sum offsets) — and the mirror
image for decompression (scatter packed frames, gather fixed-length rows).

Pipeline: every rank's shard is cut into `chunks` pieces.  All receives of the scatter are posted up front and the
root's sends are issued chunk-major, so chunk c + 1 arrives while chunk c is being compressed and chunk c - 1's
frames travel back (NCCL point-to-point runs on its own CUDA stream; `Work.wait()` only orders the compute stream
behind it).  The only host synchronisation is one read of a chunk's compressed byte count (the size of the send).

`torch.distributed` point-to-point calls: NCCL over NVLink / NVSwitch on GPUs, gloo in the CPU tests.  The per-chunk
codec call is passed in (the product passes `tamp_b200.batch` functions; the CPU tests an oracle stand-in), so the
partition / offset logic is testable without a GPU.  Bound to state with the numbers: a single root's NVLink egress
(~900 GB/s per direction nominal) carries (world - 1) / world of the input, i.e. scatter time ~ bytes / egress.
"""
from __future__ import annotations

from dataclasses import dataclass
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


def segments(n_streams: int, world: int, chunks: int) -> list[list[tuple[int, int]]]:
    """segments[r][c] = [start, end) of chunk c of rank r's shard (stream order = rank-major, chunk-minor)."""
    return [[(lo + a, lo + b) for a, b in partition(hi - lo, chunks)] for lo, hi in partition(n_streams, world)]


def _wait_all(works):
    for w in works:
        w.wait()


def _batch(ops):
    return dist.batch_isend_irecv(ops) if ops else []


@dataclass
class Packed:
    """Frames of a batch in one buffer on the root: frame i is data[offsets[i] : offsets[i] + sizes[i]].  Segments sit
    at fixed slots (first stream index x slot stride), so the buffer has gaps: it is not a prefix-sum layout."""
    data: torch.Tensor      # 1-D uint8
    offsets: torch.Tensor   # int64 [n_streams]
    sizes: torch.Tensor     # int32 [n_streams]
    status: torch.Tensor    # int8 [n_streams]
    nvlink_bytes: int = 0   # bytes this rank moved over the interconnect for the call (payload + sizes + status)


def scatter_rows(rows: torch.Tensor | None, n_streams: int, stride: int, *, src: int = 0, device=None,
                 dtype=torch.uint8, chunks: int = 1):
    """Rank `src` holds `rows` (n_streams, stride).  Returns (shard, works): `shard` is this rank's rows, `works[c]`
    the pending receives of chunk c (empty on `src`, whose shard is a view of `rows`)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    segs = segments(n_streams, world, chunks)
    lo, hi = segs[rank][0][0], segs[rank][-1][1]
    if rank == src:
        works = []
        for c in range(chunks):  # chunk-major: every peer gets its chunk 0 first
            works += _batch([dist.P2POp(dist.isend, rows[a:b], r) for r in range(world)
                             if r != src for a, b in [segs[r][c]] if b > a])
        return rows[lo:hi], [[] for _ in range(chunks)], works
    mine = torch.empty((hi - lo, stride), dtype=dtype, device=device)
    works = [_batch([dist.P2POp(dist.irecv, mine[a - lo:b - lo], src)] if b > a else []) for a, b in segs[rank]]
    return mine, works, []


def gather_rows(mine: torch.Tensor, n_streams: int, *, dst: int = 0) -> torch.Tensor | None:
    """Inverse of scatter_rows for fixed-stride per-stream results (sizes, status, output rows)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    parts = partition(n_streams, world)
    if rank == dst:
        full = torch.empty((n_streams,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
        lo, hi = parts[rank]
        full[lo:hi] = mine
        _wait_all(_batch([dist.P2POp(dist.irecv, full[a:b], r) for r, (a, b) in enumerate(parts) if r != dst and b > a]))
        return full
    if mine.shape[0]:
        _wait_all(_batch([dist.P2POp(dist.isend, mine.contiguous(), dst)]))
    return None


def compress_sharded(fn: Callable[[torch.Tensor], tuple[torch.Tensor, torch.Tensor, torch.Tensor]],
                     rows: torch.Tensor | None, n_streams: int, stride: int, slot_stride: int, *, root: int = 0,
                     device=None, chunks: int = 4) -> Packed | None:
    """scatter rows -> fn(chunk) on every rank -> gather-v of the compacted frames on `root`.

    `fn(rows_chunk)` returns (packed 1-D uint8 holding the chunk's frames back to back, sizes int32, status int8).
    `slot_stride` >= the worst-case compressed size of one stream: segment (rank, chunk) lands at byte offset
    first_stream_index * slot_stride of the root's buffer.  Returns a Packed on `root`, None elsewhere."""
    rank, world = dist.get_rank(), dist.get_world_size()
    segs = segments(n_streams, world, chunks)
    shard, recv_works, send_works = scatter_rows(rows, n_streams, stride, src=root, device=device, chunks=chunks)
    lo = segs[rank][0][0]
    dev = shard.device
    moved = 0
    if rank == root:
        data = torch.empty(max(n_streams * slot_stride, 1), dtype=torch.uint8, device=dev)
        sizes = torch.empty(n_streams, dtype=torch.int32, device=dev)
        status = torch.empty(n_streams, dtype=torch.int8, device=dev)
    results, pending = [], []
    totals = torch.zeros((chunks, world), dtype=torch.int64, device=dev)
    for c in range(chunks + 1):
        if c < chunks:  # enqueue chunk c's codec work before waiting on chunk c - 1's byte count
            a, b = segs[rank][c]
            _wait_all(recv_works[c])
            results.append(fn(shard[a - lo:b - lo]) if b > a else None)
        if c == 0:
            continue
        k = c - 1  # ship chunk k
        res = results[k]
        mine_total = torch.tensor([res[0].numel() if res is not None else 0], dtype=torch.int64, device=dev)
        row = list(totals[k].unbind(0))
        dist.all_gather([t.view(1) for t in row], mine_total)  # every rank learns every segment's byte count
        if rank == root:
            tot = totals[k].tolist()
            ops = []
            for r in range(world):
                a, b = segs[r][k]
                if b <= a:
                    continue
                base = a * slot_stride
                if r == root:
                    data[base:base + tot[r]] = res[0]
                    sizes[a:b] = res[1]
                    status[a:b] = res[2]
                else:
                    ops += [dist.P2POp(dist.irecv, sizes[a:b], r), dist.P2POp(dist.irecv, status[a:b], r)]
                    if tot[r]:
                        ops.append(dist.P2POp(dist.irecv, data[base:base + tot[r]], r))
                    moved += tot[r] + 5 * (b - a)
            pending += _batch(ops)
        elif res is not None:
            ops = [dist.P2POp(dist.isend, res[1], root), dist.P2POp(dist.isend, res[2], root)]
            if res[0].numel():
                ops.append(dist.P2POp(dist.isend, res[0], root))
            moved += res[0].numel() + 5 * res[1].numel()
            pending += _batch(ops)
    _wait_all(pending)
    _wait_all(send_works)
    if rank != root:
        return None
    moved += sum((b - a) * stride for r in range(world) if r != root for a, b in segs[r])
    # frame offsets: slot base of the stream's segment + prefix sum of the sizes inside the segment
    s64 = sizes.to(torch.int64)
    excl = torch.cumsum(s64, 0) - s64
    offsets = torch.empty(n_streams, dtype=torch.int64, device=dev)
    for r in range(world):
        for a, b in segs[r]:
            if b > a:
                offsets[a:b] = excl[a:b] - excl[a] + a * slot_stride
    return Packed(data, offsets, sizes, status, moved)


def decompress_sharded(fn: Callable[[torch.Tensor, torch.Tensor, torch.Tensor], tuple[torch.Tensor, torch.Tensor, torch.Tensor]],
                       packed: Packed | None, n_streams: int, out_stride: int, *, root: int = 0, device=None,
                       chunks: int = 4):
    """scatter frames -> fn(frames, offsets, sizes) on every rank -> gather of the fixed-length output rows on `root`.

    The root sends every segment's frames (its slot of `packed.data`, exactly the bytes in use) and sizes; a rank
    decompresses chunk by chunk with local prefix-sum offsets.  `fn` returns (rows (n, out_stride) uint8, sizes int32,
    status int8).  Returns (rows, sizes, status, nvlink_bytes) on `root`, None elsewhere."""
    rank, world = dist.get_rank(), dist.get_world_size()
    segs = segments(n_streams, world, chunks)
    lo, hi = segs[rank][0][0], segs[rank][-1][1]
    moved = 0
    # byte counts of all segments: one small broadcast from the root
    if rank == root:
        dev = packed.data.device
        s64 = packed.sizes.to(torch.int64)
        csum = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(s64, 0)])
        seg_bytes = torch.stack([torch.stack([csum[b] - csum[a] for a, b in segs[r]]) for r in range(world)])
    else:
        dev = torch.device(device) if device is not None else torch.device("cpu")
        seg_bytes = torch.empty((world, chunks), dtype=torch.int64, device=dev)
    dist.broadcast(seg_bytes, root)
    nbytes = seg_bytes.tolist()
    sends = []
    if rank == root:
        for c in range(chunks):
            ops = []
            for r in range(world):
                a, b = segs[r][c]
                if r == root or b <= a:
                    continue
                base = int(packed.offsets[a].item())
                ops.append(dist.P2POp(dist.isend, packed.sizes[a:b], r))
                if nbytes[r][c]:
                    ops.append(dist.P2POp(dist.isend, packed.data[base:base + nbytes[r][c]], r))
                moved += nbytes[r][c] + 4 * (b - a)
            sends += _batch(ops)
        out = torch.empty((n_streams, out_stride), dtype=torch.uint8, device=dev)
        osz = torch.empty(n_streams, dtype=torch.int32, device=dev)
        ost = torch.empty(n_streams, dtype=torch.int8, device=dev)
        # the receives of every peer's rows are posted now, behind the sends on the NCCL stream: a peer's chunk travels back
        # as soon as it is decompressed, whatever the root's own shard is doing
        ops = []
        for c in range(chunks):
            for r in range(world):
                a, b = segs[r][c]
                if r == root or b <= a:
                    continue
                ops += [dist.P2POp(dist.irecv, out[a:b], r), dist.P2POp(dist.irecv, osz[a:b], r),
                        dist.P2POp(dist.irecv, ost[a:b], r)]
                moved += (b - a) * (out_stride + 5)
        recvs = _batch(ops)
    else:
        frames, sizes, works = [], [], []
        for c, (a, b) in enumerate(segs[rank]):
            sizes.append(torch.empty(b - a, dtype=torch.int32, device=dev))
            frames.append(torch.empty(max(nbytes[rank][c], 1), dtype=torch.uint8, device=dev))
            ops = []
            if b > a:
                ops.append(dist.P2POp(dist.irecv, sizes[c], root))
                if nbytes[rank][c]:
                    ops.append(dist.P2POp(dist.irecv, frames[c][:nbytes[rank][c]], root))
            works.append(_batch(ops))
        mine = torch.empty((hi - lo, out_stride), dtype=torch.uint8, device=dev)
        msz = torch.empty(hi - lo, dtype=torch.int32, device=dev)
        mst = torch.empty(hi - lo, dtype=torch.int8, device=dev)
    pending = []
    for c, (a, b) in enumerate(segs[rank]):
        if b <= a:
            continue
        if rank == root:
            base = int(packed.offsets[a].item())
            fr, sz = packed.data[base:base + nbytes[rank][c]], packed.sizes[a:b]
        else:
            _wait_all(works[c])
            fr, sz = frames[c][:nbytes[rank][c]], sizes[c]
        s64 = sz.to(torch.int64)
        rows, rsz, rst = fn(fr, torch.cumsum(s64, 0) - s64, sz)
        if rank == root:
            out[a:b], osz[a:b], ost[a:b] = rows, rsz, rst
        else:
            mine[a - lo:b - lo], msz[a - lo:b - lo], mst[a - lo:b - lo] = rows, rsz, rst
            pending += _batch([dist.P2POp(dist.isend, mine[a - lo:b - lo], root), dist.P2POp(dist.isend, msz[a - lo:b - lo], root),
                               dist.P2POp(dist.isend, mst[a - lo:b - lo], root)])
            moved += (b - a) * (out_stride + 5)
    if rank == root:
        _wait_all(recvs)
        _wait_all(sends)
        return out, osz, ost, moved
    _wait_all(pending)
    return None
