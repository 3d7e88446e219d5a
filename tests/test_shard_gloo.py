"""Host-side multi-GPU plumbing on CPU: world_size=2 over gloo.  The per-shard compute is stood in by the
oracle (the CUDA path cannot run here and has no CPU fallback); what is under test is the partition and
the scatter / gather-v logic of tamp_b200/shard.py."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from tamp_b200 import shard  # noqa: E402


def test_partition_balanced():
    for n in (0, 1, 7, 8, 1 << 20, (1 << 20) + 3):
        for world in (1, 2, 4, 8):
            parts = shard.partition(n, world)
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_streams, stride, chunks, result_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    h = oracle.Harness("port")
    rows = torch.from_numpy(h.generate(oracle.TEXT, 0, n_streams, stride)) if rank == 0 else None
    slot = (stride * 9 + 7) // 8 + 32

    def comp(x):  # stand-in for batch.compress_batch + batch.compact on this rank's chunk
        out, sizes, status, _ = h.compress(x.numpy(), window=10, extended=True, threads=2, out_stride=slot)
        packed = np.concatenate([out[i, :sizes[i]] for i in range(len(sizes))] + [np.zeros(0, np.uint8)])
        return torch.from_numpy(packed.copy()), torch.from_numpy(sizes.astype(np.int32)), torch.from_numpy(status)

    def decomp(frames, offsets, sizes):  # stand-in for batch.decompress_packed
        n = sizes.numel()
        rows_in = np.zeros((n, slot), np.uint8)
        f, o, z = frames.numpy(), offsets.numpy(), sizes.numpy()
        for i in range(n):
            rows_in[i, :z[i]] = f[o[i]:o[i] + z[i]]
        back, bsz, bst, _ = h.decompress(rows_in, z.astype(np.uint32), stride, window_bits_max=10, threads=2)
        return torch.from_numpy(back), torch.from_numpy(bsz.astype(np.int32)), torch.from_numpy(bst)

    res = shard.compress_sharded(comp, rows, n_streams, stride, slot, chunks=chunks)
    back = shard.decompress_sharded(decomp, res, n_streams, stride, chunks=chunks)
    if rank == 0:
        exp, esz, est, _ = h.compress(rows.numpy(), window=10, extended=True, threads=2, out_stride=slot)
        ok = bool((res.sizes.numpy() == esz).all() and (res.status.numpy() == 0).all())
        data, off = res.data.numpy(), res.offsets.numpy()
        for i in range(n_streams):
            ok = ok and bool((data[off[i]:off[i] + esz[i]] == exp[i, :esz[i]]).all())
        out, osz, ost, moved = back
        ok = ok and bool((out.numpy() == rows.numpy()).all() and (osz.numpy() == stride).all()) and moved > 0
        # only payload crossed the interconnect on the way back: frames + 5 bytes of size / status per stream
        parts = shard.partition(n_streams, world)
        remote = sum(int(esz[a:b].sum()) + 5 * (b - a) + (b - a) * stride for r, (a, b) in enumerate(parts) if r != 0)
        ok = ok and res.nvlink_bytes == remote
        Path(result_path).write_text("ok" if ok else "mismatch")
    else:
        assert res is None and back is None
    dist.barrier()
    dist.destroy_process_group()


def test_segments_cover_the_batch_in_order():
    for n in (0, 5, 101, 4096):
        for world in (1, 2, 3, 8):
            for chunks in (1, 2, 4):
                flat = [ab for r in shard.segments(n, world, chunks) for ab in r]
                assert flat[0][0] == 0 and flat[-1][1] == n
                assert all(flat[i][1] == flat[i + 1][0] for i in range(len(flat) - 1))


@pytest.mark.parametrize("n_streams,chunks", [(101, 3), (64, 1), (5, 4)])
def test_scatter_compress_gatherv_and_back_world2(tmp_path, n_streams, chunks):
    """compress_sharded / decompress_sharded over gloo, world_size 2: the root ends up with every stream's frame at
    offsets[i] (bit-exact with a single-process run), the way back restores the rows, and the byte count that crossed
    the interconnect is payload only."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    result = tmp_path / "result.txt"
    mp.spawn(_worker, args=(2, port, n_streams, 512, chunks, str(result)), nprocs=2, join=True)
    assert result.read_text() == "ok"
