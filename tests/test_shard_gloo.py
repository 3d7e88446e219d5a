"""Host-side multi-GPU plumbing on CPU: world_size=2 over gloo.  The per-shard compute is stood in by the
oracle (the CUDA path cannot run here and has no CPU fallback); what is under test is the partition and
the scatter / gather-v logic of tamp_b200/shard.py."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
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


def _worker(rank, world, port, n_streams, stride, result_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    h = oracle.Harness("port")
    rows = torch.from_numpy(h.generate(oracle.TEXT, 0, n_streams, stride)) if rank == 0 else None

    def fn(x):  # stand-in for batch.compress_batch on this rank's shard
        out, sizes, status, _ = h.compress(x.numpy(), window=10, extended=True, threads=2)
        return torch.from_numpy(out), torch.from_numpy(sizes.astype(np.int32)), torch.from_numpy(status)

    res = shard.run_sharded(fn, rows, n_streams, stride)
    if rank == 0:
        out, sizes, status = res
        exp, esz, est, _ = h.compress(rows.numpy(), window=10, extended=True, threads=2)
        ok = bool((sizes.numpy() == esz).all() and (out.numpy() == exp).all() and (status.numpy() == 0).all())
        Path(result_path).write_text("ok" if ok else "mismatch")
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_scatter_compute_gather_world2(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    result = tmp_path / "result.txt"
    mp.spawn(_worker, args=(2, port, 101, 512, str(result)), nprocs=2, join=True)
    assert result.read_text() == "ok"
