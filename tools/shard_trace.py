#!/usr/bin/env python
"""Where does a sharded step (config 5, one window class) spend its time?  Runs compress_sharded / decompress_sharded
with host timers around their pieces (a synchronise at every mark: the sum is an upper bound of the pipelined step).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/shard_trace.py [window] [stream_len] [mib]
"""
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tamp_b200 import batch, shard  # noqa: E402

w = int(sys.argv[1]) if len(sys.argv) > 1 else 15
n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
mib = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n_streams = (mib << 20) // n
slot = (batch.compress_bound(n, 8) + 15) // 16 * 16
rows = batch.synth(0, 0, n_streams, n, device=dev) if rank == 0 else None
marks = []


def mark(name):
    torch.cuda.synchronize()
    marks.append((name, time.perf_counter()))


def comp(xc):
    mark("chunk arrived")
    r = batch.compress_batch(xc, window=w, literal=8, extended=False, out_stride=slot)
    mark("compress kernel")
    packed, _ = batch.compact(r)
    mark("compact")
    return packed, r.sizes, r.status


def decomp(frames, offsets, sizes):
    mark("frames arrived")
    r = batch.decompress_packed(frames, offsets, sizes, n, window_bits_max=w)
    mark("decompress kernel")
    return r.data, r.sizes, r.status


for it in range(2):
    marks.clear()
    dist.barrier()
    mark("start")
    p = shard.compress_sharded(comp, rows, n_streams, n, slot, device=dev, chunks=4)
    mark("compress_sharded done")
    b = shard.decompress_sharded(decomp, p, n_streams, n, device=dev, chunks=4)
    mark("decompress_sharded done")
if rank in (0, world - 1):
    t0 = marks[0][1]
    print(f"rank {rank}: " + "; ".join(f"{name} +{1e3 * (t - prev):.1f}" for (name, t), (_, prev) in zip(marks[1:], marks[:-1])) +
          f"; total {1e3 * (marks[-1][1] - t0):.1f} ms", flush=True)
dist.destroy_process_group()
