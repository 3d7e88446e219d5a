#!/usr/bin/env python
"""BASELINE.json config 5 shape: a mixed-window batch that starts on rank 0, is scattered over the GPUs of one
box with NCCL point-to-point sends, compressed per shard, and gathered back.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_sharded.py [--mib 256]

Per window class it reports MB/s of (a) kernels only, shards resident and (b) scatter + kernels + gather,
and checks on rank 0 that the gathered bytes equal a single-GPU run of the same batch.
"""
import argparse
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tamp_b200 import batch, shard  # noqa: E402

CLASSES = [(8, 1024), (10, 4096), (12, 16384), (15, 65536)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=256, help="MiB per window class (config 5 uses 4096)")
    ap.add_argument("--extended", type=int, default=1)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    ext = bool(args.extended)
    for w, n in CLASSES:
        n_streams = (args.mib << 20) // n
        stride = (batch.compress_bound(n) + 15) // 16 * 16
        rows = batch.synth(0, 0, n_streams, n, device=dev) if rank == 0 else None

        def fn(x):
            r = batch.compress_batch(x, window=w, extended=ext, out_stride=stride)
            return r.data, r.sizes, r.status

        def timed(f):
            torch.cuda.synchronize()
            dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = f()
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b)], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item()), out

        shard.run_sharded(fn, rows, n_streams, n, device=dev)  # warm-up (NCCL channels, allocations)
        t_all, res = timed(lambda: shard.run_sharded(fn, rows, n_streams, n, device=dev))
        mine = shard.scatter_rows(rows, n_streams, n, device=dev)
        fn(mine)
        t_kernel, _ = timed(lambda: fn(mine))
        if rank == 0:
            out, sizes, status = res
            ref = batch.compress_batch(rows, window=w, extended=ext, out_stride=stride)
            torch.cuda.synchronize()
            ok = bool(torch.equal(ref.sizes, sizes)) and bool((status == 0).all())
            col = torch.arange(stride, device=dev)[None, :] < sizes[:, None]
            ok = ok and bool(torch.equal(ref.data[col], out[col]))
            mb = n_streams * n / 1e6
            print(json.dumps({"window": w, "stream_len": n, "n_streams": n_streams, "n_gpus": world, "extended": int(ext),
                              "kernels_only_MBps": round(mb / t_kernel * 1e3, 1),
                              "scatter_compress_gather_MBps": round(mb / t_all * 1e3, 1),
                              "kernel_ms": round(t_kernel, 2), "end_to_end_ms": round(t_all, 2),
                              "gathered_equals_single_gpu": ok}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
