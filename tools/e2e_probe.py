#!/usr/bin/env python
"""Host <-> device copy ceiling of the box, the number the end-to-end (host-pointer) bench leg is bounded by.

    python tools/e2e_probe.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/e2e_probe.py   # N ranks at once

Per rank: pinned H2D alone, D2H alone, and both directions at once (two CUDA streams), in 64 MiB chunks like the
pipelined host path of the library; with N ranks all of them copy at the same time (barrier before every leg), so the
aggregate shows what the host's PCIe / memory paths give N GPUs together.  Prints one JSON line (rank 0).
Optionally pins the process to the cores of the GPU's NUMA node first (--pin), the way bench.py does.
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tamp_b200 import hostpin  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=1024)
    ap.add_argument("--chunk-mib", type=int, default=64)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--pin", type=int, default=1)
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    pin = hostpin.pin_to_gpu_node(local) if args.pin else None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.mib << 20
    chunk = args.chunk_mib << 20
    h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_in.fill_(7)
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.full((n,), 3, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def leg(h2d, d2h):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            for off in range(0, n, chunk):
                if h2d:
                    with torch.cuda.stream(s1):
                        d_a[off:off + chunk].copy_(h_in[off:off + chunk], non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s2):
                        h_out[off:off + chunk].copy_(d_b[off:off + chunk], non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.reps
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return n / 1e9 / dt  # GB/s per direction per rank, slowest rank

    leg(True, True)  # warm-up
    res = {"ranks": world, "mib_per_direction": args.mib, "chunk_mib": args.chunk_mib, "pinned_to": pin,
           "h2d_GBps_per_rank": round(leg(True, False), 2), "d2h_GBps_per_rank": round(leg(False, True), 2),
           "both_GBps_per_direction_per_rank": round(leg(True, True), 2)}
    res["both_GBps_per_direction_all_ranks"] = round(res["both_GBps_per_direction_per_rank"] * world, 2)
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
