#!/usr/bin/env python
"""BASELINE.json config 5 on its own (bench.py carries the same measurement in its `configs["5"]` key at N > 1):

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_sharded.py [--c5-mib 256]

A mixed-window batch that starts and ends on rank 0: NCCL scatter of the input rows, compress per shard, gather-v of
the device-compacted frames, and the way back; NCCL transfers inside the timed region, kernels-only time beside it.
"""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import oracle  # noqa: E402  (checker: parity sample against the reference C)
from tamp_b200 import batch, shard  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c5-mib", type=int, default=256, help="MiB per window class (config 5 uses 4096)")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    peak, _ = bench.measured_peak_gbs()
    res = bench.config5(args, torch, batch, shard, oracle, np, dev, peak, rank, world, dist)
    if rank == 0:
        print(json.dumps(res), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
