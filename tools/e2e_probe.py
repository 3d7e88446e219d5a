#!/usr/bin/env python
"""Where the host-pointer (e2e) step of bench.py spends its time: raw pinned H2D / D2H bandwidth of the box, then the
compress call and the decompress call of the bench workload timed separately."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tamp_b200 import batch  # noqa: E402

n_streams, n = 1 << 20, 1024
x = batch.synth(0, 0, n_streams, n)
hx = torch.empty((n_streams, n), dtype=torch.uint8, pin_memory=True)
hx.copy_(x)
dx = torch.empty_like(x)
for name, fn in (("H2D", lambda: dx.copy_(hx, non_blocking=True)), ("D2H", lambda: hx.copy_(dx, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"{name}: {n_streams * n / 1e9 / dt:.1f} GB/s ({dt * 1e3:.1f} ms per GiB)")
hx.copy_(x)
stride = (batch.compress_bound(n, 8) + 15) // 16 * 16
hcomp = torch.empty((n_streams, stride), dtype=torch.uint8, pin_memory=True)
hback = torch.empty((n_streams, n), dtype=torch.uint8, pin_memory=True)
for it in range(4):
    t0 = time.perf_counter()
    r = batch.compress_batch(hx, window=10, literal=8, extended=False, out=hcomp)
    t1 = time.perf_counter()
    d = batch.decompress_batch(hcomp, r.sizes, n, window_bits_max=10, out=hback)
    t2 = time.perf_counter()
    print(f"iter {it}: compress call {1e3 * (t1 - t0):.1f} ms, decompress call {1e3 * (t2 - t1):.1f} ms")
assert torch.equal(hback, hx)
