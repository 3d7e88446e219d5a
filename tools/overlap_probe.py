#!/usr/bin/env python
"""Do a host-pointer compress call and a host-pointer decompress call overlap when issued from two threads?
Times each alone and both together on the bench workload (2^20 x 1 KiB, window 10, v1)."""
import sys
import threading
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tamp_b200 import batch  # noqa: E402

n, ns = 1024, 1 << 20
x = batch.synth(0, 0, ns, n)
hx = torch.empty((ns, n), dtype=torch.uint8, pin_memory=True)
hx.copy_(x)
r0 = batch.compress_batch(hx, window=10, extended=False)
hcomp = [torch.empty_like(r0.data).pin_memory() for _ in range(2)]
hback = torch.empty((ns, n), dtype=torch.uint8, pin_memory=True)
hcomp[0].copy_(r0.data)
sizes = r0.sizes


def comp():
    batch.compress_batch(hx, window=10, extended=False, out=hcomp[1])


def dec():
    batch.decompress_batch(hcomp[0], sizes, n, window_bits_max=10, out=hback)


def timed(fns, reps=4):
    best = 1e9
    for _ in range(reps):
        ths = [threading.Thread(target=f) for f in fns]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


comp(); dec()
print("compress alone ms", round(timed([comp]), 2))
print("decompress alone ms", round(timed([dec]), 2))
print("both together ms", round(timed([comp, dec]), 2))
print("two compress calls together ms (same direction: serialised by design)", round(timed([comp, comp]), 2))
assert torch.equal(hback, hx)

# the same through the packed entry points (contiguous frames + offsets)
hpacked = [torch.empty(int(sizes.sum().item() * 1.02) + (1 << 20), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
hoffs = [torch.empty(ns + 1, dtype=torch.int64) for _ in range(2)]
res = batch.compress_batch_packed(hx, window=10, extended=False, packed=hpacked[0], offsets=hoffs[0])


def compp():
    batch.compress_batch_packed(hx, window=10, extended=False, packed=hpacked[1], offsets=hoffs[1])


def decp():
    batch.decompress_packed(hpacked[0], hoffs[0], res[2], n, window_bits_max=10, out=hback)


compp(); decp()
print("packed: compress alone ms", round(timed([compp]), 2))
print("packed: decompress alone ms", round(timed([decp]), 2))
print("packed: both together ms", round(timed([compp, decp]), 2))
assert torch.equal(hback, hx)
