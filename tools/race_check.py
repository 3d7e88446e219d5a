#!/usr/bin/env python
"""Small workload for `compute-sanitizer --tool racecheck`: the position-parallel compressor in its three modes
(v1, lazy, extended), the pick-up pass (run-heavy generator), the decompressor and the compaction kernels.
Racecheck is slow; tools/sanitize.py is the (larger) memcheck workload."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tamp_b200 import batch  # noqa: E402

ONLY_NEW = "--segmented-only" in sys.argv  # (just the workloads added with the segmented calls)
for w, n in [] if ONLY_NEW else [(10, 1024)]:
    for gen in (0, 5, 3):
        x = batch.synth(gen, 3, 48, n)
        for ext in (False, True):
            r = batch.compress_batch(x, window=w, extended=ext)
            d = batch.decompress_batch(r.data, r.sizes, n + 16, window_bits_max=w)
            torch.cuda.synchronize()
            assert torch.equal(d.data[:, :n], x)
        r = batch.compress_batch(x, window=w, extended=False, lazy_matching=True)
        torch.cuda.synchronize()
if not ONLY_NEW:  # lap variants, wide decompressor (default dispatch since round 2)
    batch.set_kernel_mode(0)
    for w, n, ext, lazy in [(10, 4096, False, False), (8, 1024, False, True), (10, 1024, True, False), (13, 6000, True, False),
                            (15, 9000, False, False)]:
        x = batch.synth(0, 5, 24, n)
        kw = {"lazy_matching": True} if lazy else {}
        r = batch.compress_batch(x, window=w, extended=ext, **kw)
        d = batch.decompress_batch(r.data, r.sizes, n + 16, window_bits_max=w)
        torch.cuda.synchronize()
        assert torch.equal(d.data[:, :n], x), (w, n, ext, lazy)
    batch.set_kernel_mode(0)
# history walks (lane per segment / warp per walker) on streams of several chunks, the long split decompressor (mode 6)
for mode in () if ONLY_NEW else (0, 6):
    batch.set_kernel_mode(mode)
    for w, n in [(8, 3000), (10, 5000), (11, 5000), (12, 9000), (15, 20000)]:
        for gen in (0, 3):
            x = batch.synth(gen, 11, 12, n // 16 * 16)
            r = batch.compress_batch(x, window=w, extended=False)
            d = batch.decompress_batch(r.data, r.sizes, x.shape[1] + 16, window_bits_max=w)
            torch.cuda.synchronize()
            assert torch.equal(d.data[:, :x.shape[1]], x), (mode, w, n, gen)
batch.set_kernel_mode(0)
# segments of one stream: append-mode frames out of the walk kernels, segment headers into the split decompressors
for w, seg, n in [(10, 1024, 20_011), (10, 4096, 30_001), (12, 8192, 40_003)]:
    data = batch.synth(0, 17, (n + 1023) // 1024, 1024).reshape(-1)[:n].clone()
    stream, offs = batch.compress_segmented(data, seg, window=w, extended=False)
    back = batch.decompress_segmented(stream, offs, seg, out_size=n)
    torch.cuda.synchronize()
    assert torch.equal(back, data), ("segmented", w, seg)
# the split decompressor proper (rows no longer than the window: its copy phase keeps the row in shared memory), both formats
for ext in (False, True):
    x = batch.synth(0, 21, 96, 1024)
    r = batch.compress_batch(x, window=10, extended=ext)
    d = batch.decompress_batch(r.data, r.sizes, 1024, window_bits_max=10)
    torch.cuda.synchronize()
    assert torch.equal(d.data, x), ("split", ext)
x = batch.synth(0, 9, 3000, 512)
r = batch.compress_batch(x, window=9, extended=True)
packed, offsets = batch.compact(r)
torch.cuda.synchronize()
print("race workload ok")
