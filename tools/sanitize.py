#!/usr/bin/env python
"""Small workload touching every kernel once; meant to run under compute-sanitizer
(memcheck / racecheck / synccheck):  compute-sanitizer --tool racecheck python tools/sanitize.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tamp_b200 import batch  # noqa: E402
from tamp_b200.capi import CCompressor, CDecompressor  # noqa: E402

for mode in (0, 1, 2, 4, 6):  # 6: the long split decompressor for every batch
    batch.set_kernel_mode(mode)
    for w, n, ns in [(8, 700, 96), (10, 1024, 128), (10, 3000, 64), (12, 5000, 16), (13, 3000, 8), (15, 9000, 4)]:
        for gen in (0, 5):
            x = batch.synth(gen, 3, ns, (n + 15) // 16 * 16)
            for ext in (False, True):
                r = batch.compress_batch(x, window=w, extended=ext)
                d = batch.decompress_batch(r.data, r.sizes, x.shape[1], window_bits_max=w)
                torch.cuda.synchronize()
                assert torch.equal(d.data, x), (mode, w, n, gen, ext)
# lazy matching through the position-parallel kernel (TAMP_LAZY_MATCHING flavour of the library)
batch.set_kernel_mode(0)
for gen in (0, 5):
    x = batch.synth(gen, 7, 64, 1024)
    r = batch.compress_batch(x, window=10, extended=False, lazy_matching=True)
    d = batch.decompress_batch(r.data, r.sizes, 1024, window_bits_max=10)
    torch.cuda.synchronize()
    assert torch.equal(d.data, x), ("lazy", gen)
# output compaction and the packed-frame input layout
x = batch.synth(0, 9, 3000, 512)
r = batch.compress_batch(x, window=9, extended=True)
packed, offsets = batch.compact(r)
d = batch.decompress_packed(packed, offsets[:-1], r.sizes, 512 + 16, window_bits_max=9)
torch.cuda.synchronize()
assert torch.equal(d.data[:, :512], x)
# one long stream as a batch of dictionary_reset segments (append-mode frames, segment headers in the decompressors), with
# a last segment that is not a multiple of 16 bytes and an input that ends at the last byte of its allocation
for mode in (0, 1, 2, 6):
    batch.set_kernel_mode(mode)
    for w, seg, n, ext in [(10, 1024, 50_003, False), (10, 4096, 70_001, True), (12, 8192, 100_007, False), (8, 512, 9_999, True)]:
        data = batch.synth(0, 11, (n + 1023) // 1024, 1024).reshape(-1)[:n].clone()
        stream, offs = batch.compress_segmented(data, seg, window=w, extended=ext)
        back = batch.decompress_segmented(stream, offs, seg, out_size=n)
        torch.cuda.synchronize()
        assert torch.equal(back, data), ("segmented", mode, w, seg, ext)
batch.set_kernel_mode(0)
x = batch.synth(0, 13, 64, 1024)
r = batch.compress_batch(x, window=10, extended=False, dictionary_reset=True, append=True, write_token=True)
torch.cuda.synchronize()
assert bool((r.status == 0).all())
c = CCompressor(window=10)
out, _, res = c.compress_and_flush(b"hello hello hello world" * 20, 1000, True)
assert res == 0
back, _, res = CDecompressor(window_bits=10).decompress(out, 1000)
assert back == b"hello hello hello world" * 20
print("sanitize workload ok")
