#!/bin/bash
# Round 2, GPU session 38: host-pointer segmented compress, pinned input AND output (the wrapper's fresh pageable output
# buffer dominated session 37's figure)
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 300 python - <<'P' 2>&1 | tee gpurun_out/s38_host_speed.log
import time, torch
from tamp_b200 import batch
n = 1 << 30
x = batch.synth(0, 0, n // 1024, 1024).reshape(-1).cpu().pin_memory()
out = torch.empty(int(n * 1.14) + (1 << 20), dtype=torch.uint8).pin_memory()
for seg in (1024, 4096, 65536):
    batch.compress_segmented(x, seg, window=10, extended=False, out=out)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); s, o = batch.compress_segmented(x, seg, window=10, extended=False, out=out); ts.append(time.perf_counter() - t0)
    back = batch.decompress_segmented(s.cuda(), o.cuda(), seg)
    print(f"host-pointer compress_segmented, pinned 1 GiB in / pinned out, segment {seg}: {n / min(ts) / 1e9:.1f} GB/s ({min(ts) * 1e3:.1f} ms), ratio {s.numel() / n:.4f}, round trip {bool(torch.equal(back.cpu(), x))}")
P
