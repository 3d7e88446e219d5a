#!/bin/bash
# Round 2, GPU session 37: host-pointer segmented compress through the pipelined packed path
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
( timeout 900 python -m pytest tests/test_gpu_segmented.py -q -x 2>&1 | tail -8 ) > gpurun_out/s37_seg_tests.log
tail -4 gpurun_out/s37_seg_tests.log
timeout 300 python - <<'P' 2>&1 | tee gpurun_out/s37_host_speed.log
import time, torch
from tamp_b200 import batch
n = 1 << 30
x = batch.synth(0, 0, n // 1024, 1024).reshape(-1).cpu().pin_memory()
for seg in (1024, 4096, 65536):
    batch.compress_segmented(x, seg, window=10, extended=False)
    t0 = time.perf_counter(); s, o = batch.compress_segmented(x, seg, window=10, extended=False); t1 = time.perf_counter()
    print(f"host-pointer compress_segmented, pinned 1 GiB, segment {seg}: {n / (t1 - t0) / 1e9:.1f} GB/s, ratio {s.numel() / n:.4f}")
P
