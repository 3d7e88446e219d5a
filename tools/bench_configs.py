#!/usr/bin/env python
"""Secondary measurements: BASELINE.json configs 3-5 shapes (not the headline bench line).

    python tools/bench_configs.py [--mib 256] [--mode 0|1]

For each (window, stream length) class it times compress and decompress kernels (CUDA events, data resident
in HBM, 3 runs after 1 warm-up) and checks the round trip.  Prints one JSON line per class.
"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tamp_b200 import batch  # noqa: E402

CLASSES = [(8, 1024), (10, 1024), (10, 4096), (12, 16384), (15, 65536)]


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=256)
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--gen", type=int, default=0)
    ap.add_argument("--classes", default="")
    ap.add_argument("--v1-only", action="store_true")
    args = ap.parse_args()
    batch.set_kernel_mode(args.mode)
    classes = CLASSES
    if args.classes:
        classes = [tuple(int(v) for v in c.split(":")) for c in args.classes.split(",")]
    for w, n in classes:
        n_streams = max(1, (args.mib << 20) // n)
        x = batch.synth(args.gen, 0, n_streams, n)
        for ext in ((False,) if args.v1_only else (False, True)):
            t_c, r = timed(lambda: batch.compress_batch(x, window=w, extended=ext))
            t_d, d = timed(lambda: batch.decompress_batch(r.data, r.sizes, n, window_bits_max=w))
            ok = bool(torch.equal(d.data, x)) and bool((r.status == 0).all())
            mb = n_streams * n / 1e6
            print(json.dumps({"window": w, "stream_len": n, "n_streams": n_streams, "extended": int(ext),
                              "ratio": round(r.sizes.double().sum().item() / (n_streams * n), 4),
                              "compress_ms": round(t_c, 3), "compress_GBps": round(mb / t_c, 2),
                              "decompress_ms": round(t_d, 3), "decompress_GBps": round(mb / t_d, 2),
                              "round_trip_ok": ok, "kernel_mode": args.mode}), flush=True)


if __name__ == "__main__":
    main()
