#!/usr/bin/env python
"""ONE long stream, segment-parallel (tamp_b200_compress_segmented / _decompress_segmented).

    python tools/bench_segmented.py [--mib 1024] [--cpu-mib 8]

Times the device-resident calls (CUDA events around the whole call: kernels, compaction, the status read-back) on G_text
for a few (window, segment size) pairs and prints one JSON line each: uncompressed GB/s both ways, the ratio, what the
unsegmented stream's ratio would be (reference C on a sample), the reference C on ONE host core (a single stream cannot
use more), and the parity of a sample against the reference.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import oracle  # noqa: E402  (checker / CPU baseline only)
from tamp_b200 import batch  # noqa: E402

CASES = [(10, 1024), (10, 4096), (10, 65536), (12, 65536), (15, 65536), (15, 1 << 20)]


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
    ap.add_argument("--mib", type=int, default=1024)
    ap.add_argument("--cpu-mib", type=int, default=4)
    ap.add_argument("--extended", type=int, default=0)
    args = ap.parse_args()
    n = args.mib << 20
    data = batch.synth(0, 0, n // 1024, 1024).reshape(-1)  # 1 KiB G_text pieces back to back: one long text
    ref = oracle.Ref() if oracle.ref_available() else None
    sample = data[:args.cpu_mib << 20].cpu().numpy().tobytes()
    for window, seg in CASES:
        kw = dict(window=window, extended=bool(args.extended))
        t_c, (stream, offs) = timed(lambda: batch.compress_segmented(data, seg, **kw))
        t_d, out = timed(lambda: batch.decompress_segmented(stream, offs, seg, window_bits_max=window))
        ok = bool(torch.equal(out, data))
        line = {"window": window, "segment_size": seg, "extended": args.extended, "bytes": n, "segments": offs.numel() - 1,
                "ratio": round(stream.numel() / n, 4), "compress_ms": round(t_c, 2), "compress_GBps": round(n / t_c / 1e6, 2),
                "decompress_ms": round(t_d, 2), "decompress_GBps": round(n / t_d / 1e6, 2), "round_trip_ok": ok}
        if ref is not None:
            # parity sample: the same stream cut at the sample's end is what one reference compressor writes for the sample
            k = len(sample) // seg
            want = b""
            c = oracle.RefCompressor(ref, dictionary_reset=True, **kw)
            t0 = time.perf_counter()
            for i in range(k):
                if i:
                    want += c.reset_dictionary(64)[0]
                want += c.compress(sample[i * seg:(i + 1) * seg], seg * 9 // 8 + 64)[0]
            want += c.flush(64, True)[0]
            dt = time.perf_counter() - t0
            got = stream[:int(offs[k])].cpu().numpy().tobytes()
            line["parity"] = "bit-exact vs one reference compressor on the first %d segments" % k if got == want else "MISMATCH"
            line["cpu_one_core_compress_MBps"] = round(k * seg / dt / 1e6, 2)
            t0 = time.perf_counter()
            back, res = ref.decompress(want, window_bits_max=window, cap=k * seg + 64)
            line["cpu_one_core_decompress_MBps"] = round(k * seg / (time.perf_counter() - t0) / 1e6, 2)
            assert back == sample[:k * seg]
            whole = ref.compress(sample[:k * seg], window=window, extended=bool(args.extended))
            line["ratio_unsegmented_sample"] = round(len(whole) / (k * seg), 4)
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
