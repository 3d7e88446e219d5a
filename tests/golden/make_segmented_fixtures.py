#!/usr/bin/env python
"""Golden vectors for the segmented single-stream path (tamp_b200_compress_segmented), generated from the UNMODIFIED
reference C (oracle/_ref/libtamp_ref.so, built from /root/reference by oracle/Makefile):

    one TampCompressor with conf.dictionary_reset = 1:  compress(segment) ; tamp_compressor_reset_dictionary() between the
    segments (compressor.c:847-881) ; flush(write_token = true) at the end.

Inputs come from the synthetic generators (oracle harness), so only sizes and SHA-256 digests are stored.

    python tests/golden/make_segmented_fixtures.py      # writes tests/golden/ref_segmented.json
"""
import hashlib
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402

CASES = [  # (window, literal, extended, segment_size, generator, k, n)
    (10, 8, False, 1024, 0, 7, 300_000),
    (10, 8, False, 4096, 0, 8, 262_144),
    (10, 8, True, 4096, 0, 9, 100_001),
    (10, 8, True, 1024, 5, 10, 50_000),
    (8, 8, False, 1024, 0, 11, 65_536 + 5),
    (9, 7, False, 512, 0, 12, 20_000),
    (12, 8, False, 16384, 0, 13, 400_000),
    (12, 8, True, 16384, 3, 14, 150_000),
    (15, 8, False, 65536, 0, 15, 1_000_000),
    (10, 8, False, 65536, 0, 16, 200_000),
    (10, 8, True, 64, 0, 17, 1_000),
    (10, 8, False, 4096, 0, 18, 5),
    (10, 8, False, 4096, 0, 19, 0),
    (11, 8, False, 2048, 2, 20, 30_000),
]


def ref_segmented(ref, data, seg, *, window, literal, extended):
    c = oracle.RefCompressor(ref, window=window, literal=literal, extended=extended, dictionary_reset=True)
    assert c.init_res == 0
    out, frames = b"", [0]
    for i in range(0, max(len(data), 1), seg):
        s = data[i:i + seg]
        if i:
            o, res = c.reset_dictionary(64)
            assert res == 0
            out += o
            frames.append(len(out) - 2)  # the second FLUSH of the reset (55 80) opens the next segment
        o, consumed, res = c.compress(s, len(s) * 9 // 8 + 64)
        assert res == 0 and consumed == len(s)
        out += o
    o, res = c.flush(64, True)
    assert res == 0
    out += o
    return out, frames + [len(out)]


def main():
    ref = oracle.Ref()
    h = oracle.Harness("port")
    rows = []
    for window, literal, extended, seg, gen, k, n in CASES:
        data = h.generate(gen, k, 1, max(n, 1))[0].tobytes()[:n]
        if literal < 8:
            data = bytes(b & ((1 << literal) - 1) for b in data)
        out, offsets = ref_segmented(ref, data, seg, window=window, literal=literal, extended=extended)
        back, res = ref.decompress(out, window_bits_max=window, cap=n + 64)
        assert back == data, (window, seg, n, res)
        rows.append(dict(window=window, literal=literal, extended=extended, segment_size=seg, gen=gen, k=k, n=n,
                         size=len(out), sha256=hashlib.sha256(out).hexdigest(),
                         offsets_sha256=hashlib.sha256(json.dumps(offsets).encode()).hexdigest(),
                         input_sha256=hashlib.sha256(data).hexdigest()))
    (ROOT / "tests" / "golden" / "ref_segmented.json").write_text(json.dumps(rows, indent=1) + "\n")
    print(f"{len(rows)} segmented streams written")


if __name__ == "__main__":
    main()
