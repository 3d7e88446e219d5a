#!/usr/bin/env python
"""Differential fuzz of the segment-walk compressor's SOURCE on the CPU SIMT emulator (tests/emu) against the oracle:
random windows / literal widths / generators / lengths, spliced repeats, custom dictionaries.  Test infrastructure."""
import ctypes as C
import random
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle  # noqa: E402
import test_emulated_kernels as T  # noqa: E402
from conftest import gen_stream  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
lib = C.CDLL(str(ROOT / "tests/emu/_build/libemu_kernels.so"))
lib.emu_walk_compress.restype = C.c_int
lib.emu_walk_compress.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64,
                                  C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint, C.c_uint64]
h = oracle.Harness("port")
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 12345)
ext = len(sys.argv) > 3 and sys.argv[3] == "ext"  # k_walk_compress<extended format> with crafted runs / long repeats
t0, total, deferred = time.time(), 0, 0
for rnd in range(rounds):
    window = rng.choice([8, 9, 10, 10])
    W = 1 << window
    lit = rng.choice([8, 8, 8, 7])
    streams = []
    for i in range(30):
        n = rng.choice([W, W, W - 1, rng.randrange(0, W + 1), rng.randrange(0, 64)])
        kind = rng.choice([0, 0, 0, 1, 2, 3, 4, 5]) if lit == 8 else rng.choice([0, 1, 2, 5])
        s = bytearray(gen_stream(h, kind, rng.randrange(1 << 20), n))
        for _ in range(rng.randrange(0, 6)):
            if n > 8:
                a, b = rng.randrange(0, n - 4), rng.randrange(0, n - 4)
                ln = rng.randrange(1, min(20, n - max(a, b)))
                s[b:b + ln] = s[a:a + ln]
        if ext and rng.random() < 0.5:
            s = bytearray(T._crafted(h, rng, max(n, 1), rng.randrange(1 << 16))[:n])
        if lit == 7:
            s = bytearray(b & 127 for b in s)
        streams.append(bytes(s))
    dic = bytes(rng.choice(b"abcde \n tiens") for _ in range(W)) if rng.random() < 0.3 else None
    wt = rng.random() < 0.3
    got = T.ppar(lib, T.WALK_EXT if ext else T.WALK, streams, window=window, literal=lit, dictionary=dic, seed=rnd, grid=rng.choice([1, 2]),
                 max_pairs=rng.choice([8192, 100000]), write_token=wt)
    for s, g in zip(streams, got):
        total += 1
        if g is None:
            deferred += 1
            continue
        want = oracle.compress(s, window=window, literal=lit, extended=ext, dictionary=dic, write_token=wt)
        assert g == (want, 0), (rnd, window, lit, len(s))
print("ok", total, "streams,", deferred, "deferred,", round(time.time() - t0, 1), "s")
