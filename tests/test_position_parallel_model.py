"""CPU check of the claims the position-parallel compressor (tamp_b200/csrc/cuda/ppar_compress.cu) rests on.

A small pure-Python model computes, for EVERY input offset p independently, the best match of input[p...] in
``window_p[x] = input[x] if x < p else dictionary[x]`` (exhaustive search, lowest index on ties) and then derives the
token stream by a walk over that table — greedy for v1, the cached-match rule for lazy matching, run / short-run /
extended-match rules for the extended format (with the "a run longer than 8 bytes makes the window parse-dependent"
bail-out).  The bytes must equal the oracle's, i.e. the reference's serial compressor, for streams no longer than
the window.  No GPU involved: this pins the algorithm, tests/test_gpu_batch.py pins the kernel.
"""
import random

import pytest

import oracle

CODE = [0x00, 0x03, 0x08, 0x0b, 0x14, 0x24, 0x26, 0x2b, 0x4b, 0x54, 0x94, 0x95, 0xaa, 0x27, 0xab]
BITS = [2, 3, 5, 5, 6, 7, 7, 7, 8, 8, 9, 9, 9, 7, 9]


class Bits:
    def __init__(self):
        self.acc, self.n, self.out = 0, 0, bytearray()

    def put(self, v, k):
        self.acc = (self.acc << k) | v
        self.n += k
        while self.n >= 8:
            self.out.append((self.acc >> (self.n - 8)) & 0xFF)
            self.n -= 8
        self.acc &= (1 << self.n) - 1

    def exthuff(self, v, t):  # write_extended_huffman, compressor.c:257-263
        i = v >> t
        self.put((CODE[i] << t) | (v & ((1 << t) - 1)), BITS[i] - 1 + t)

    def finish(self):
        if self.n:
            self.out.append((self.acc << (8 - self.n)) & 0xFF)
        return bytes(self.out)


def match_table(inp, dic, W, maxlen, boundary_shift=0):
    """best[p] = (len, index) of input[p + shift ...] in the window whose first p positions hold input bytes."""
    N = len(inp)
    table = []
    for p in range(N):
        q = p + boundary_shift
        L = min(maxlen, N - q)
        best = (0, 0)
        if L >= 2:
            for x in range(W - 1):
                n = 0
                while n < L and x + n < W and (inp[x + n] if x + n < p else dic[x + n]) == inp[q + n]:
                    n += 1
                if n > best[0]:
                    best = (n, x)
                    if n == L:
                        break
        table.append(best if best[0] >= 2 else (0, 0))
    return table


def model_v1(inp, dic, wbits, lazy):
    W, N = 1 << wbits, len(inp)
    A = match_table(inp, dic, W, 15)
    B = match_table(inp, dic, W, 15, boundary_shift=1) if lazy else None
    w = Bits()
    w.put(((wbits - 8) << 5) | (3 << 3) | 4, 8)  # custom dictionary flag set by the caller of this model
    p, cached = 0, None
    while p < N:
        ln, idx = cached if cached is not None else A[p]
        cached = None
        r = min(16, N - p)
        if lazy and 2 <= ln <= 8 and r > ln + 2:  # compressor.c:576-619
            nl, ni = B[p]
            if nl > ln and not (ni <= p < ni + nl):
                cached, ln = (nl, ni), 0
        if ln < 2:
            w.put(0x100 | inp[p], 9)
            p += 1
        else:
            w.put((CODE[ln - 2] << wbits) | idx, BITS[ln - 2] + wbits)
            p += ln
    return w.finish()


def model_ext(inp, dic, wbits):
    """None = the stream emits a run of more than 8 bytes before its end (left to the serial kernel)."""
    W, N = 1 << wbits, len(inp)
    A = match_table(inp, dic, W, 16)
    w = Bits()
    w.put(((wbits - 8) << 5) | (3 << 3) | 4 | 2, 8)
    p = rle = 0
    while p < N:
        r = min(16, N - p)
        last = inp[p - 1] if p else dic[W - 1]
        if rle or inp[p] == last:  # compressor.c:471-523
            avail = 0
            while avail < 16 and p + avail < N and inp[p + avail] == last:
                avail += 1
            avail = min(avail, r, 241 - rle)
            total = rle + avail
            ended = avail < r or total >= 241
            if not ended and total > 0:
                rle, p = total, p + avail
                continue
            if total >= 2:
                use = not (total == avail and total <= 6 and A[p][0] > total)
                if use:
                    if total > 8 and p + avail < N:
                        return None
                    w.put(CODE[12], BITS[12])
                    w.exthuff(total - 2, 4)
                    p, rle = p + avail, 0
                    continue
            elif rle == 1:
                w.put(0x100 | last, 9)
                rle = 0
                continue
        ln, idx = A[p]
        if ln < 2:
            w.put(0x100 | inp[p], 9)
            p += 1
        elif ln <= 13:
            w.put((CODE[ln - 2] << wbits) | idx, BITS[ln - 2] + wbits)
            p += ln
        else:  # extended match: longest match up to 133 bytes against the window as it is now, lowest index
            if ln == 16:
                cap = min(N - p, 133)
                best = (0, 0)
                for x in range(W - 1):
                    n = 0
                    while n < cap and x + n < W and (inp[x + n] if x + n < p else dic[x + n]) == inp[p + n]:
                        n += 1
                    if n > best[0]:
                        best = (n, x)
                ln, idx = best
            w.put(CODE[13], BITS[13])
            w.exthuff(ln - 14, 3)
            w.put(idx, wbits)
            p += ln
    if rle == 1:
        w.put(0x100 | inp[N - 1], 9)
    elif rle >= 2:
        w.put(CODE[12], BITS[12])
        w.exthuff(rle - 2, 4)
    return w.finish()


def _streams(rng, W):
    words = [bytes(rng.choice(b"abcdefghij") for _ in range(rng.randrange(2, 7))) for _ in range(24)]
    for kind in range(12):
        n = rng.choice([W, W - 1, W // 2, 40, 17, 16, 3, 1])
        if kind % 3 == 0:  # word text
            s = b" ".join(rng.choice(words) for _ in range(n))[:n]
        elif kind % 3 == 1:  # text with runs and repeats (RLE tokens, extended matches)
            s = bytearray(b" ".join(rng.choice(words[:6]) for _ in range(n))[:n])
            for _ in range(4):
                at, ln = rng.randrange(0, max(1, n - 12)), rng.randrange(2, 12)
                s[at:at + ln] = bytes([s[at]]) * len(s[at:at + ln])
            s = bytes(s)
        else:  # few symbols: long chains, overlapping matches
            s = bytes(rng.choice(b"ab\n") for _ in range(n))
        yield s


@pytest.mark.parametrize("wbits,seed", [(8, 2026), (8, 7), (8, 99), (9, 5)])
def test_match_table_plus_walk_reproduces_the_serial_compressor(wbits, seed):
    rng = random.Random(seed)
    W = 1 << wbits
    dic = bytes(rng.choice(b"abcdefghij \n") for _ in range(W))  # custom dictionary: the window's initial content
    deferred = 0
    for s in _streams(rng, W):
        ref = oracle.compress(s, window=wbits, literal=8, extended=False, dictionary=dic)
        assert model_v1(s, dic, wbits, lazy=False) == ref
        ref_lazy = oracle.compress(s, window=wbits, literal=8, extended=False, dictionary=dic, lazy_matching=True)
        assert model_v1(s, dic, wbits, lazy=True) == ref_lazy
        ext = model_ext(s, dic, wbits)
        if ext is None:
            deferred += 1
        else:
            assert ext == oracle.compress(s, window=wbits, literal=8, extended=True, dictionary=dic)
    assert deferred < 12
