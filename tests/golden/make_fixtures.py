#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs oracle/_ref/libtamp_ref*.so, i.e. /root/reference):

    python -m oracle  # or: make -C oracle all
    PYTHONPATH=. python tests/golden/make_fixtures.py

Outputs (committed; they travel to the GPU box where /root/reference does not exist):

* reference_kats.json    — known-answer vectors TRANSCRIBED from the reference's own tests (citations in
                           each entry); this script re-checks every one against libtamp_ref before writing.
* ref_fixtures.json      — (generator kind, stream index, length, conf) -> size + SHA-256 of the bytes the
                           reference C produces; inputs are regenerated from the seeded generators
                           (SURVEY.md 8d), so only digests are stored.
* ref_api_sequences.json — recorded call-by-call traces of the reference C API (sink/poll/flush/compress/
                           decompress with small output buffers), used to pin the per-call C-ABI drop-in.
"""
from __future__ import annotations

import hashlib
import json
import random
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))

import oracle  # noqa: E402
from oracle import Harness, Ref, RefCompressor, RefDecompressor, pack_conf  # noqa: E402

OUT = Path(__file__).resolve().parent

SEED_256 = (b"\x00.//r.0. t>\n/>snas.trnr i\x00r/a\x00snat./.r\x00i o.s tneo>.as>\na.ta\x00 aa\x00\x00\x000oe ri\x00a>eatsi\n.\ni."
            b"str\n//snesr.ost<  \x00\ni\neoa\x00se0.o\n\n>aori>n0.>./.oonen0<\x00<r o\n\naas0< ai\n0\x00na\x00e><.\noas to \n></se"
            b">>ts/oreatinter.n0 >s\n/.e.><. r si<>/<san\x00ae t 0.r.o/0./a r/ttn nn.<re.t0 \x00r\x00ro")


def dict_with(size, fill, *patches):
    d = bytearray([fill]) * size
    for off, data in patches:
        d[off:off + len(data)] = data
    return bytes(d)


def kats():
    z256 = bytes(256)
    K = []

    def enc(name, cite, data, out_hex, **conf):
        K.append(dict(kind="compress", name=name, cite=cite, input=data.hex(), expected=out_hex.replace(" ", "").lower(),
                      conf={k: (v.hex() if isinstance(v, (bytes, bytearray)) else v) for k, v in conf.items()}))

    def dec(name, cite, comp_hex, out, status, dictionary=None, window_bits_max=15):
        K.append(dict(kind="decompress", name=name, cite=cite, input=comp_hex.replace(" ", "").lower(),
                      expected=out.hex(), status=status, dictionary=dictionary.hex() if dictionary else None,
                      window_bits_max=window_bits_max))

    enc("foo_v1", "tests/test_compressor.py:66-110", b"foo foo foo", "58 B3 04 1C 81 00 03 00 00",
        window=10, literal=8, extended=False)
    enc("foo_7bit", "tests/test_compressor.py:145-174", b"foo foo foo", "50 E6 08 3A 04 00 0C 00",
        window=10, literal=7, extended=False)
    enc("custom_dict_11", "tests/test_compressor.py:176-203", b"foo foo foo", "14 54 00",
        window=8, literal=7, extended=False, dictionary=b"foo foo foo" + bytes(256 - 11))
    enc("oob_2_byte", "tests/test_compressor.py:212-237", b"Q\x00Q", "58 A8 C0 2A 20", window=10, literal=8,
        extended=False)
    enc("rle_A20", "tests/test_compressor.py:313-337", b"A" * 20, "5A A0 AA B1", window=10, literal=8, extended=True)
    enc("rle_B5", "tests/test_compressor.py:339-361", b"B" * 5, "5A A1 2A 84", window=10, literal=8, extended=True)
    enc("ext_match_14", "tests/test_compressor.py:363-395", b"abcdefghijklmn", "1E 4E 00 00", window=8, literal=8,
        extended=True, dictionary=b"abcdefghijklmn" + bytes(256 - 14))
    enc("ext_match_16", "tests/test_compressor.py:397-425", b"abcdefghijklmnop", "1E 4E 40 00", window=8, literal=8,
        extended=True, dictionary=b"abcdefghijklmnop" + bytes(256 - 16))
    enc("finder_window_edge", "ctests/test_compressor.c:802-811", b"WXYZ!!", "1C 47 E4 86 42", window=8, literal=8,
        extended=False, dictionary=dict_with(256, ord("a"), (250, b"UVWXYZ")))
    enc("finder_alignment_phases", "ctests/test_compressor.c:813-824", b"Qabcd", "1C 59 40", window=8, literal=8,
        extended=False, dictionary=dict_with(256, 0xFF, (3, b"Qa"), (13, b"Qab"), (26, b"Qabc"), (40, b"Qabcd")))
    enc("finder_swar_bytes", "ctests/test_compressor.c:826-838", b"\x00\x80\x7f\x01\xff", "1C 5E 40", window=8,
        literal=8, extended=False,
        dictionary=dict_with(256, 0x01, (10, b"\x00\x7f"), (50, b"\x80\x00\x80"), (100, b"\x00\x00\x00"),
                             (200, b"\x00\x80\x7f\x01\xff")))
    enc("finder_max_pattern_early_exit", "ctests/test_compressor.c:840-848", b"ABCDEFGHIJKLMNOP", "1C 4E 3D 50",
        window=8, literal=8, extended=False, dictionary=dict_with(256, ord("z"), (30, b"ABCDEFGHIJKLMNOP")))

    dec("foo_v1", "ctests/test_decompressor.c:14-26", "58 B3 04 1C 81 00 03 00 00", b"foo foo foo", oracle.INPUT_EXHAUSTED)
    dec("malicious_oob", "ctests/test_decompressor.c:79-97", "58 3F F0", b"", oracle.OOB, window_bits_max=10)
    dec("rle_A20", "ctests/test_decompressor.c:105-110", "5A A0 AA B1", b"A" * 20, oracle.INPUT_EXHAUSTED)
    dec("ext_match_14", "ctests/test_decompressor.c:139-144", "1E 4E 00 00", b"abcdefghijklmn", oracle.INPUT_EXHAUSTED,
        dictionary=b"abcdefghijklmn" + bytes(256 - 14))
    dec("flushing", "tests/test_decompressor.py:99-113", "58 A8 AA C0 AB AA C0", b"QW", oracle.INPUT_EXHAUSTED)
    dec("overlap_snapshot", "tests/test_decompressor.py:124-158", "5C B0 B0 00", b"aabc", oracle.INPUT_EXHAUSTED,
        dictionary=b"abcd" + bytes(1020))
    K.append(dict(kind="dictionary", name="seed_256", cite="tests/test_pseudorandom.py:21-23", literal=8,
                  expected=SEED_256.hex()))
    del z256
    return K


def check_kats(K, ref: Ref):
    for k in K:
        if k["kind"] == "compress":
            conf = dict(k["conf"])
            if "dictionary" in conf:
                conf["dictionary"] = bytes.fromhex(conf["dictionary"])
            got = ref.compress(bytes.fromhex(k["input"]), **conf)
            assert got.hex() == k["expected"], (k["name"], got.hex())
        elif k["kind"] == "decompress":
            d = bytes.fromhex(k["dictionary"]) if k["dictionary"] else None
            got, st = ref.decompress(bytes.fromhex(k["input"]), dictionary=d, window_bits_max=k["window_bits_max"])
            assert got.hex() == k["expected"] and st == k["status"], (k["name"], got, st)
        else:
            buf = (oracle.C.c_char * 256)()
            ref.L.tamp_initialize_dictionary(buf, 256, k["literal"])
            assert bytes(buf).hex() == k["expected"], k["name"]


def grid_fixtures(ref: Ref, refl: Ref, h: Harness):
    F = []
    rng = random.Random(20261017)

    def add(kind, k, n, **conf):
        data = h.generate(kind, k, 1, max(n, 1))[0].tobytes()[:n]
        lit = conf.get("literal", 8)
        if lit < 8:
            data = bytes(b & ((1 << lit) - 1) for b in data)
        r = (refl if conf.get("lazy_matching") else ref).compress(data, **conf)
        dec, st = ref.decompress(r, cap=n + 16)
        assert dec == data and st == oracle.INPUT_EXHAUSTED
        _, st_exact = ref.decompress(r, cap=n)
        F.append(dict(gen=kind, k=k, n=n, conf=conf, in_sha=hashlib.sha256(data).hexdigest()[:16],
                      size=len(r), sha=hashlib.sha256(r).hexdigest(), status_exact_cap=st_exact))

    # the SURVEY 8(c) known answers + BASELINE configs' shapes
    for (w, n, k, ext) in [(10, 4096, 0, 0), (10, 4096, 0, 1), (10, 1024, 0, 0), (10, 1024, 0, 1), (10, 1024, 12345, 1),
                           (8, 4096, 0, 1), (12, 4096, 0, 1), (15, 65536, 0, 0), (15, 65536, 0, 1), (8, 1024, 7, 0),
                           (12, 16384, 3, 1), (12, 16384, 3, 0)]:
        add(oracle.TEXT, k, n, window=w, literal=8, extended=bool(ext))
    # every generator x window x format, ragged lengths incl. 0/1/15/16/17
    for gen in range(6):
        for w in (8, 9, 10, 11, 12, 13, 14, 15):
            for ext in (False, True):
                for n in (0, 1, 2, 15, 16, 17, 33, 257, 1024, 1500, 4099):
                    if w >= 13 and n not in (0, 17, 1024, 4099):
                        continue
                    add(gen, rng.randrange(1 << 20), n, window=w, literal=8, extended=ext)
    # literal widths (7-bit clean generators only), dictionary_reset header, flush token, lazy matching
    for lit in (5, 6, 7):
        for w in (8, 10, 11, 12, 15):
            for ext in (False, True):
                add(oracle.TEXT, rng.randrange(1 << 20), 2000, window=w, literal=lit, extended=ext)
    for w in (8, 10, 12):
        for ext in (False, True):
            add(oracle.TEXT, rng.randrange(1 << 20), 3000, window=w, literal=8, extended=ext, dictionary_reset=True)
            add(oracle.RUNS, rng.randrange(1 << 20), 3000, window=w, literal=8, extended=ext, write_token=True)
            add(oracle.TEXT, rng.randrange(1 << 20), 3000, window=w, literal=8, extended=ext, lazy_matching=True)
            add(oracle.RUNS, rng.randrange(1 << 20), 3000, window=w, literal=8, extended=ext, lazy_matching=True)
    return F


def api_sequences(ref: Ref, h: Harness):
    """Record call-by-call traces of the reference C API."""
    S = []
    rng = random.Random(4242)
    for case in range(60):
        w = rng.choice([8, 10, 12])
        ext = rng.random() < 0.6
        dr = rng.random() < 0.25
        gen = rng.choice([oracle.TEXT, oracle.RUNS, oracle.PERIODIC, oracle.ALPHA16])
        n = rng.choice([40, 200, 900, 2500])
        k = rng.randrange(1 << 20)
        data = h.generate(gen, k, 1, n)[0].tobytes()
        c = RefCompressor(ref, window=w, extended=ext, dictionary_reset=dr)
        ops = []
        pos = 0
        style = rng.choice(["sinkpoll", "compress", "mixed"])
        while pos < n:
            if style == "sinkpoll" or (style == "mixed" and rng.random() < 0.5):
                m = rng.choice([1, 3, 7, 16, 20])
                chunk = data[pos:pos + m]
                took = c.sink(chunk)
                pos += took
                ops.append(dict(op="sink", n=len(chunk), consumed=took))
                if c.full() or rng.random() < 0.15:
                    cap = rng.choice([0, 1, 2, 3, 8, 8, 8])
                    out, res = c.poll(cap)
                    ops.append(dict(op="poll", cap=cap, out=out.hex(), res=res))
            else:
                m = rng.choice([5, 16, 64, 300])
                cap = rng.choice([1, 4, 16, 1000])
                chunk = data[pos:pos + m]
                out, took, res = c.compress(chunk, cap)
                pos += took
                ops.append(dict(op="compress", n=len(chunk), cap=cap, out=out.hex(), consumed=took, res=res))
            if rng.random() < 0.04:
                cap = rng.choice([2, 5, 64])
                out, res = c.flush(cap, True)
                ops.append(dict(op="flush", cap=cap, write_token=True, out=out.hex(), res=res))
        for _ in range(200):
            cap = rng.choice([1, 3, 6, 64])
            out, res = c.flush(cap, False)
            ops.append(dict(op="flush", cap=cap, write_token=False, out=out.hex(), res=res))
            if res == oracle.OK:
                break
        S.append(dict(kind="compress", window=w, extended=ext, dictionary_reset=dr, gen=gen, k=k, n=n, ops=ops,
                      final_state=c.state.raw[8:].hex(), final_window_sha=hashlib.sha256(c.window.raw).hexdigest()))
    # decompress traces: chunked input, small output buffers
    for case in range(60):
        w = rng.choice([8, 10, 12])
        ext = rng.random() < 0.6
        gen = rng.choice([oracle.TEXT, oracle.RUNS, oracle.PERIODIC])
        n = rng.choice([40, 300, 2000])
        k = rng.randrange(1 << 20)
        data = h.generate(gen, k, 1, n)[0].tobytes()
        comp = ref.compress(data, window=w, extended=ext)
        if case % 10 == 9:  # corrupt one byte: status codes must agree too
            comp = bytearray(comp)
            comp[rng.randrange(1, len(comp))] ^= 1 << rng.randrange(8)
            comp = bytes(comp)
        d = RefDecompressor(ref, window_bits=w)
        ops = []
        pos = 0
        guard = 0
        while guard < 20000:
            guard += 1
            m = rng.choice([0, 1, 2, 5, 33, 500]) if pos < len(comp) else 0
            cap = rng.choice([0, 1, 2, 7, 50, 300])
            chunk = comp[pos:pos + m]
            out, took, res = d.decompress(chunk, cap)
            pos += took
            ops.append(dict(n=len(chunk), cap=cap, out=out.hex(), consumed=took, res=res))
            if res < 0 or (res == oracle.INPUT_EXHAUSTED and pos >= len(comp) and m == 0):
                break
        S.append(dict(kind="decompress", window=w, comp=comp.hex(), ops=ops, final_state=d.state.raw[8:].hex(),
                      final_window_sha=hashlib.sha256(d.window.raw).hexdigest()))
    return S


def main():
    oracle.build(ref=True)
    ref, refl, h = Ref(), Ref(lazy=True), Harness("reference")
    K = kats()
    check_kats(K, ref)
    (OUT / "reference_kats.json").write_text(json.dumps(K, indent=1))
    F = grid_fixtures(ref, refl, h)
    (OUT / "ref_fixtures.json").write_text(json.dumps(F, indent=0))
    S = api_sequences(ref, h)
    (OUT / "ref_api_sequences.json").write_text(json.dumps(S, indent=0))
    print(f"kats={len(K)} fixtures={len(F)} sequences={len(S)}")


if __name__ == "__main__":
    main()
