"""The oracle (oracle/tamp_oracle.c) pinned against the reference: its own tests' golden vectors,
fixtures generated from the unmodified reference C (tests/golden/make_fixtures.py), and — when
oracle/_ref is present — the reference itself, live."""
import hashlib
import random

import pytest

import oracle
from conftest import gen_stream


def _conf(k):
    conf = dict(k["conf"])
    if "dictionary" in conf:
        conf["dictionary"] = bytes.fromhex(conf["dictionary"])
    return conf


def test_reference_kats(kats):
    """Every golden bitstream the reference's tests hold for this path."""
    seen = 0
    for k in kats:
        if k["kind"] == "compress":
            got = oracle.compress(bytes.fromhex(k["input"]), **_conf(k))
            assert got.hex() == k["expected"], k["name"]
            conf = _conf(k)
            back, st = oracle.decompress(got, dictionary=conf.get("dictionary"))
            assert back == bytes.fromhex(k["input"]) and st == oracle.INPUT_EXHAUSTED, k["name"]
        elif k["kind"] == "decompress":
            d = bytes.fromhex(k["dictionary"]) if k["dictionary"] else None
            got, st = oracle.decompress(bytes.fromhex(k["input"]), dictionary=d, window_bits_max=k["window_bits_max"])
            assert (got.hex(), st) == (k["expected"], k["status"]), k["name"]
        else:
            assert oracle.initialize_dictionary(256, k["literal"]).hex() == k["expected"]
        seen += 1
    assert seen >= 19


def test_min_pattern_size():
    # common.c:54-56
    for w in range(8, 16):
        for lit in range(5, 9):
            assert oracle.min_pattern_size(w, lit) == 2 + (w > 10 + 2 * (lit - 5))


def test_ref_fixtures(ref_fixtures, harness):
    """Digests of reference-C output on the seeded generators; inputs are regenerated here."""
    for f in ref_fixtures:
        conf = dict(f["conf"])
        data = gen_stream(harness, f["gen"], f["k"], f["n"], conf.get("literal", 8))
        assert hashlib.sha256(data).hexdigest()[:16] == f["in_sha"], "generator drift"
        got = oracle.compress(data, **conf)
        assert len(got) == f["size"] and hashlib.sha256(got).hexdigest() == f["sha"], f
        back, st = oracle.decompress(got, cap=f["n"] + 16)
        assert back == data and st == oracle.INPUT_EXHAUSTED
        _, st = oracle.decompress(got, cap=f["n"])
        assert st == f["status_exact_cap"]


def test_find_best_match_spec():
    # fuzz/esp32_host/differential.cpp:50-67 semantics on the ctests' edge dictionaries
    d = bytearray(b"a" * 256)
    d[250:256] = b"UVWXYZ"
    assert oracle.find_best_match(bytes(d), b"WXYZ!!") == (252, 4)
    d = bytearray(b"\xff" * 256)
    d[3:5], d[13:16], d[26:30], d[40:45] = b"Qa", b"Qab", b"Qabc", b"Qabcd"
    assert oracle.find_best_match(bytes(d), b"Qabcd") == (40, 5)
    assert oracle.find_best_match(bytes(d), b"Qa") == (3, 2)       # ties -> lowest index
    assert oracle.find_best_match(b"ab" * 128, b"b") == (0, 0)     # needs >= 2 bytes
    assert oracle.find_best_match(b"xy" + b"ab" * 127, b"ba", 2) == (3, 2)


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built (no /root/reference here)")
def test_live_differential_vs_reference(harness):
    ref, refl = oracle.Ref(), oracle.Ref(lazy=True)
    rng = random.Random(99)
    for it in range(400):
        w = rng.choice([8, 9, 10, 11, 12, 15])
        lit = rng.choice([5, 6, 7, 8, 8])
        conf = dict(window=w, literal=lit, extended=rng.random() < 0.6, dictionary_reset=rng.random() < 0.2,
                    lazy_matching=rng.random() < 0.3, write_token=rng.random() < 0.3)
        n = rng.choice([0, 1, 16, 17, 100, 1024, 2600])
        data = gen_stream(harness, rng.randrange(6), rng.randrange(1 << 30), n, lit)
        if rng.random() < 0.2:
            conf["dictionary"] = bytes(rng.choice(data) if data and rng.random() < .7 else rng.randrange(1 << lit)
                                       for _ in range(1 << w))
        a = oracle.compress(data, **conf)
        b = (refl if conf["lazy_matching"] else ref).compress(data, **conf)
        assert a == b, (it, conf.keys(), w, lit, n)
        dic = conf.get("dictionary")
        cap = rng.choice([None, len(data), max(0, len(data) // 2)])
        assert oracle.decompress(a, dictionary=dic, cap=cap) == ref.decompress(a, dictionary=dic, cap=cap)


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built (no /root/reference here)")
def test_streaming_encoder_vs_reference(harness):
    """write()/flush(token) call sequences: bytes and final window equal the reference's."""
    ref = oracle.Ref()
    rng = random.Random(7)
    for it in range(150):
        w, ext, dr = rng.choice([8, 10, 12]), rng.random() < 0.6, rng.random() < 0.3
        n = rng.randrange(1, 3000)
        data = gen_stream(harness, rng.choice([0, 5, 3, 2]), rng.randrange(1 << 30), n)
        e = oracle.Encoder(window=w, extended=ext, dictionary_reset=dr)
        r = oracle.RefCompressor(ref, window=w, extended=ext, dictionary_reset=dr)
        out, pos = b"", 0
        while pos < n:
            chunk = data[pos:pos + rng.choice([1, 2, 5, 16, 17, 100, 700])]
            pos += len(chunk)
            e.write(chunk)
            o, m, res = r.compress(chunk, 10000)
            assert res == 0 and m == len(chunk)
            out += o
            if rng.random() < 0.2:
                e.flush(True)
                o, res = r.flush(100, True)
                out += o
        e.flush(False)
        out += r.flush(100, False)[0]
        assert e.getvalue() == out and e.window() == r.window.raw
        assert oracle.decompress(out) == (data, oracle.INPUT_EXHAUSTED)


def _oracle_segmented(data, seg, **kw):
    """The segmented stream composed from the oracle: segment 0 = a dictionary_reset stream, every later segment the same
    with the append-mode marker (FLUSH padded to 16 bits, compressor.c:227-234) where its two header bytes would be."""
    out, offsets = b"", [0]
    for i in range(0, max(len(data), 1), seg):
        f = oracle.compress(data[i:i + seg], dictionary_reset=True, write_token=True, **kw)
        out += f if i == 0 else b"\x55\x80" + f[2:]
        offsets.append(len(out))
    return out, offsets


def test_segmented_streams_composed_from_the_oracle_match_the_reference_digests(harness):
    """tests/golden/ref_segmented.json was recorded from ONE unmodified reference compressor with
    tamp_compressor_reset_dictionary() between the segments (make_segmented_fixtures.py).  The oracle restatement,
    composed segment by segment, gives the same bytes and segment offsets, and decodes them front to back as one
    stream: this is what the GPU tests and bench.py's one_stream leg check the CUDA path against."""
    import hashlib
    import json
    from conftest import GOLDEN
    for c in json.loads((GOLDEN / "ref_segmented.json").read_text()):
        data = gen_stream(harness, c["gen"], c["k"], c["n"], c["literal"])
        assert hashlib.sha256(data).hexdigest() == c["input_sha256"]
        out, offsets = _oracle_segmented(data, c["segment_size"], window=c["window"], literal=c["literal"], extended=c["extended"])
        assert len(out) == c["size"] and hashlib.sha256(out).hexdigest() == c["sha256"]
        assert hashlib.sha256(json.dumps(offsets).encode()).hexdigest() == c["offsets_sha256"]
        assert oracle.decompress(out, window_bits_max=c["window"], cap=c["n"] + 64) == (data, oracle.INPUT_EXHAUSTED)


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built (no /root/reference here)")
def test_append_mode_frame_identity_vs_reference(harness):
    """An append-mode frame of the reference == 55 80 + the body of the dictionary_reset frame of the same input, for
    every non-empty input; the empty one is the marker alone (its FLUSH is still the last token: no second one)."""
    ref = oracle.Ref()
    rng = random.Random(11)
    for it in range(60):
        w, ext, wt = rng.choice([8, 10, 12, 15]), rng.random() < 0.5, rng.random() < 0.7
        data = gen_stream(harness, rng.choice([0, 5, 3, 2, 1]), rng.randrange(1 << 30), rng.choice([1, 2, 16, 17, 1000, 5000]))
        c = oracle.RefCompressor(ref, window=w, extended=ext, dictionary_reset=True, append=True)
        want, m, res = c.compress_and_flush(data, len(data) * 2 + 64, wt)
        assert res == 0 and want == b"\x55\x80" + oracle.compress(data, window=w, extended=ext, dictionary_reset=True, write_token=wt)[2:]
    c = oracle.RefCompressor(ref, window=10, dictionary_reset=True, append=True)
    assert c.compress_and_flush(b"", 64, True)[0] == b"\x55\x80"
