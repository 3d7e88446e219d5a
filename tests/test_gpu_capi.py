"""GPU parity of the per-call drop-in C API (tamp_compressor_* / tamp_decompressor_*), called through
the C-ABI of libtamp_b200.so.  Every call is a CUDA kernel launch on a batch of one."""
import hashlib
import random

import pytest

import oracle
from conftest import gen_stream
from tamp_b200 import _lib
from tamp_b200.capi import CCompressor, CDecompressor

pytestmark = pytest.mark.gpu


def _conf(k):
    conf = dict(k["conf"])
    if "dictionary" in conf:
        conf["dictionary"] = bytes.fromhex(conf["dictionary"])
    conf.pop("lazy_matching", None)
    return conf


def test_kats_through_c_api(kats):
    """The reference's own golden bitstreams, via init + compress_and_flush / init + decompress."""
    n = 0
    for k in kats:
        if k["kind"] == "compress":
            conf = _conf(k)
            c = CCompressor(**conf)
            data = bytes.fromhex(k["input"])
            out, consumed, res = c.compress_and_flush(data, 64, False)
            assert (res, consumed, out.hex()) == (0, len(data), k["expected"]), k["name"]
            d = CDecompressor(dictionary=conf.get("dictionary"), window_bits=conf["window"])
            back, _, res = d.decompress(out, 64)
            assert (res, back) == (_lib.INPUT_EXHAUSTED, data), k["name"]
            n += 1
        elif k["kind"] == "decompress":
            dic = bytes.fromhex(k["dictionary"]) if k["dictionary"] else None
            wb = k["window_bits_max"] if dic is None else (len(dic).bit_length() - 1)
            d = CDecompressor(dictionary=dic, window_bits=wb)
            comp = bytes.fromhex(k["input"])
            back, _, res = d.decompress(comp, 64)
            assert (res, back.hex()) == (k["status"], k["expected"]), k["name"]
            n += 1
    assert n >= 18
    assert _lib.lib().tamp_b200_launch_count() >= n


def test_decompress_byte_by_byte():
    """ctests/test_decompressor.c:14-77: feed one input byte per call."""
    comp = bytes.fromhex("58b3041c8100030000")
    d = CDecompressor(window_bits=10)
    out = b""
    for i in range(len(comp)):
        chunk, consumed, res = d.decompress(comp[i:i + 1], 32 - len(out))
        assert res >= 0 and consumed == 1
        out += chunk
    assert out == b"foo foo foo"


def test_api_traces_replay(api_sequences, harness):
    """Call-by-call traces recorded from the reference C (tests/golden/make_fixtures.py): the same calls
    on the CUDA-backed API must return the same bytes, status, consumed counts, and leave the same
    state struct and window behind."""
    n_calls = 0
    for s in api_sequences:
        if s["kind"] == "compress":
            data = gen_stream(harness, s["gen"], s["k"], s["n"])
            c = CCompressor(window=s["window"], extended=s["extended"], dictionary_reset=s["dictionary_reset"])
            pos = 0
            for op in s["ops"]:
                n_calls += 1
                if op["op"] == "sink":
                    took = c.sink(data[pos:pos + op["n"]])
                    assert took == op["consumed"]
                    pos += took
                elif op["op"] == "poll":
                    out, res = c.poll(op["cap"])
                    assert (out.hex(), res) == (op["out"], op["res"]), (op, pos)
                elif op["op"] == "compress":
                    out, took, res = c.compress(data[pos:pos + op["n"]], op["cap"])
                    assert (out.hex(), took, res) == (op["out"], op["consumed"], op["res"]), (op, pos)
                    pos += took
                else:
                    out, res = c.flush(op["cap"], op["write_token"])
                    assert (out.hex(), res) == (op["out"], op["res"]), (op, pos)
            assert c.state_bytes().hex() == s["final_state"]
            assert hashlib.sha256(c.window.raw).hexdigest() == s["final_window_sha"]
        else:
            comp = bytes.fromhex(s["comp"])
            d = CDecompressor(window_bits=s["window"])
            pos = 0
            for op in s["ops"]:
                n_calls += 1
                out, took, res = d.decompress(comp[pos:pos + op["n"]], op["cap"])
                assert (out.hex(), took, res) == (op["out"], op["consumed"], op["res"]), (op, pos)
                pos += took
            assert d.state_bytes().hex() == s["final_state"]
            assert hashlib.sha256(d.window.raw).hexdigest() == s["final_window_sha"]
    assert n_calls > 1000


def test_one_shot_vs_oracle_all_windows(harness):
    rng = random.Random(11)
    for w in range(8, 16):
        for ext in (False, True):
            lit = rng.choice([5, 6, 7, 8, 8])
            n = rng.choice([0, 1, 17, 700, 3000])
            data = gen_stream(harness, rng.randrange(6), rng.randrange(1 << 20), n, lit)
            c = CCompressor(window=w, literal=lit, extended=ext)
            out, consumed, res = c.compress_and_flush(data, len(data) * 2 + 64, False)
            assert res == 0 and consumed == n
            assert out == oracle.compress(data, window=w, literal=lit, extended=ext), (w, ext, lit, n)
            d = CDecompressor(window_bits=w)
            back, _, res = d.decompress(out, n + 8)
            assert (back, res) == (data, _lib.INPUT_EXHAUSTED)


def test_reset_dictionary_and_mid_stream_flush(harness):
    """compressor.c:847-881 + decompressor.c:501-514: segments cut by double-FLUSH."""
    data = gen_stream(harness, oracle.TEXT, 5, 3000)
    c = CCompressor(window=10, extended=True, dictionary_reset=True)
    out = b""
    for i in range(0, 3000, 1000):
        o, took, res = c.compress(data[i:i + 1000], 4000)
        assert res == 0 and took == 1000
        out += o
        o, res = c.reset_dictionary(64)
        assert res == 0
        out += o
    o, res = c.flush(64, False)
    out += o
    assert oracle.decompress(out) == (data, oracle.INPUT_EXHAUSTED)
    d = CDecompressor(window_bits=10)
    back, _, res = d.decompress(out, 4000)
    assert (back, res) == (data, _lib.INPUT_EXHAUSTED)
    if oracle.ref_available():
        r = oracle.RefCompressor(oracle.Ref(), window=10, extended=True, dictionary_reset=True)
        exp = b""
        for i in range(0, 3000, 1000):
            exp += r.compress(data[i:i + 1000], 4000)[0]
            exp += r.reset_dictionary(64)[0]
        exp += r.flush(64, False)[0]
        assert out == exp


@pytest.mark.parametrize("window,extended", [(10, True), (10, False), (12, True), (8, False)])
def test_append_mode_segments(harness, window, extended):
    """compressor.c:227-234: a compressor opened with `append` starts with FLUSH padded to two bytes
    instead of a header, so that its output, appended to a dictionary_reset stream that ended with a FLUSH token,
    reads as one stream (decompressor.c:501-514 resets the dictionary on the double FLUSH).  Bytes against the
    reference C, call by call; the concatenation against the oracle decoder and through the CUDA decoder."""
    parts = [gen_stream(harness, oracle.TEXT, 21 + i, n) for i, n in enumerate((1500, 700, 1, 2300))]
    ref = oracle.Ref() if oracle.ref_available() else None
    stream = b""
    for i, data in enumerate(parts):
        kw = dict(window=window, extended=extended, dictionary_reset=True, append=i > 0)
        c = CCompressor(**kw)
        if i > 0:
            assert c.state_bytes() != CCompressor(**dict(kw, append=False)).state_bytes()
        out, took, res = c.compress(data, 2 * len(data) + 64)
        assert res == 0 and took == len(data)
        tail, res = c.flush(64, True)  # the trailing FLUSH the next segment's leading FLUSH pairs with
        assert res == 0
        seg = out + tail
        if i > 0:
            assert seg[:2] == bytes([0xAB >> 1, (0xAB & 1) << 7])  # 9-bit FLUSH code, zero padded to 16 bits
        if ref is not None:
            r = oracle.RefCompressor(ref, **kw)
            exp = r.compress(data, 2 * len(data) + 64)[0] + r.flush(64, True)[0]
            assert seg == exp, (i, window, extended)
            assert c.state_bytes() == r.state.raw[8:]
        stream += seg
    whole = b"".join(parts)
    assert oracle.decompress(stream) == (whole, oracle.INPUT_EXHAUSTED)
    d = CDecompressor(window_bits=window)
    back, _, res = d.decompress(stream, len(whole) + 16)
    assert (back, res) == (whole, _lib.INPUT_EXHAUSTED)
    # an immediate flush on a fresh append compressor must not add a second FLUSH (last_was_flush, compressor.c:232)
    c = CCompressor(window=window, extended=extended, dictionary_reset=True, append=True)
    out, res = c.flush(16, True)
    assert res == 0 and out == bytes([0xAB >> 1, (0xAB & 1) << 7])
    if ref is not None:
        assert oracle.RefCompressor(ref, window=window, extended=extended, dictionary_reset=True,
                                    append=True).flush(16, True)[0] == out


def test_excess_bits_and_callbacks():
    import ctypes as C
    c = CCompressor(window=10, literal=7, extended=False)
    out, consumed, res = c.compress_and_flush(b"abc\xff" + b"d" * 20, 100, False)
    assert res == _lib.EXCESS_BITS
    calls = []
    CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_size_t, C.c_size_t)
    cb = CB(lambda u, a, b: calls.append((a, b)) or 0)
    c = CCompressor(window=10)
    out, consumed, res = c.compress_and_flush(b"hello hello hello hello", 100, False, callback=cb)
    assert res == 0 and calls[-1] == (23, 23)
    abort = CB(lambda u, a, b: 101)
    c = CCompressor(window=10)
    assert c.compress_and_flush(b"hello hello hello hello", 100, False, callback=abort)[2] == 101


def test_lazy_matching_flavour(ref_fixtures, harness):
    """lazy matching (compressor.c:176-189, :576-619) through the TAMP_LAZY_MATCHING=1
    flavour of the library, against digests recorded from the reference built the same way."""
    import hashlib
    from tamp_b200 import batch
    import numpy as np
    import torch
    n = 0
    for f in ref_fixtures:
        conf = dict(f["conf"])
        if not conf.get("lazy_matching"):
            continue
        data = gen_stream(harness, f["gen"], f["k"], f["n"])
        c = CCompressor(window=conf["window"], literal=conf.get("literal", 8), extended=conf["extended"],
                        lazy_matching=True)
        out, consumed, res = c.compress_and_flush(data, len(data) * 2 + 64, False)
        assert res == 0 and consumed == len(data)
        assert len(out) == f["size"] and hashlib.sha256(out).hexdigest() == f["sha"], conf
        assert out == oracle.compress(data, window=conf["window"], extended=conf["extended"], lazy_matching=True)
        # same through the batch entry point (general kernel: the specialised ones do not do lazy matching)
        x = torch.from_numpy(np.frombuffer(data, dtype=np.uint8).copy())[None, :].cuda()
        r = batch.compress_batch(x, window=conf["window"], extended=conf["extended"], lazy_matching=True)
        torch.cuda.synchronize()
        assert bytes(r.data[0, :int(r.sizes[0])].cpu().numpy()) == out
        # lazy_matching = False in the lazy flavour equals the default flavour
        c0 = CCompressor(window=conf["window"], extended=conf["extended"], lazy_matching=False)
        assert c0.compress_and_flush(data, len(data) * 2 + 64, False)[0] == oracle.compress(
            data, window=conf["window"], extended=conf["extended"])
        n += 1
    assert n >= 12
