"""The file-like Python binding (tamp_b200.binding), exercised the way the reference's own
tests/test_compressor.py / test_decompressor.py / test_compressor_decompressor.py exercise tamp.Compressor
and tamp.Decompressor — same vectors, same call patterns — but on the CUDA path."""
import io
import random

import pytest

import oracle
import tamp_b200
from conftest import gen_stream

pytestmark = pytest.mark.gpu


def test_compressor_default_and_chunked_writes():
    # tests/test_compressor.py:66-143
    expected = bytes.fromhex("58b3041c8100030000")
    with io.BytesIO() as f:
        c = tamp_b200.Compressor(f, extended=False)
        n = c.write(b"foo foo foo")
        n += c.flush(write_token=False)
        assert f.getvalue() == expected and n == len(expected)
    with io.BytesIO() as f, tamp_b200.Compressor(f, extended=False) as c:
        for piece in (b"f", b"oo", b" fo", b"o foo"):
            c.write(piece)
        c.flush(write_token=False)
        assert f.getvalue() == expected


def test_compressor_7bit_custom_dictionary_and_errors():
    # tests/test_compressor.py:145-237, :239-300
    assert tamp_b200.compress(b"foo foo foo", literal=7, extended=False).hex() == "50e6083a04000c00"
    dictionary = bytearray(256)
    dictionary[:11] = b"foo foo foo"
    assert tamp_b200.compress(b"foo foo foo", window=8, literal=7, dictionary=dictionary, extended=False).hex() == "145400"
    with pytest.raises(ValueError):
        tamp_b200.Compressor(io.BytesIO(), window=9, literal=7, dictionary=bytearray(256))
    with pytest.raises(tamp_b200.ExcessBitsError):
        tamp_b200.compress(b"\xff" * 40, literal=7)
    for bad in (dict(window=7), dict(window=16), dict(literal=4), dict(literal=9)):
        with pytest.raises(ValueError):
            tamp_b200.Compressor(io.BytesIO(), **bad)


def test_extended_vectors_and_text_mode():
    # tests/test_compressor.py:313-425
    assert tamp_b200.compress(b"A" * 20).hex() == "5aa0aab1"
    assert tamp_b200.compress(b"B" * 5).hex() == "5aa12a84"
    d = bytearray(256)
    d[:16] = b"abcdefghijklmnop"
    assert tamp_b200.compress(b"abcdefghijklmnop", window=8, dictionary=bytearray(d)).hex() == "1e4e4000"
    assert tamp_b200.compress("foo foo foo", extended=False).hex() == "58b3041c8100030000"
    with io.BytesIO(bytes.fromhex("58b3041c8100030000")) as f:
        assert tamp_b200.open(f, "r").read() == "foo foo foo"


def test_decompressor_reads():
    # tests/test_decompressor.py:30-158
    comp = bytes.fromhex("58b3041c8100030000")
    assert tamp_b200.decompress(comp) == b"foo foo foo"
    with io.BytesIO(comp) as f:
        d = tamp_b200.Decompressor(f)
        assert d.read(4) == b"foo " and d.read(2) == b"fo" and d.read(-1) == b"o foo"
    assert tamp_b200.decompress(bytes.fromhex("58a8aac0abaac0")) == b"QW"            # FLUSH handling
    custom = bytearray(1024)
    custom[:4] = b"abcd"
    data = bytes.fromhex("5cb0b000")                                                 # overlap: "aabc" not "aaaa"
    assert tamp_b200.decompress(data, dictionary=bytearray(custom)) == b"aabc"
    with io.BytesIO(data) as f:
        d = tamp_b200.Decompressor(f, dictionary=bytearray(custom))
        assert [bytes(d.read(1)) for _ in range(4)] == [b"a", b"a", b"b", b"c"]
    with pytest.raises(ValueError):
        tamp_b200.Decompressor(io.BytesIO(bytes([0b000_10_1_0_0])))                  # dictionary required
    with pytest.raises(ValueError):
        tamp_b200.decompress(bytes.fromhex("583ff0"))                                # hostile offset -> OOB


def test_round_trips_and_dictionary_reset(harness):
    # tests/test_compressor_decompressor.py: every conf round-trips; reset_dictionary keeps decoding in sync
    rng = random.Random(0)
    for _ in range(12):
        w, lit = rng.choice([8, 10, 12, 15]), rng.choice([7, 8])
        ext, lazy = rng.random() < 0.5, rng.random() < 0.4
        data = gen_stream(harness, rng.choice([0, 5]), rng.randrange(1 << 20), rng.randrange(1, 5000), lit)
        comp = tamp_b200.compress(data, window=w, literal=lit, extended=ext, lazy_matching=lazy)
        assert comp == oracle.compress(data, window=w, literal=lit, extended=ext, lazy_matching=lazy)
        assert tamp_b200.decompress(comp) == data
    data = gen_stream(harness, 0, 77, 4000)
    with io.BytesIO() as f:
        c = tamp_b200.Compressor(f, dictionary_reset=True)
        c.write(data[:1500])
        c.reset_dictionary()
        c.write(data[1500:])
        c.close()
        comp = f.getvalue()
    assert tamp_b200.decompress(comp) == data
    assert oracle.decompress(comp)[0] == data


def test_list_of_bytes_batch_calls(harness):
    """tamp_b200.compress_batch(list[bytes]) == [tamp.compress(c) for c in chunks], and back."""
    import random
    rng = random.Random(3)
    chunks = [gen_stream(harness, rng.randrange(6), 500 + i, rng.choice([0, 1, 17, 300, 1024, 1500, 5000])) for i in range(200)]
    for kw in (dict(), dict(extended=False), dict(window=12, extended=False), dict(window=8, literal=7)):
        data = [bytes(b & 127 for b in c) for c in chunks] if kw.get("literal") == 7 else chunks
        got = tamp_b200.compress_batch(data, **kw)
        okw = dict(window=kw.get("window", 10), literal=kw.get("literal", 8), extended=kw.get("extended", True))
        assert got == [oracle.compress(c, **okw) for c in data], kw
        assert tamp_b200.decompress_batch(got, 5000) == data
    assert tamp_b200.compress_batch([]) == [] and tamp_b200.decompress_batch([], 10) == []
    with pytest.raises(ValueError):
        tamp_b200.decompress_batch(tamp_b200.compress_batch([b"x" * 100]), 50)
    with pytest.raises(tamp_b200.ExcessBitsError):
        tamp_b200.compress_batch([b"abc", b"ab\xffc"], literal=7)
