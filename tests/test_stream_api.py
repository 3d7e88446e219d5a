"""tamp_compress_stream / tamp_decompress_stream through the C ABI: callback readers and writers on
the host, the codec calls behind them on the GPU.  The bytes that reach the write callback, concatenated, must be the
reference's (reference compressor.c:891-955, decompressor.c:585-640); counters, error codes and the progress
callback's abort rule as documented in include/tamp/*.h."""
import ctypes as C
import io

import pytest

import oracle
from conftest import gen_stream
from tamp_b200 import _lib
from tamp_b200.capi import CCompressor, CDecompressor, compress_stream, decompress_stream


def _reader(data, step):
    f = io.BytesIO(data)
    return lambda size: f.read(min(size, step))


@pytest.mark.gpu
@pytest.mark.parametrize("window,extended,n,step", [(10, True, 50000, 1 << 20), (10, False, 20000, 777), (8, True, 3000, 1),
                                                    (12, True, 40000, 5000), (10, True, 0, 64)])
def test_stream_round_trip_matches_the_reference_bytes(harness, window, extended, n, step):
    data = gen_stream(harness, 0, 11, n)
    want = oracle.compress(data, window=window, extended=extended)
    out = bytearray()
    seen = []
    comp = CCompressor(window=window, extended=extended)
    res, consumed, written = compress_stream(comp, _reader(data, step), lambda b: (out.extend(b), len(b))[1],
                                             lambda done, total: seen.append((done, total)) or 0)
    assert (res, consumed, written) == (_lib.OK, n, len(want))
    assert bytes(out) == want
    assert all(t == 0 for _, t in seen) and [d for d, _ in seen] == sorted(d for d, _ in seen)
    if n:
        assert seen[-1][0] == n

    back = bytearray()
    dec = CDecompressor(window_bits=window)
    res, consumed, written = decompress_stream(dec, _reader(want, max(step // 2, 1)), lambda b: (back.extend(b), len(b))[1])
    assert (res, consumed, written) == (_lib.OK, len(want), n)
    assert bytes(back) == data


@pytest.mark.gpu
def test_stream_errors_and_abort(harness):
    data = gen_stream(harness, 0, 12, 40000)
    # read error, short write, negative write, progress abort (user codes live in [100, 127])
    assert compress_stream(CCompressor(window=10), lambda size: None, lambda b: len(b))[0] == _lib.READ_ERROR
    assert compress_stream(CCompressor(window=10), _reader(data, 4096), lambda b: len(b) - 1)[0] == _lib.WRITE_ERROR
    assert compress_stream(CCompressor(window=10), _reader(data, 4096), lambda b: -1)[0] == _lib.WRITE_ERROR
    res, consumed, _ = compress_stream(CCompressor(window=10), _reader(data, 4096), lambda b: len(b), lambda d, t: 101)
    assert res == 101 and 0 < consumed <= 16384
    # literal that does not fit: the codec's own error comes through
    bad = bytes([0x41] * 100 + [0xF0] + [0x41] * 100)
    assert compress_stream(CCompressor(window=10, literal=7), _reader(bad, 50), lambda b: len(b))[0] == _lib.EXCESS_BITS
    frame = oracle.compress(data, window=10)
    assert decompress_stream(CDecompressor(window_bits=10), lambda size: None, lambda b: len(b))[0] == _lib.READ_ERROR
    assert decompress_stream(CDecompressor(window_bits=10), _reader(frame, 999), lambda b: 0)[0] == _lib.WRITE_ERROR
    assert decompress_stream(CDecompressor(window_bits=10), _reader(frame, 999), lambda b: len(b), lambda d, t: -100)[0] == -100
    # a frame whose window does not fit the decompressor's buffer
    assert decompress_stream(CDecompressor(window_bits=8), _reader(frame, 999), lambda b: len(b))[0] == _lib.INVALID_CONF


@pytest.mark.gpu
def test_memory_and_stdio_handlers_drive_a_stream(harness, tmp_path):
    """The built-in handlers as a C caller would use them: TampMemReader -> FILE*, then FILE* -> TampMemWriter."""
    L = _lib.lib()
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    data = gen_stream(harness, 0, 13, 30000)
    src = C.create_string_buffer(data, len(data))
    reader = _lib.TampMemReader(C.cast(src, C.c_void_p), len(data), 0)
    path = str(tmp_path / "out.tamp").encode()
    f = libc.fopen(path, b"wb")
    comp = CCompressor(window=10)
    consumed, written = C.c_size_t(0), C.c_size_t(0)
    res = L.tamp_compress_stream(C.byref(comp.state), C.cast(L.tamp_stream_mem_read, C.c_void_p), C.byref(reader),
                                 C.cast(L.tamp_stream_stdio_write, C.c_void_p), f, C.byref(consumed), C.byref(written),
                                 None, None)
    libc.fclose(f)
    want = oracle.compress(data, window=10)
    assert res == _lib.OK and consumed.value == len(data) and written.value == len(want)
    assert open(path, "rb").read() == want

    dst = C.create_string_buffer(len(data))
    writer = _lib.TampMemWriter(C.cast(dst, C.c_void_p), len(data), 0)
    f = libc.fopen(path, b"rb")
    dec = CDecompressor(window_bits=10)
    res = L.tamp_decompress_stream(C.byref(dec.state), C.cast(L.tamp_stream_stdio_read, C.c_void_p), f,
                                   C.cast(L.tamp_stream_mem_write, C.c_void_p), C.byref(writer), None, None, None, None)
    libc.fclose(f)
    assert res == _lib.OK and writer.pos == len(data) and dst.raw == data
    # a writer that is one byte short refuses the chunk that would overflow it
    small = _lib.TampMemWriter(C.cast(dst, C.c_void_p), len(data) - 1, 0)
    f = libc.fopen(path, b"rb")
    dec = CDecompressor(window_bits=10)
    res = L.tamp_decompress_stream(C.byref(dec.state), C.cast(L.tamp_stream_stdio_read, C.c_void_p), f,
                                   C.cast(L.tamp_stream_mem_write, C.c_void_p), C.byref(small), None, None, None, None)
    libc.fclose(f)
    assert res == _lib.WRITE_ERROR


def test_memory_handlers_on_the_host():
    """No GPU needed: the handlers are plain host code."""
    L = _lib.lib()
    src = C.create_string_buffer(b"0123456789", 10)
    r = _lib.TampMemReader(C.cast(src, C.c_void_p), 10, 0)
    buf = C.create_string_buffer(8)
    assert L.tamp_stream_mem_read(C.byref(r), buf, 8) == 8 and buf.raw == b"01234567"
    assert L.tamp_stream_mem_read(C.byref(r), buf, 8) == 2 and buf.raw[:2] == b"89"
    assert L.tamp_stream_mem_read(C.byref(r), buf, 8) == 0
    dst = C.create_string_buffer(6)
    w = _lib.TampMemWriter(C.cast(dst, C.c_void_p), 6, 0)
    assert L.tamp_stream_mem_write(C.byref(w), b"abcd", 4) == 4
    assert L.tamp_stream_mem_write(C.byref(w), b"efg", 3) == -1 and w.pos == 4
    assert L.tamp_stream_mem_write(C.byref(w), b"ef", 2) == 2 and dst.raw == b"abcdef"


def test_stream_calls_fail_loudly_without_gpu():
    L = _lib.lib()
    if L.tamp_b200_device_count() > 0:
        pytest.skip("CUDA device present")
    res, _, _ = compress_stream(CCompressor(window=10), _reader(b"hello hello hello hello", 64), lambda b: len(b))
    assert res == _lib.ERROR and "no CUDA device" in _lib.last_error()
