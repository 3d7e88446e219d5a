"""File-like Python binding over the drop-in C API — the interface of the reference's Cython classes
(tamp/_c_compressor.pyx:13-199, tamp/_c_decompressor.pyx:12-187) running on the CUDA path.

Same constructor arguments, method names, return values and exception mapping
(tamp/_c_common.pyx: OUTPUT_FULL/INPUT_EXHAUSTED -> IndexError, EXCESS_BITS -> ExcessBitsError,
INVALID_CONF/OOB -> ValueError), so the reference's tests/test_compressor.py style cases can be run against
it unchanged.  Every write()/flush()/read() is one or more calls of tamp_compressor_compress /
tamp_compressor_flush / tamp_decompressor_decompress in libtamp_b200.so (CUDA, batch of one).
"""
from __future__ import annotations

import ctypes as C
from io import BytesIO

from . import _lib

CHUNK_SIZE = 1 << 20


class ExcessBitsError(Exception):
    """Provided data has more bits than expected ``literal`` bits."""


ERROR_LOOKUP = {
    _lib.OUTPUT_FULL: IndexError,
    _lib.INPUT_EXHAUSTED: IndexError,
    _lib.ERROR: RuntimeError,
    _lib.EXCESS_BITS: ExcessBitsError,
    _lib.INVALID_CONF: ValueError,
    _lib.OOB: ValueError,
}


def _raise(res):
    exc = ERROR_LOOKUP.get(res, NotImplementedError)
    if res == _lib.ERROR:
        raise exc(_lib.last_error())
    raise exc


class Compressor:
    def __init__(self, f, *, window=10, literal=8, dictionary=None, lazy_matching=False, extended=True,
                 dictionary_reset=False, append=False):
        if dictionary is not None and len(dictionary) != (1 << window):
            raise ValueError("Dictionary-window size mismatch.")
        if not hasattr(f, "write"):  # path-like
            f = open(str(f), "wb")
            self._close_f_on_close = True
        else:
            self._close_f_on_close = False
        self.f = f
        # lazy matching lives in the TAMP_LAZY_MATCHING=1 flavour of the library (different conf layout)
        self._lazy = bool(lazy_matching)
        self._L = _lib.lib(lazy=self._lazy)
        self._state = (_lib.TampCompressorLazy if self._lazy else _lib.TampCompressor)()
        if not 0 <= window <= 15 or not 0 <= literal <= 15:
            raise ValueError
        if self._lazy:
            conf = _lib.TampConfLazy(window, literal, int(dictionary is not None), int(extended),
                                     int(dictionary_reset), int(append), 1)
        else:
            conf = _lib.TampConf(window, literal, int(dictionary is not None), int(extended), int(dictionary_reset),
                                 int(append))
        if dictionary is not None:
            if not isinstance(dictionary, bytearray):
                dictionary = bytearray(dictionary)
            self._window_buffer = dictionary  # used in place, like the reference
        else:
            self._window_buffer = bytearray(1 << max(window, 8))
        self._window = (C.c_char * len(self._window_buffer)).from_buffer(self._window_buffer)
        self._dictionary_reset = bool(dictionary_reset)
        res = self._L.tamp_compressor_init(C.byref(self._state), C.byref(conf), self._window)
        if res < 0:
            _raise(res)

    def write(self, data) -> int:
        data = bytes(data)
        remaining = len(data)
        if remaining == 0:
            return 0
        out = C.create_string_buffer(CHUNK_SIZE)
        view = memoryview(data)
        # one call never consumes more input than CHUNK_SIZE of output holds (a literal costs at most 9 bits): hand the
        # compressor that much and no more, so a large write is not re-sliced and re-uploaded on every iteration
        step = CHUNK_SIZE * 8 // 9
        pos = 0
        written_total = 0
        while remaining:
            n_out, n_in = C.c_size_t(0), C.c_size_t(0)
            take = min(remaining, step)
            piece = bytes(view[pos:pos + take])
            res = self._L.tamp_compressor_compress_cb(C.byref(self._state), out, CHUNK_SIZE, C.byref(n_out),
                                                      piece, take, C.byref(n_in), None, None)
            if res < 0:
                _raise(res)
            self.f.write(out.raw[:n_out.value])
            written_total += n_out.value
            pos += n_in.value
            remaining -= n_in.value
        return written_total

    def flush(self, write_token: bool = True) -> int:
        out = C.create_string_buffer(32)
        n_out = C.c_size_t(0)
        res = self._L.tamp_compressor_flush(C.byref(self._state), out, 32, C.byref(n_out), write_token)
        if res < 0:
            _raise(res)
        if n_out.value:
            self.f.write(out.raw[:n_out.value])
        self.f.flush()
        return n_out.value

    def reset_dictionary(self) -> int:
        out = C.create_string_buffer(32)
        n_out = C.c_size_t(0)
        res = self._L.tamp_compressor_reset_dictionary(C.byref(self._state), out, 32, C.byref(n_out))
        if res < 0:
            _raise(res)
        if n_out.value:
            self.f.write(out.raw[:n_out.value])
        self.f.flush()
        return n_out.value

    def close(self) -> int:
        # with dictionary_reset, always end with a FLUSH so an append-mode compressor can form a double FLUSH
        n = self.flush(write_token=self._dictionary_reset)
        if self._close_f_on_close:
            self.f.close()
        return n

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_value, traceback):
        self.close()


class TextCompressor(Compressor):
    def write(self, data: str) -> int:
        return super().write(data.encode())


def compress(data, *args, **kwargs) -> bytes:
    with BytesIO() as f:
        c = (TextCompressor if isinstance(data, str) else Compressor)(f, *args, **kwargs)
        c.write(data)
        c.flush(write_token=False)
        return f.getvalue()


class Decompressor:
    def __init__(self, f, *, dictionary=None):
        if not hasattr(f, "read"):
            f = open(str(f), "rb")
            self._close_f_on_close = True
        else:
            self._close_f_on_close = False
        self.f = f
        self._L = _lib.lib()
        self._input = b""
        self._input_pos = 0
        conf = _lib.TampConf()
        header = bytearray()
        while True:  # read the header one byte at a time: no seek support needed
            b = f.read(1)
            if not b:
                _raise(_lib.INPUT_EXHAUSTED)
            header += b
            used = C.c_size_t(0)
            res = self._L.tamp_decompressor_read_header(C.byref(conf), bytes(header), len(header), C.byref(used))
            if res == _lib.OK:
                break
            if res != _lib.INPUT_EXHAUSTED:
                _raise(res)
        if conf.use_custom_dictionary and dictionary is None:
            raise ValueError
        if dictionary is not None and len(dictionary) < (1 << conf.window):
            raise ValueError("Dictionary-window size mismatch.")
        if dictionary is not None:
            if not isinstance(dictionary, bytearray):
                dictionary = bytearray(dictionary)
            self._window_buffer = dictionary
        else:
            self._window_buffer = bytearray(1 << conf.window)
        self._window = (C.c_char * len(self._window_buffer)).from_buffer(self._window_buffer)
        self._state = _lib.TampDecompressor()
        res = self._L.tamp_decompressor_init(C.byref(self._state), C.byref(conf), self._window, conf.window)
        if res < 0:
            _raise(res)

    def readinto(self, buf: bytearray) -> int:
        size = len(buf)
        done = 0
        while done < size:
            want = min(CHUNK_SIZE, size - done)
            out = C.create_string_buffer(want)
            n_out, n_in = C.c_size_t(0), C.c_size_t(0)
            chunk = self._input[self._input_pos:]
            res = self._L.tamp_decompressor_decompress_cb(C.byref(self._state), out, want, C.byref(n_out), chunk,
                                                          len(chunk), C.byref(n_in), None, None)
            self._input_pos += n_in.value
            buf[done:done + n_out.value] = out.raw[:n_out.value]
            done += n_out.value
            if res == _lib.INPUT_EXHAUSTED:
                self._input = self.f.read(CHUNK_SIZE)
                self._input_pos = 0
                if len(self._input) > CHUNK_SIZE:
                    raise ValueError("read() returned more bytes than requested.")
                if not self._input:
                    break
            elif res < 0:
                _raise(res)
        return done

    def read(self, size: int = -1) -> bytearray:
        if size == 0:
            return bytearray()
        chunk_size = CHUNK_SIZE
        out = []
        while True:
            buf = bytearray(chunk_size if size < 0 else size)
            chunk_size <<= 1
            n = self.readinto(buf)
            if size > 0:
                del buf[n:]
                out.append(buf)
                break
            if n < len(buf):
                if n:
                    out.append(buf[:n])
                break
            out.append(buf)
        return out[0] if len(out) == 1 else bytearray(b"".join(out))

    def close(self):
        if self._close_f_on_close:
            self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_value, traceback):
        self.close()


class TextDecompressor(Decompressor):
    def read(self, *args, **kwargs) -> str:
        return super().read(*args, **kwargs).decode()


def decompress(data: bytes, *args, **kwargs) -> bytearray:
    with BytesIO(bytes(data)) as f:
        return Decompressor(f, *args, **kwargs).read()


def open(f, mode="rb", **kwargs):  # noqa: A001 - mirrors tamp.open (tamp/__init__.py:93-105)
    if "r" in mode and "w" in mode:
        raise ValueError
    if "r" in mode:
        return Decompressor(f, **kwargs) if "b" in mode else TextDecompressor(f, **kwargs)
    if "w" in mode:
        return Compressor(f, **kwargs) if "b" in mode else TextCompressor(f, **kwargs)
    raise ValueError


def compress_batch(chunks, *, window: int = 10, literal: int = 8, dictionary=None, extended: bool = True) -> list[bytes]:
    """``[tamp.compress(c, window=..., literal=..., dictionary=..., extended=...) for c in chunks]`` in one batch on the
    GPU: every element is an independent Tamp stream, bit-identical to what the reference's
    ``tamp.compress`` returns for it (`tamp/_c_compressor.pyx:97-116`; no flush token, as there)."""
    import numpy as np
    import torch

    from . import batch
    chunks = [bytes(c) for c in chunks]
    if not chunks:
        return []
    stride = max(16, (max(len(c) for c in chunks) + 15) // 16 * 16)
    rows = np.zeros((len(chunks), stride), np.uint8)
    sizes = np.array([len(c) for c in chunks], np.int32)
    for i, c in enumerate(chunks):
        rows[i, :len(c)] = np.frombuffer(c, np.uint8)
    dic = None if dictionary is None else torch.frombuffer(bytearray(dictionary), dtype=torch.uint8).cuda()
    r = batch.compress_batch(torch.from_numpy(rows).cuda(), window=window, literal=literal, extended=extended,
                             dictionary=dic, sizes=torch.from_numpy(sizes).cuda())
    st = r.status.cpu().numpy()
    for i in np.nonzero(st != 0)[0][:1]:
        _raise(int(st[i]))
    data, osz = r.data.cpu().numpy(), r.sizes.cpu().numpy()
    return [data[i, :osz[i]].tobytes() for i in range(len(chunks))]


def decompress_batch(frames, max_size: int, *, dictionary=None) -> list[bytes]:
    """``[tamp.decompress(f, dictionary=...) for f in frames]`` in one batch on the GPU; ``max_size`` bounds the output
    of one stream (a stream that produces more raises, as a full output buffer would)."""
    import numpy as np
    import torch

    from . import batch
    frames = [bytes(f) for f in frames]
    if not frames:
        return []
    stride = max(16, (max(len(f) for f in frames) + 15) // 16 * 16)
    rows = np.zeros((len(frames), stride), np.uint8)
    sizes = np.array([len(f) for f in frames], np.int32)
    for i, f in enumerate(frames):
        rows[i, :len(f)] = np.frombuffer(f, np.uint8)
    dic = None if dictionary is None else torch.frombuffer(bytearray(dictionary), dtype=torch.uint8).cuda()
    cap = (max_size + 1 + 15) // 16 * 16  # one byte of room: a stream of exactly max_size bytes still ends INPUT_EXHAUSTED
    r = batch.decompress_batch(torch.from_numpy(rows).cuda(), torch.from_numpy(sizes).cuda(), cap, dictionary=dic)
    st, osz = r.status.cpu().numpy(), r.sizes.cpu().numpy()
    for i in range(len(frames)):
        if st[i] < 0:
            _raise(int(st[i]))
        if st[i] == _lib.OUTPUT_FULL or osz[i] > max_size:
            raise ValueError(f"stream {i} decompresses to more than max_size = {max_size} bytes")
    data = r.data.cpu().numpy()
    return [data[i, :osz[i]].tobytes() for i in range(len(frames))]
