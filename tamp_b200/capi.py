"""Thin object wrappers over the Tamp C API exported by libtamp_b200.so.

Method names and argument meaning follow the C functions one to one (include/tamp/compressor.h,
include/tamp/decompressor.h), so that tests can replay the reference's ctests against the CUDA path.
Each object owns its state struct and window exactly like a C caller would.
"""
from __future__ import annotations

import ctypes as C

from . import _lib
from ._lib import TampCompressor, TampCompressorLazy, TampConf, TampConfLazy, TampDecompressor


def make_conf(window=10, literal=8, use_custom_dictionary=False, extended=False, dictionary_reset=False,
              append=False, lazy_matching=None):
    """TampConf image; pass lazy_matching (True/False) to get the TAMP_LAZY_MATCHING=1 layout."""
    if lazy_matching is None:
        return TampConf(window, literal, int(use_custom_dictionary), int(extended), int(dictionary_reset), int(append))
    return TampConfLazy(window, literal, int(use_custom_dictionary), int(extended), int(dictionary_reset), int(append),
                        int(lazy_matching))


class CCompressor:
    def __init__(self, *, window=10, literal=8, extended=True, dictionary=None, dictionary_reset=False,
                 append=False, default_conf=False, lazy_matching=None):
        """lazy_matching=None: default library flavour; True/False: the TAMP_LAZY_MATCHING=1 flavour."""
        self.L = _lib.lib(lazy=lazy_matching is not None)
        self.state = TampCompressor() if lazy_matching is None else TampCompressorLazy()
        self.window = C.create_string_buffer(1 << window)
        if dictionary is not None:
            if len(dictionary) != 1 << window:
                raise ValueError("dictionary must be 1 << window bytes")
            self.window.raw = bytes(dictionary)
        self.conf = make_conf(window, literal, dictionary is not None, extended, dictionary_reset, append, lazy_matching)
        self.init_res = self.L.tamp_compressor_init(C.byref(self.state), None if default_conf else C.byref(self.conf),
                                                    self.window)

    def sink(self, data: bytes) -> int:
        n = C.c_size_t(0)
        self.L.tamp_compressor_sink(C.byref(self.state), data, len(data), C.byref(n))
        return n.value

    def full(self) -> bool:
        return bool(self.L.tamp_compressor_full(C.byref(self.state)))

    def poll(self, cap: int):
        out = C.create_string_buffer(max(cap, 1))
        n = C.c_size_t(0)
        r = self.L.tamp_compressor_poll(C.byref(self.state), out, cap, C.byref(n))
        return out.raw[:n.value], r

    def flush(self, cap: int, write_token: bool):
        out = C.create_string_buffer(max(cap, 1))
        n = C.c_size_t(0)
        r = self.L.tamp_compressor_flush(C.byref(self.state), out, cap, C.byref(n), write_token)
        return out.raw[:n.value], r

    def compress(self, data: bytes, cap: int, callback=None):
        out = C.create_string_buffer(max(cap, 1))
        n, m = C.c_size_t(0), C.c_size_t(0)
        r = self.L.tamp_compressor_compress_cb(C.byref(self.state), out, cap, C.byref(n), data, len(data),
                                               C.byref(m), callback, None)
        return out.raw[:n.value], m.value, r

    def compress_and_flush(self, data: bytes, cap: int, write_token: bool, callback=None):
        out = C.create_string_buffer(max(cap, 1))
        n, m = C.c_size_t(0), C.c_size_t(0)
        r = self.L.tamp_compressor_compress_and_flush_cb(C.byref(self.state), out, cap, C.byref(n), data, len(data),
                                                         C.byref(m), write_token, callback, None)
        return out.raw[:n.value], m.value, r

    def reset_dictionary(self, cap: int):
        out = C.create_string_buffer(max(cap, 1))
        n = C.c_size_t(0)
        r = self.L.tamp_compressor_reset_dictionary(C.byref(self.state), out, cap, C.byref(n))
        return out.raw[:n.value], r

    def state_bytes(self) -> bytes:
        """State after the window pointer (what the reference traces record)."""
        return bytes(self.state)[8:]


class CDecompressor:
    def __init__(self, *, dictionary=None, window_bits=15, conf: TampConf | None = None):
        self.L = _lib.lib()
        self.state = TampDecompressor()
        self.window = C.create_string_buffer(1 << window_bits)
        if dictionary is not None:
            self.window.raw = bytes(dictionary) + bytes((1 << window_bits) - len(dictionary))
        self.init_res = self.L.tamp_decompressor_init(C.byref(self.state), C.byref(conf) if conf is not None else None,
                                                      self.window, window_bits)

    def decompress(self, data: bytes, cap: int, callback=None):
        out = C.create_string_buffer(max(cap, 1))
        n, m = C.c_size_t(0), C.c_size_t(0)
        r = self.L.tamp_decompressor_decompress_cb(C.byref(self.state), out, cap, C.byref(n), data, len(data),
                                                   C.byref(m), callback, None)
        return out.raw[:n.value], m.value, r

    def state_bytes(self) -> bytes:
        return bytes(self.state)[8:]


def read_header(data: bytes):
    conf = TampConf()
    n = C.c_size_t(0)
    r = _lib.lib().tamp_decompressor_read_header(C.byref(conf), data, len(data), C.byref(n))
    return conf, n.value, r


# ---- stream API (tamp_compress_stream / tamp_decompress_stream) -------------------------------------------------------

def _stream(fn, state, read, write, progress):
    """Drive a tamp_*_stream function with Python callables: read(n) -> bytes (b"" at the end, None = error),
    write(bytes) -> bytes accepted (negative = error), progress(done, total) -> non-zero aborts."""
    def c_read(_h, buf, size):
        chunk = read(size)
        if chunk is None:
            return -1
        C.memmove(buf, chunk, len(chunk))
        return len(chunk)

    def c_write(_h, buf, size):
        return write(C.string_at(buf, size))

    rcb, wcb = _lib.READ_CB(c_read), _lib.WRITE_CB(c_write)
    pcb = _lib.PROGRESS_CB(lambda _u, done, total: progress(done, total)) if progress else None
    consumed, written = C.c_size_t(0), C.c_size_t(0)
    res = fn(C.byref(state), C.cast(rcb, C.c_void_p), None, C.cast(wcb, C.c_void_p), None, C.byref(consumed),
             C.byref(written), C.cast(pcb, C.c_void_p) if pcb else None, None)
    return res, consumed.value, written.value


def compress_stream(comp: "CCompressor", read, write, progress=None):
    return _stream(comp.L.tamp_compress_stream, comp.state, read, write, progress)


def decompress_stream(dec: "CDecompressor", read, write, progress=None):
    return _stream(dec.L.tamp_decompress_stream, dec.state, read, write, progress)
