"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

ctypes bindings for the CPU oracle (``oracle/tamp_oracle.c``), for the unmodified reference C
compiled into ``oracle/_ref`` (when present), and for the threaded CPU batch harness.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  ``tamp_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
BUILD = HERE / "_build"
REFDIR = HERE / "_ref"

OK, OUTPUT_FULL, INPUT_EXHAUSTED = 0, 1, 2
ERROR, EXCESS_BITS, INVALID_CONF, OOB = -1, -2, -3, -4

TEXT, RAND, ALPHA16, PERIODIC, BINARY, RUNS = range(6)


def build(ref: bool = True) -> None:
    """Compile the oracle (always) and the reference (only if /root/reference exists)."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", str(HERE)] + targets, check=True)


class OracleConf(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("window", "literal", "use_custom_dictionary", "extended", "dictionary_reset", "lazy_matching")]


_u8p = C.POINTER(C.c_uint8)
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = BUILD / "liboracle.so"
        if not path.exists():
            build(ref=False)
        L = C.CDLL(str(path))
        L.oracle_compress.restype = C.c_long
        L.oracle_compress.argtypes = [C.POINTER(OracleConf), C.c_char_p, C.c_char_p, C.c_size_t, C.c_char_p,
                                      C.c_size_t, C.c_int]
        L.oracle_decompress.restype = C.c_long
        L.oracle_decompress.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t,
                                        C.POINTER(C.c_int)]
        L.oracle_enc_new.restype = C.c_void_p
        L.oracle_enc_new.argtypes = [C.POINTER(OracleConf), C.c_char_p, C.POINTER(C.c_int)]
        L.oracle_enc_write.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.oracle_enc_flush.argtypes = [C.c_void_p, C.c_int]
        L.oracle_enc_size.restype = C.c_size_t
        L.oracle_enc_size.argtypes = [C.c_void_p]
        L.oracle_enc_data.restype = C.c_void_p
        L.oracle_enc_data.argtypes = [C.c_void_p]
        L.oracle_enc_window.restype = C.c_void_p
        L.oracle_enc_window.argtypes = [C.c_void_p]
        L.oracle_enc_free.argtypes = [C.c_void_p]
        L.oracle_initialize_dictionary.argtypes = [C.c_char_p, C.c_size_t, C.c_int]
        L.oracle_min_pattern_size.argtypes = [C.c_int, C.c_int]
        L.oracle_find_best_match.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int)]
        _lib = L
    return _lib


def _conf(window, literal, extended, dictionary, dictionary_reset, lazy_matching):
    return OracleConf(window, literal, int(dictionary is not None), int(bool(extended)), int(bool(dictionary_reset)),
                      int(bool(lazy_matching)))


class OracleError(Exception):
    def __init__(self, status):
        super().__init__(f"oracle status {status}")
        self.status = status


def initialize_dictionary(size: int, literal: int = 8) -> bytes:
    buf = C.create_string_buffer(size)
    lib().oracle_initialize_dictionary(buf, size, literal)
    return buf.raw


def min_pattern_size(window: int, literal: int) -> int:
    return lib().oracle_min_pattern_size(window, literal)


def find_best_match(window: bytes, pattern: bytes, max_len: int | None = None):
    idx = C.c_int(0)
    n = lib().oracle_find_best_match(bytes(window), len(window), bytes(pattern),
                                     len(pattern) if max_len is None else max_len, C.byref(idx))
    return idx.value, n


def compress(data: bytes, *, window=10, literal=8, extended=True, dictionary=None, dictionary_reset=False,
             lazy_matching=False, write_token=False) -> bytes:
    """init + compress_and_flush(write_token), whole buffer."""
    data = bytes(data)
    cf = _conf(window, literal, extended, dictionary, dictionary_reset, lazy_matching)
    cap = len(data) * 9 // 8 + 64
    out = C.create_string_buffer(cap)
    n = lib().oracle_compress(C.byref(cf), bytes(dictionary) if dictionary is not None else None, data, len(data),
                              out, cap, int(write_token))
    if n < 0:
        raise OracleError(n)
    return out.raw[:n]


def decompress(data: bytes, *, dictionary=None, window_bits_max=15, cap=None):
    """tamp_decompressor_init(NULL conf) + one decompress call.  Returns (bytes, status)."""
    data = bytes(data)
    if cap is None:
        cap = max(64, len(data) * 140)  # extended match: 134 bytes from ~3 bytes of input
    out = C.create_string_buffer(max(cap, 1))
    st = C.c_int(0)
    n = lib().oracle_decompress(bytes(dictionary) if dictionary is not None else None, window_bits_max, data,
                                len(data), out, cap, C.byref(st))
    return out.raw[:n], st.value


class Encoder:
    """Streaming oracle encoder: write()/flush() like tamp.Compressor, unbounded output."""

    def __init__(self, *, window=10, literal=8, extended=True, dictionary=None, dictionary_reset=False,
                 lazy_matching=False):
        cf = _conf(window, literal, extended, dictionary, dictionary_reset, lazy_matching)
        st = C.c_int(0)
        self._h = lib().oracle_enc_new(C.byref(cf), bytes(dictionary) if dictionary is not None else None,
                                       C.byref(st))
        if not self._h:
            raise OracleError(st.value)
        self._window = 1 << window

    def write(self, data: bytes) -> None:
        r = lib().oracle_enc_write(self._h, bytes(data), len(data))
        if r != OK:
            raise OracleError(r)

    def flush(self, write_token=True) -> None:
        r = lib().oracle_enc_flush(self._h, int(write_token))
        if r != OK:
            raise OracleError(r)

    def getvalue(self) -> bytes:
        n = lib().oracle_enc_size(self._h)
        return C.string_at(lib().oracle_enc_data(self._h), n) if n else b""

    def window(self) -> bytes:
        return C.string_at(lib().oracle_enc_window(self._h), self._window)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_enc_free(self._h)
            self._h = None


# --------------------------------------------------------------------------------------------
# The unmodified reference C, compiled from /root/reference into oracle/_ref (when available).
# --------------------------------------------------------------------------------------------

def pack_conf(window=10, literal=8, use_custom_dictionary=False, extended=False, dictionary_reset=False,
              append=False, lazy_matching=False) -> int:
    """TampConf bit-field image (common.h:170-182), GCC x86-64 LSB-first allocation."""
    return (window | literal << 4 | int(use_custom_dictionary) << 8 | int(extended) << 9 |
            int(dictionary_reset) << 10 | int(append) << 11 | int(lazy_matching) << 12)


class Ref:
    """Thin ctypes view of libtamp_ref.so (reference C API, compressor.h / decompressor.h)."""

    COMPRESSOR_SIZE = 48   # sizeof(TampCompressor), x86-64, measured
    DECOMPRESSOR_SIZE = 24

    def __init__(self, lazy=False):
        path = REFDIR / ("libtamp_ref_lazy.so" if lazy else "libtamp_ref.so")
        if not path.exists():
            raise FileNotFoundError(path)
        L = C.CDLL(str(path))
        sz, szp, vp, cp = C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p, C.c_char_p
        L.tamp_compressor_init.restype = C.c_int8
        L.tamp_compressor_init.argtypes = [vp, vp, vp]
        L.tamp_compressor_sink.restype = None
        L.tamp_compressor_sink.argtypes = [vp, cp, sz, szp]
        L.tamp_compressor_poll.restype = C.c_int8
        L.tamp_compressor_poll.argtypes = [vp, vp, sz, szp]
        L.tamp_compressor_full.restype = C.c_bool
        L.tamp_compressor_full.argtypes = [vp]
        L.tamp_compressor_flush.restype = C.c_int8
        L.tamp_compressor_flush.argtypes = [vp, vp, sz, szp, C.c_bool]
        L.tamp_compressor_compress_cb.restype = C.c_int8
        L.tamp_compressor_compress_cb.argtypes = [vp, vp, sz, szp, cp, sz, szp, vp, vp]
        L.tamp_compressor_compress_and_flush_cb.restype = C.c_int8
        L.tamp_compressor_compress_and_flush_cb.argtypes = [vp, vp, sz, szp, cp, sz, szp, C.c_bool, vp, vp]
        L.tamp_compressor_reset_dictionary.restype = C.c_int8
        L.tamp_compressor_reset_dictionary.argtypes = [vp, vp, sz, szp]
        L.tamp_decompressor_read_header.restype = C.c_int8
        L.tamp_decompressor_read_header.argtypes = [vp, cp, sz, szp]
        L.tamp_decompressor_init.restype = C.c_int8
        L.tamp_decompressor_init.argtypes = [vp, vp, vp, C.c_uint8]
        L.tamp_decompressor_decompress_cb.restype = C.c_int8
        L.tamp_decompressor_decompress_cb.argtypes = [vp, vp, sz, szp, cp, sz, szp, vp, vp]
        L.tamp_initialize_dictionary.restype = None
        L.tamp_initialize_dictionary.argtypes = [vp, sz, C.c_uint8]
        L.tamp_compute_min_pattern_size.restype = C.c_int8
        L.tamp_compute_min_pattern_size.argtypes = [C.c_uint8, C.c_uint8]
        self.L = L
        self.lazy = lazy

    # -- one-shot helpers -------------------------------------------------------------------
    def compress(self, data: bytes, *, window=10, literal=8, extended=True, dictionary=None,
                 dictionary_reset=False, lazy_matching=False, write_token=False) -> bytes:
        c = RefCompressor(self, window=window, literal=literal, extended=extended, dictionary=dictionary,
                          dictionary_reset=dictionary_reset, lazy_matching=lazy_matching)
        cap = len(data) * 9 // 8 + 64
        out, consumed, res = c.compress_and_flush(bytes(data), cap, write_token)
        assert res == OK and consumed == len(data), (res, consumed)
        return out

    def decompress(self, data: bytes, *, dictionary=None, window_bits_max=15, cap=None):
        if cap is None:
            cap = max(64, len(data) * 140)
        d = RefDecompressor(self, dictionary=dictionary, window_bits=window_bits_max)
        out, consumed, res = d.decompress(bytes(data), cap)
        return out, res


class RefCompressor:
    def __init__(self, ref: Ref, *, window=10, literal=8, extended=True, dictionary=None, dictionary_reset=False,
                 append=False, lazy_matching=False, default_conf=False):
        self.ref = ref
        self.state = C.create_string_buffer(Ref.COMPRESSOR_SIZE)
        self.window = C.create_string_buffer(1 << window)
        if dictionary is not None:
            assert len(dictionary) == 1 << window
            self.window.raw = bytes(dictionary)
        conf = C.c_uint16(pack_conf(window, literal, dictionary is not None, extended, dictionary_reset, append,
                                    lazy_matching and ref.lazy))
        self.init_res = ref.L.tamp_compressor_init(self.state, None if default_conf else C.byref(conf), self.window)

    def sink(self, data: bytes) -> int:
        n = C.c_size_t(0)
        self.ref.L.tamp_compressor_sink(self.state, data, len(data), C.byref(n))
        return n.value

    def full(self) -> bool:
        return bool(self.ref.L.tamp_compressor_full(self.state))

    def poll(self, cap: int):
        out = C.create_string_buffer(max(cap, 1))
        n = C.c_size_t(0)
        r = self.ref.L.tamp_compressor_poll(self.state, out, cap, C.byref(n))
        return out.raw[:n.value], r

    def flush(self, cap: int, write_token: bool):
        out = C.create_string_buffer(max(cap, 1))
        n = C.c_size_t(0)
        r = self.ref.L.tamp_compressor_flush(self.state, out, cap, C.byref(n), write_token)
        return out.raw[:n.value], r

    def compress(self, data: bytes, cap: int):
        out = C.create_string_buffer(max(cap, 1))
        n, m = C.c_size_t(0), C.c_size_t(0)
        r = self.ref.L.tamp_compressor_compress_cb(self.state, out, cap, C.byref(n), data, len(data), C.byref(m),
                                                   None, None)
        return out.raw[:n.value], m.value, r

    def compress_and_flush(self, data: bytes, cap: int, write_token: bool):
        out = C.create_string_buffer(max(cap, 1))
        n, m = C.c_size_t(0), C.c_size_t(0)
        r = self.ref.L.tamp_compressor_compress_and_flush_cb(self.state, out, cap, C.byref(n), data, len(data),
                                                             C.byref(m), write_token, None, None)
        return out.raw[:n.value], m.value, r

    def reset_dictionary(self, cap: int):
        out = C.create_string_buffer(max(cap, 1))
        n = C.c_size_t(0)
        r = self.ref.L.tamp_compressor_reset_dictionary(self.state, out, cap, C.byref(n))
        return out.raw[:n.value], r


class RefDecompressor:
    def __init__(self, ref: Ref, *, dictionary=None, window_bits=15, conf=None):
        self.ref = ref
        self.state = C.create_string_buffer(Ref.DECOMPRESSOR_SIZE)
        self.window = C.create_string_buffer(1 << window_bits)
        if dictionary is not None:
            self.window.raw = bytes(dictionary) + bytes((1 << window_bits) - len(dictionary))
        cptr = None
        if conf is not None:
            self._conf = C.c_uint16(conf)
            cptr = C.byref(self._conf)
        self.init_res = ref.L.tamp_decompressor_init(self.state, cptr, self.window, window_bits)

    def decompress(self, data: bytes, cap: int):
        out = C.create_string_buffer(max(cap, 1))
        n, m = C.c_size_t(0), C.c_size_t(0)
        r = self.ref.L.tamp_decompressor_decompress_cb(self.state, out, cap, C.byref(n), data, len(data),
                                                       C.byref(m), None, None)
        return out.raw[:n.value], m.value, r


def ref_available(lazy=False) -> bool:
    return (REFDIR / ("libtamp_ref_lazy.so" if lazy else "libtamp_ref.so")).exists()


# --------------------------------------------------------------------------------------------
# Threaded CPU batch harness (cpu_baseline + bulk expected-output generation for parity tests)
# --------------------------------------------------------------------------------------------

class Harness:
    def __init__(self, kind: str = "auto"):
        """kind: 'reference' (oracle/_ref), 'port' (oracle restatement) or 'auto' (reference if built)."""
        ref_path = REFDIR / "libharness_ref.so"
        port_path = BUILD / "libharness_port.so"
        if kind == "auto":
            kind = "reference" if ref_path.exists() else "port"
        if kind == "port" and not port_path.exists():
            build(ref=False)
        path = ref_path if kind == "reference" else port_path
        L = C.CDLL(str(path))
        vp, sz = C.c_void_p, C.c_size_t
        L.harness_compress.restype = C.c_double
        L.harness_compress.argtypes = [C.c_int, C.c_int, C.c_int, vp, sz, vp, sz, vp, sz, vp, vp, C.c_int]
        L.harness_decompress.restype = C.c_double
        L.harness_decompress.argtypes = [C.c_int, vp, sz, vp, sz, vp, sz, vp, vp, C.c_int]
        L.harness_generate.restype = C.c_double
        L.harness_generate.argtypes = [C.c_int, C.c_uint64, sz, sz, vp, C.c_int]
        L.harness_kind.restype = C.c_char_p
        self.L = L
        self.kind = L.harness_kind().decode()

    @staticmethod
    def _ptr(a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    def generate(self, kind: int, first_k: int, n_streams: int, stream_len: int, threads: int | None = None):
        out = np.empty((n_streams, stream_len), dtype=np.uint8)
        self.L.harness_generate(kind, first_k, n_streams, stream_len, self._ptr(out), threads or os.cpu_count())
        return out

    def compress(self, data: np.ndarray, *, window=10, literal=8, extended=False, sizes=None, out_stride=None,
                 threads: int | None = None):
        """data: (n_streams, stride) uint8.  Returns (out (n, out_stride), out_sizes, status, seconds)."""
        assert data.dtype == np.uint8 and data.ndim == 2 and data.flags.c_contiguous
        n, stride = data.shape
        if out_stride is None:
            out_stride = (stride * 9 + 7) // 8 + 32
        out = np.zeros((n, out_stride), dtype=np.uint8)
        osz = np.zeros(n, dtype=np.uint32)
        st = np.zeros(n, dtype=np.int8)
        if sizes is not None:
            sizes = np.ascontiguousarray(sizes, dtype=np.uint32)
        t = self.L.harness_compress(window, literal, int(extended), self._ptr(data), stride, self._ptr(sizes), n,
                                    self._ptr(out), out_stride, self._ptr(osz), self._ptr(st),
                                    threads or os.cpu_count())
        return out, osz, st, t

    def decompress(self, comp: np.ndarray, sizes: np.ndarray, out_stride: int, *, window_bits_max=15,
                   threads: int | None = None):
        assert comp.dtype == np.uint8 and comp.ndim == 2 and comp.flags.c_contiguous
        n, stride = comp.shape
        sizes = np.ascontiguousarray(sizes, dtype=np.uint32)
        out = np.zeros((n, out_stride), dtype=np.uint8)
        osz = np.zeros(n, dtype=np.uint32)
        st = np.zeros(n, dtype=np.int8)
        t = self.L.harness_decompress(window_bits_max, self._ptr(comp), stride, self._ptr(sizes), n, self._ptr(out),
                                      out_stride, self._ptr(osz), self._ptr(st), threads or os.cpu_count())
        return out, osz, st, t
