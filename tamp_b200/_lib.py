"""Loader for the C-ABI shared library (tamp_b200/_build/libtamp_b200.so).

The library holds the host C API layer and the sm_100a CUDA kernels.  There is NO Python or CPU
implementation of the codec in this package: if the library is missing the import fails loudly
(build it with ``python -c 'import __graft_entry__ as g; g.build()'`` or ``make -C tamp_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
# (TAMP_B200_BUILD_DIR: A/B measurements of two builds of the library on one box, e.g. tools/sessions/r02_s32.sh)
_BUILD_DIR = PKG / os.environ.get("TAMP_B200_BUILD_DIR", "_build")
LIB_PATH = _BUILD_DIR / "libtamp_b200.so"
LIB_PATH_LAZY = _BUILD_DIR / "libtamp_b200_lazy.so"  # same kernels, TAMP_LAZY_MATCHING=1 struct layouts

OK, OUTPUT_FULL, INPUT_EXHAUSTED = 0, 1, 2
ERROR, EXCESS_BITS, INVALID_CONF, OOB = -1, -2, -3, -4
IO_ERROR, READ_ERROR, WRITE_ERROR = -10, -11, -12


class TampConf(C.Structure):
    """include/tamp/common.h TampConf (reference common.h:170-182)."""
    _fields_ = [("window", C.c_uint16, 4), ("literal", C.c_uint16, 4), ("use_custom_dictionary", C.c_uint16, 1),
                ("extended", C.c_uint16, 1), ("dictionary_reset", C.c_uint16, 1), ("append", C.c_uint16, 1)]


class TampCompressor(C.Structure):
    """include/tamp/compressor.h TampCompressor (reference compressor.h:13-66), 48 bytes."""
    _fields_ = [("window", C.c_void_p), ("bit_buffer", C.c_uint32), ("window_pos", C.c_uint16),
                ("bit_buffer_pos", C.c_uint8), ("input_size", C.c_uint8), ("input_pos", C.c_uint8),
                ("input", C.c_uint8 * 16), ("min_pattern_size", C.c_uint8), ("conf", TampConf),
                ("extended_match_position", C.c_uint16), ("rle_count", C.c_uint8),
                ("extended_match_count", C.c_uint8), ("last_was_flush", C.c_uint8)]


class TampConfLazy(C.Structure):
    """TampConf when the library and its caller are built with TAMP_LAZY_MATCHING=1 (common.h:178-181)."""
    _fields_ = [("window", C.c_uint16, 4), ("literal", C.c_uint16, 4), ("use_custom_dictionary", C.c_uint16, 1),
                ("extended", C.c_uint16, 1), ("dictionary_reset", C.c_uint16, 1), ("append", C.c_uint16, 1),
                ("lazy_matching", C.c_uint16, 1)]


class TampCompressorLazy(C.Structure):
    """TampCompressor with the lazy-matching cache fields (compressor.h:45-53); still 48 bytes."""
    _fields_ = [("window", C.c_void_p), ("bit_buffer", C.c_uint32), ("window_pos", C.c_uint16),
                ("bit_buffer_pos", C.c_uint8), ("input_size", C.c_uint8), ("input_pos", C.c_uint8),
                ("input", C.c_uint8 * 16), ("min_pattern_size", C.c_uint8), ("conf", TampConfLazy),
                ("cached_match_index", C.c_int16), ("extended_match_position", C.c_uint16),
                ("cached_match_size", C.c_uint8), ("rle_count", C.c_uint8), ("extended_match_count", C.c_uint8),
                ("last_was_flush", C.c_uint8)]


class TampDecompressor(C.Structure):
    """include/tamp/decompressor.h TampDecompressor (reference decompressor.h:13-57), 24 bytes."""
    _fields_ = [("window", C.c_void_p), ("bit_buffer", C.c_uint32), ("window_pos", C.c_uint16),
                ("bit_buffer_pos", C.c_uint8), ("token_state", C.c_uint8), ("pending_window_offset", C.c_uint16),
                ("pending_match_size", C.c_uint16), ("conf_window", C.c_uint8, 4), ("conf_literal", C.c_uint8, 4),
                ("min_pattern_size", C.c_uint8, 2), ("conf_extended", C.c_uint8, 1),
                ("conf_dictionary_reset", C.c_uint8, 1), ("skip_bytes", C.c_uint8),
                ("window_bits_max", C.c_uint8, 4), ("configured", C.c_uint8, 1), ("header_bytes_read", C.c_uint8, 2),
                ("last_was_flush", C.c_uint8, 1)]


class TampB200Batch(C.Structure):
    """include/tamp_b200.h TampB200Batch."""
    _fields_ = [("in_", C.c_void_p), ("in_offsets", C.c_void_p), ("in_sizes", C.c_void_p), ("in_stride", C.c_uint64),
                ("out", C.c_void_p), ("out_stride", C.c_uint64), ("out_sizes", C.c_void_p), ("status", C.c_void_p),
                ("n_streams", C.c_uint64)]


assert C.sizeof(TampConf) == 2 and C.sizeof(TampCompressor) == 48 and C.sizeof(TampDecompressor) == 24
assert C.sizeof(TampConfLazy) == 2 and C.sizeof(TampCompressorLazy) == 48

# Every symbol include/*.h declares (checked by tests/test_abi.py).
READ_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_ubyte), C.c_size_t)   # tamp_read_t
WRITE_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_ubyte), C.c_size_t)  # tamp_write_t
PROGRESS_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_size_t, C.c_size_t)         # tamp_callback_t


class TampMemReader(C.Structure):
    _fields_ = [("data", C.c_void_p), ("size", C.c_size_t), ("pos", C.c_size_t)]


class TampMemWriter(C.Structure):
    _fields_ = [("data", C.c_void_p), ("capacity", C.c_size_t), ("pos", C.c_size_t)]


EXPORTS = [
    "tamp_initialize_dictionary", "tamp_compute_min_pattern_size", "tamp_window_copy",
    "tamp_compressor_init", "tamp_compressor_sink", "tamp_compressor_poll", "tamp_compressor_full",
    "tamp_compressor_flush", "tamp_compressor_reset_dictionary", "tamp_compressor_compress_cb",
    "tamp_compressor_compress_and_flush_cb",
    "tamp_decompressor_read_header", "tamp_decompressor_init", "tamp_decompressor_decompress_cb",
    "tamp_compress_stream", "tamp_decompress_stream",
    "tamp_stream_mem_read", "tamp_stream_mem_write", "tamp_stream_stdio_read", "tamp_stream_stdio_write",
    "tamp_b200_compress_bound", "tamp_b200_compress_batch", "tamp_b200_decompress_batch", "tamp_b200_compress_batch_packed",
    "tamp_b200_compress_batch_device", "tamp_b200_decompress_batch_device", "tamp_b200_compact_batch_device",
    "tamp_b200_set_kernel_mode",
    "tamp_b200_segment_count", "tamp_b200_segmented_bound", "tamp_b200_compress_segmented", "tamp_b200_compress_segmented_device",
    "tamp_b200_decompress_segmented", "tamp_b200_decompress_segmented_device",
    "tamp_b200_synth_device", "tamp_b200_device_count", "tamp_b200_set_device", "tamp_b200_last_error",
    "tamp_b200_launch_count", "tamp_b200_copy_bytes", "tamp_b200_version",
]

_lib = None
_lib_lazy = None


def build() -> None:
    subprocess.run(["make", "-s", "-j8", "-C", str(PKG / "csrc")], check=True)


def lib(lazy: bool = False) -> C.CDLL:
    global _lib, _lib_lazy
    if lazy and _lib_lazy is not None:
        return _lib_lazy
    if not lazy and _lib is not None:
        return _lib
    path = LIB_PATH_LAZY if lazy else LIB_PATH
    if not path.exists():
        raise ImportError(f"{path} is missing: the CUDA extension is not built and tamp_b200 has no fallback "
                          f"(run `make -C {PKG / 'csrc'}`)")
    L = C.CDLL(str(path))
    vp, sz, szp, cp, u8 = C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_char_p, C.c_uint8
    i8 = C.c_int8
    sig = {
        "tamp_initialize_dictionary": (None, [vp, sz, u8]),
        "tamp_compute_min_pattern_size": (i8, [u8, u8]),
        "tamp_window_copy": (None, [vp, C.POINTER(C.c_uint16), C.c_uint16, u8, C.c_uint16]),
        "tamp_compressor_init": (i8, [vp, vp, vp]),
        "tamp_compressor_sink": (None, [vp, cp, sz, szp]),
        "tamp_compressor_poll": (i8, [vp, vp, sz, szp]),
        "tamp_compressor_full": (C.c_bool, [vp]),
        "tamp_compressor_flush": (i8, [vp, vp, sz, szp, C.c_bool]),
        "tamp_compressor_reset_dictionary": (i8, [vp, vp, sz, szp]),
        "tamp_compressor_compress_cb": (i8, [vp, vp, sz, szp, cp, sz, szp, vp, vp]),
        "tamp_compressor_compress_and_flush_cb": (i8, [vp, vp, sz, szp, cp, sz, szp, C.c_bool, vp, vp]),
        "tamp_decompressor_read_header": (i8, [vp, cp, sz, szp]),
        "tamp_decompressor_init": (i8, [vp, vp, vp, u8]),
        "tamp_decompressor_decompress_cb": (i8, [vp, vp, sz, szp, cp, sz, szp, vp, vp]),
        "tamp_compress_stream": (i8, [vp, vp, vp, vp, vp, szp, szp, vp, vp]),
        "tamp_decompress_stream": (i8, [vp, vp, vp, vp, vp, szp, szp, vp, vp]),
        "tamp_stream_mem_read": (C.c_int, [vp, vp, sz]),
        "tamp_stream_mem_write": (C.c_int, [vp, vp, sz]),
        "tamp_stream_stdio_read": (C.c_int, [vp, vp, sz]),
        "tamp_stream_stdio_write": (C.c_int, [vp, vp, sz]),
        "tamp_b200_compress_bound": (sz, [vp, sz]),
        "tamp_b200_compress_batch": (i8, [vp, vp, vp, C.c_bool]),
        "tamp_b200_decompress_batch": (i8, [vp, u8, vp]),
        "tamp_b200_compress_batch_packed": (i8, [vp, vp, vp, C.c_bool, vp, C.c_uint64, vp]),
        "tamp_b200_compress_batch_device": (i8, [vp, vp, vp, C.c_bool, vp]),
        "tamp_b200_decompress_batch_device": (i8, [vp, u8, vp, vp]),
        "tamp_b200_set_kernel_mode": (None, [C.c_int]),
        "tamp_b200_synth_device": (i8, [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, vp, vp]),
        "tamp_b200_compact_batch_device": (i8, [vp, vp, C.c_uint64, vp, vp]),
        "tamp_b200_segment_count": (C.c_uint64, [C.c_uint64, C.c_uint64]),
        "tamp_b200_segmented_bound": (C.c_uint64, [vp, C.c_uint64, C.c_uint64]),
        "tamp_b200_compress_segmented": (i8, [vp, vp, C.c_uint64, C.c_uint64, vp, C.c_uint64, vp, C.POINTER(C.c_uint64)]),
        "tamp_b200_compress_segmented_device": (i8, [vp, vp, C.c_uint64, C.c_uint64, vp, C.c_uint64, vp,
                                                    C.POINTER(C.c_uint64), vp]),
        "tamp_b200_decompress_segmented": (i8, [vp, vp, C.c_uint64, C.c_uint64, u8, vp, C.c_uint64, C.POINTER(C.c_uint64)]),
        "tamp_b200_decompress_segmented_device": (i8, [vp, vp, C.c_uint64, C.c_uint64, u8, vp, C.c_uint64,
                                                      C.POINTER(C.c_uint64), vp]),
        "tamp_b200_device_count": (C.c_int, []),
        "tamp_b200_set_device": (i8, [C.c_int]),
        "tamp_b200_last_error": (C.c_char_p, []),
        "tamp_b200_launch_count": (C.c_uint64, []),
        "tamp_b200_copy_bytes": (None, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
        "tamp_b200_version": (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here == a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lazy:
        _lib_lazy = L
    else:
        _lib = L
    return L


def last_error() -> str:
    return lib().tamp_b200_last_error().decode()
