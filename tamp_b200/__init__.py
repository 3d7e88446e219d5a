"""tamp_b200 — B200-native batch Tamp codec behind the reference's C API.

* ``tamp_b200.capi``  — object wrappers over the drop-in C API (tamp_compressor_* / tamp_decompressor_*).
* ``tamp_b200.batch`` — batch entry points on torch tensors (device-resident or host).
* ``compress`` / ``decompress`` — one-shot helpers with the reference Python package's signature
  (tamp/_c_compressor.pyx:189-199, tamp/_c_decompressor.pyx:184-187) running on the CUDA path.

All codec work runs in ``_build/libtamp_b200.so`` (sm_100a CUDA); there is no CPU fallback.
"""
from __future__ import annotations

from . import _lib
from ._lib import EXCESS_BITS, INPUT_EXHAUSTED, INVALID_CONF, OK, OOB, OUTPUT_FULL  # noqa: F401


class ExcessBitsError(Exception):
    """Provided data has more bits than expected ``literal`` bits."""


def compress(data: bytes, *, window=10, literal=8, dictionary=None, extended=True, dictionary_reset=False) -> bytes:
    from .capi import CCompressor
    if isinstance(data, str):
        data = data.encode()
    c = CCompressor(window=window, literal=literal, extended=extended, dictionary=dictionary,
                    dictionary_reset=dictionary_reset)
    if c.init_res != OK:
        raise ValueError(f"invalid configuration (status {c.init_res})")
    cap = len(data) * 9 // 8 + 64
    out, consumed, res = c.compress_and_flush(bytes(data), cap, False)
    if res == EXCESS_BITS:
        raise ExcessBitsError
    if res != OK or consumed != len(data):
        raise RuntimeError(f"compress failed: status {res}: {_lib.last_error()}")
    return out


def decompress(data: bytes, *, dictionary=None) -> bytearray:
    from .capi import CDecompressor
    data = bytes(data)
    d = CDecompressor(dictionary=dictionary, window_bits=15)
    out = bytearray()
    pos = 0
    while True:
        chunk, consumed, res = d.decompress(data[pos:], 1 << 20)
        out += chunk
        pos += consumed
        if res < 0:
            raise ValueError(f"decompress failed: status {res}")
        if res == INPUT_EXHAUSTED:
            return out
