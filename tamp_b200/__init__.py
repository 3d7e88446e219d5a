"""tamp_b200 — B200-native batch Tamp codec behind the reference's C API.

* ``tamp_b200.capi``  — object wrappers over the drop-in C API (tamp_compressor_* / tamp_decompressor_*).
* ``tamp_b200.batch`` — batch entry points on torch tensors (device-resident or host).
* ``tamp_b200.binding`` — ``Compressor`` / ``Decompressor`` / ``compress`` / ``decompress`` / ``open`` (+ ``compress_batch`` /
  ``decompress_batch`` over lists of bytes) with the
  reference Python package's interface (tamp/_c_compressor.pyx, tamp/_c_decompressor.pyx) on the CUDA path.

All codec work runs in ``_build/libtamp_b200.so`` (sm_100a CUDA); there is no CPU fallback.
"""
from __future__ import annotations

from . import _lib
from ._lib import EXCESS_BITS, INPUT_EXHAUSTED, INVALID_CONF, OK, OOB, OUTPUT_FULL  # noqa: F401


from .binding import (Compressor, Decompressor, ExcessBitsError, TextCompressor, TextDecompressor,  # noqa: E402,F401
                      compress, compress_batch, decompress, decompress_batch, open)
