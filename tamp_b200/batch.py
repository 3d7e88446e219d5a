"""Batch (de)compression of independent Tamp streams on one B200 — the measured product path.

Thin torch plumbing over the C-ABI batch entry points (include/tamp_b200.h): tensors supply device
memory and the current CUDA stream; every byte of codec work happens in libtamp_b200.so's kernels.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import TampB200Batch
from .capi import make_conf


class TampError(RuntimeError):
    def __init__(self, status, what=""):
        super().__init__(f"tamp_b200 {what} failed: status {status}: {_lib.last_error()}")
        self.status = status


def compress_bound(n: int, literal: int = 8) -> int:
    conf = make_conf(10, literal)
    return int(_lib.lib().tamp_b200_compress_bound(C.byref(conf), n))


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_handle(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


@dataclass
class BatchResult:
    data: torch.Tensor     # (n_streams, out_stride) uint8
    sizes: torch.Tensor    # (n_streams,) int32 (bit pattern of uint32)
    status: torch.Tensor   # (n_streams,) int8 tamp_res per stream


def _check_2d(x):
    if x.dtype != torch.uint8 or x.dim() != 2 or not x.is_contiguous():
        raise ValueError("expected a contiguous (n_streams, stride) uint8 tensor")


def compress_batch(data: torch.Tensor, *, window=10, literal=8, extended=True, dictionary: torch.Tensor | None = None,
                   dictionary_reset=False, write_token=False, sizes: torch.Tensor | None = None,
                   out: torch.Tensor | None = None, out_stride: int | None = None,
                   lazy_matching: bool = False, append: bool = False) -> BatchResult:
    """Compress every row of ``data`` as an independent stream.

    Per stream the bytes equal ``tamp_compressor_init`` + ``tamp_compressor_compress_and_flush``
    of the reference C library.  ``data`` may live on the GPU (resident path, no copies, enqueued on
    the current stream) or in host memory (host entry point: H2D/D2H inside the call)."""
    _check_2d(data)
    n, stride = data.shape
    if out_stride is None:
        out_stride = out.shape[1] if out is not None else (compress_bound(stride, literal) + 15) // 16 * 16
    dev = data.device
    if out is None:
        out = torch.empty((n, out_stride), dtype=torch.uint8, device=dev, pin_memory=(dev.type == "cpu"))
    osz = torch.empty(n, dtype=torch.int32, device=dev)
    st = torch.empty(n, dtype=torch.int8, device=dev)
    if sizes is not None:
        sizes = sizes.to(device=dev, dtype=torch.int32).contiguous()
    # lazy matching (compressor.c:576-619) needs the TAMP_LAZY_MATCHING=1 flavour of the library (conf layout)
    conf = make_conf(window, literal, dictionary is not None, extended, dictionary_reset, append,
                     lazy_matching=True if lazy_matching else None)
    b = TampB200Batch(_ptr(data), None, _ptr(sizes), stride, _ptr(out), out_stride, _ptr(osz), _ptr(st), n)
    L = _lib.lib(lazy=bool(lazy_matching))
    if dev.type == "cuda":
        if dictionary is not None:
            dictionary = dictionary.to(dev).contiguous()
        with torch.cuda.device(dev):
            r = L.tamp_b200_compress_batch_device(C.byref(conf), _ptr(dictionary), C.byref(b), write_token,
                                                  _stream_handle(dev))
    else:
        if dictionary is not None:
            dictionary = dictionary.cpu().contiguous()
        r = L.tamp_b200_compress_batch(C.byref(conf), _ptr(dictionary), C.byref(b), write_token)
    if r != 0:
        raise TampError(r, "compress_batch")
    return BatchResult(out, osz, st)


def _window_bits_max(headers: torch.Tensor | None, dictionary: torch.Tensor | None, window_bits_max) -> int:
    """The ``window_bits_max`` of a decompress call (decompressor.h:67-79: the size of the window buffer the caller
    provides).  With a custom dictionary it is the dictionary's size, and the C entry point reads exactly
    ``1 << window_bits_max`` bytes of it.  ``None``: the largest window any frame header of the batch asks for (one
    small reduction + device->host read; pass the value to avoid it) — so that frames with windows <= 10 reach the
    specialised kernels instead of the general ones."""
    if dictionary is not None:
        n = dictionary.numel()
        if n < 256 or n > 32768 or n & (n - 1):
            raise ValueError("a custom dictionary holds 2**8 .. 2**15 bytes")
        bits = n.bit_length() - 1
        if window_bits_max is not None and window_bits_max != bits:
            if (1 << window_bits_max) > n:
                raise ValueError(f"window_bits_max={window_bits_max} needs a dictionary of {1 << window_bits_max} bytes, got {n}")
            bits = window_bits_max
        return bits
    if window_bits_max is not None:
        return int(window_bits_max)
    if headers is None or headers.numel() == 0:
        return 15
    return min(15, max(8, int((headers >> 5).max().item()) + 8))


def compress_batch_packed(data: torch.Tensor, *, window=10, literal=8, extended=True, dictionary: torch.Tensor | None = None,
                          dictionary_reset=False, write_token=False, sizes: torch.Tensor | None = None,
                          packed: torch.Tensor | None = None, offsets: torch.Tensor | None = None):
    """Host tensors only: compress the rows of ``data`` (pinned host memory) into contiguous frames in ``packed`` (pinned;
    allocated at the worst case if None).  Returns ``(packed, offsets, sizes, status)``: frame i is
    ``packed[offsets[i]:offsets[i] + sizes[i]]``; ``offsets`` has n + 1 int64 entries."""
    _check_2d(data)
    if data.device.type != "cpu":
        raise ValueError("compress_batch_packed takes host tensors; on the device use compress_batch + compact")
    n, stride = data.shape
    if packed is None:
        packed = torch.empty(n * compress_bound(stride, literal), dtype=torch.uint8, pin_memory=True)
    if offsets is None:
        offsets = torch.empty(n + 1, dtype=torch.int64)
    osz = torch.empty(n, dtype=torch.int32)
    st = torch.empty(n, dtype=torch.int8)
    if sizes is not None:
        sizes = sizes.to(dtype=torch.int32).contiguous()
    conf = make_conf(window, literal, dictionary is not None, extended, dictionary_reset)
    b = TampB200Batch(_ptr(data), None, _ptr(sizes), stride, None, 0, _ptr(osz), _ptr(st), n)
    if dictionary is not None:
        dictionary = dictionary.cpu().contiguous()
    r = _lib.lib().tamp_b200_compress_batch_packed(C.byref(conf), _ptr(dictionary), C.byref(b), write_token, _ptr(packed),
                                                   packed.numel(), _ptr(offsets))
    if r != 0:
        raise TampError(r, "compress_batch_packed")
    return packed, offsets, osz, st


def decompress_batch(comp: torch.Tensor, sizes: torch.Tensor | None, out_stride: int, *, window_bits_max=None,
                     dictionary: torch.Tensor | None = None, out: torch.Tensor | None = None) -> BatchResult:
    """Decompress every row of ``comp`` (``sizes[i]`` valid bytes each) into rows of ``out_stride`` bytes.

    Per stream: ``tamp_decompressor_init(conf=NULL)`` + one ``tamp_decompressor_decompress`` call with
    ``out_stride`` bytes of room; ``status`` carries that call's tamp_res (2 = input exhausted is the
    normal completion, 1 = output filled first).  ``window_bits_max``: see :func:`_window_bits_max`."""
    _check_2d(comp)
    n, stride = comp.shape
    dev = comp.device
    window_bits_max = _window_bits_max(comp[:, 0] if stride else None, dictionary, window_bits_max)
    if out is None:
        out = torch.empty((n, out_stride), dtype=torch.uint8, device=dev, pin_memory=(dev.type == "cpu"))
    osz = torch.empty(n, dtype=torch.int32, device=dev)
    st = torch.empty(n, dtype=torch.int8, device=dev)
    if sizes is not None:
        sizes = sizes.to(device=dev, dtype=torch.int32).contiguous()
    b = TampB200Batch(_ptr(comp), None, _ptr(sizes), stride, _ptr(out), out_stride, _ptr(osz), _ptr(st), n)
    L = _lib.lib()
    if dev.type == "cuda":
        if dictionary is not None:
            dictionary = dictionary.to(dev).contiguous()
        with torch.cuda.device(dev):
            r = L.tamp_b200_decompress_batch_device(_ptr(dictionary), window_bits_max, C.byref(b), _stream_handle(dev))
    else:
        if dictionary is not None:
            dictionary = dictionary.cpu().contiguous()
        r = L.tamp_b200_decompress_batch(_ptr(dictionary), window_bits_max, C.byref(b))
    if r != 0:
        raise TampError(r, "decompress_batch")
    return BatchResult(out, osz, st)


def compact(r: BatchResult, capacity: int | None = None):
    """Pack the fixed-stride rows of a compressed batch (device tensors) into contiguous frames.

    Returns ``(packed, offsets)``: frame i is ``packed[offsets[i]:offsets[i + 1]]``; ``offsets`` has n + 1 int64
    entries.  ``capacity`` defaults to the exact total (one device->host read of the size sum)."""
    dev = r.data.device
    if dev.type != "cuda":
        raise ValueError("compact() works on device-resident batches")
    n, stride = r.data.shape
    if capacity is None:
        capacity = int(r.sizes.to(torch.int64).sum().item())
    packed = torch.empty(max(capacity, 1), dtype=torch.uint8, device=dev)
    offsets = torch.empty(n + 1, dtype=torch.int64, device=dev)
    b = TampB200Batch(None, None, None, 0, _ptr(r.data), stride, _ptr(r.sizes), None, n)
    with torch.cuda.device(dev):
        rc = _lib.lib().tamp_b200_compact_batch_device(C.byref(b), _ptr(packed), capacity, _ptr(offsets),
                                                       _stream_handle(dev))
    if rc != 0:
        raise TampError(rc, "compact_batch")
    return packed[:capacity], offsets


def decompress_packed(packed: torch.Tensor, offsets: torch.Tensor, sizes: torch.Tensor, out_stride: int, *,
                      window_bits_max=None, dictionary: torch.Tensor | None = None, out: torch.Tensor | None = None) -> BatchResult:
    """Decompress contiguous frames (the output of :func:`compact`): frame i is ``sizes[i]`` bytes at
    ``packed[offsets[i]]``.  Device tensors, or host tensors (the output of :func:`compress_batch_packed`).
    ``window_bits_max``: see :func:`_window_bits_max`."""
    dev = packed.device
    n = sizes.numel()
    if dev.type == "cpu":
        if out is None:
            out = torch.empty((n, out_stride), dtype=torch.uint8, pin_memory=True)
        osz = torch.empty(n, dtype=torch.int32)
        st = torch.empty(n, dtype=torch.int8)
        sizes = sizes.to(dtype=torch.int32).contiguous()
        offsets = offsets.to(dtype=torch.int64).contiguous()
        if window_bits_max is None and dictionary is None and n:
            nz = offsets[:n][sizes > 0]
            window_bits_max = _window_bits_max(packed[nz] if nz.numel() else None, None, None)
        else:
            window_bits_max = _window_bits_max(None, dictionary, window_bits_max)
        b = TampB200Batch(_ptr(packed), _ptr(offsets), _ptr(sizes), 0, _ptr(out), out_stride, _ptr(osz), _ptr(st), n)
        if dictionary is not None:
            dictionary = dictionary.cpu().contiguous()
        rc = _lib.lib().tamp_b200_decompress_batch(_ptr(dictionary), window_bits_max, C.byref(b))
        if rc != 0:
            raise TampError(rc, "decompress_batch")
        return BatchResult(out, osz, st)
    if window_bits_max is None and dictionary is None and n:
        nz = offsets[:n].to(dev)[sizes.to(dev) > 0]
        window_bits_max = _window_bits_max(packed[nz] if nz.numel() else None, None, None)
    else:
        window_bits_max = _window_bits_max(None, dictionary, window_bits_max)
    if out is None:
        out = torch.empty((n, out_stride), dtype=torch.uint8, device=dev)
    osz = torch.empty(n, dtype=torch.int32, device=dev)
    st = torch.empty(n, dtype=torch.int8, device=dev)
    sizes = sizes.to(device=dev, dtype=torch.int32).contiguous()
    offsets = offsets.to(device=dev, dtype=torch.int64).contiguous()
    b = TampB200Batch(_ptr(packed), _ptr(offsets), _ptr(sizes), 0, _ptr(out), out_stride, _ptr(osz), _ptr(st), n)
    if dictionary is not None:
        dictionary = dictionary.to(dev).contiguous()
    with torch.cuda.device(dev):
        rc = _lib.lib().tamp_b200_decompress_batch_device(_ptr(dictionary), window_bits_max, C.byref(b),
                                                          _stream_handle(dev))
    if rc != 0:
        raise TampError(rc, "decompress_batch")
    return BatchResult(out, osz, st)


def compress_segmented(data, segment_size: int = 65536, *, window=10, literal=8, extended=True, out: torch.Tensor | None = None):
    """ONE long input as ONE Tamp stream, compressed segment-parallel.

    ``data``: bytes-like, or a 1-D uint8 tensor (on the GPU: resident path).  The input is cut into segments of
    ``segment_size`` bytes (a multiple of 16); the stream written is what ONE reference compressor with
    ``dictionary_reset=True`` writes when ``tamp_compressor_reset_dictionary()`` (compressor.c:847-881) is called
    between the segments and ``flush(write_token=True)`` at the end — any Tamp decompressor reads it front to back.
    Returns ``(stream, seg_offsets)``: bytes + a list of ints for bytes-like input, tensors (uint8, int64) on the
    input's device otherwise; ``seg_offsets`` has one entry per segment plus the total, the index
    ``decompress_segmented`` uses to work segment-parallel.  ``out``: a uint8 buffer on the input's device to write into
    (pinned for host input: a fresh pageable buffer costs more than the call), at least ``tamp_b200_segmented_bound`` bytes
    or ``TAMP_OUTPUT_FULL``."""
    as_bytes = not isinstance(data, torch.Tensor)
    t = torch.frombuffer(bytearray(data), dtype=torch.uint8) if as_bytes and len(data) else \
        (torch.empty(0, dtype=torch.uint8) if as_bytes else data)
    if t.dtype != torch.uint8 or t.dim() != 1 or not t.is_contiguous():
        raise ValueError("expected bytes or a contiguous 1-D uint8 tensor")
    n = t.numel()
    L = _lib.lib()
    conf = make_conf(window, literal, False, extended, True)
    nseg = int(L.tamp_b200_segment_count(n, segment_size))
    if nseg == 0:
        raise ValueError("segment_size must be a positive multiple of 16")
    cap = int(L.tamp_b200_segmented_bound(C.byref(conf), n, segment_size))
    dev = t.device
    if out is None:
        out = torch.empty(cap, dtype=torch.uint8, device=dev)
    elif out.dtype != torch.uint8 or out.dim() != 1 or not out.is_contiguous() or out.device != dev:
        raise ValueError("out: a contiguous 1-D uint8 tensor on the input's device")
    cap = out.numel()
    offs = torch.empty(nseg + 1, dtype=torch.int64, device=dev)
    total = C.c_uint64(0)
    if dev.type == "cuda":
        with torch.cuda.device(dev):
            r = L.tamp_b200_compress_segmented_device(C.byref(conf), _ptr(t), n, segment_size, _ptr(out), cap, _ptr(offs),
                                                      C.byref(total), _stream_handle(dev))
    else:
        r = L.tamp_b200_compress_segmented(C.byref(conf), _ptr(t), n, segment_size, _ptr(out), cap, _ptr(offs), C.byref(total))
    if r != 0:
        raise TampError(r, "compress_segmented")
    out = out[:total.value]
    if as_bytes:
        return out.numpy().tobytes(), offs.tolist()
    return out, offs


def decompress_segmented(stream, seg_offsets, segment_size: int, *, window_bits_max: int | None = None,
                         out_size: int | None = None):
    """Segment-parallel decompress of a stream written by ``compress_segmented`` (or by a reference compressor that
    reset its dictionary every ``segment_size`` input bytes), given the segment offsets.  ``out_size``: the room to
    provide (default: segments x segment_size).  Returns bytes for bytes-like input, a uint8 tensor otherwise."""
    as_bytes = not isinstance(stream, torch.Tensor)
    t = torch.frombuffer(bytearray(stream), dtype=torch.uint8) if as_bytes else stream
    dev = t.device
    offs = torch.as_tensor(seg_offsets, dtype=torch.int64).to(dev).contiguous()
    nseg = offs.numel() - 1
    if nseg < 1 or t.dtype != torch.uint8 or t.dim() != 1 or not t.is_contiguous():
        raise ValueError("expected a 1-D uint8 stream and at least two segment offsets")
    if window_bits_max is None:
        window_bits_max = int(t[int(offs[0])].item() >> 5) + 8 if t.numel() else 15
    cap = nseg * segment_size if out_size is None else out_size
    out = torch.empty(cap, dtype=torch.uint8, device=dev)
    total = C.c_uint64(0)
    L = _lib.lib()
    if dev.type == "cuda":
        with torch.cuda.device(dev):
            r = L.tamp_b200_decompress_segmented_device(_ptr(t), _ptr(offs), nseg, segment_size, window_bits_max, _ptr(out), cap,
                                                        C.byref(total), _stream_handle(dev))
    else:
        r = L.tamp_b200_decompress_segmented(_ptr(t), _ptr(offs), nseg, segment_size, window_bits_max, _ptr(out), cap,
                                             C.byref(total))
    if r != 0:
        raise TampError(r, "decompress_segmented")
    out = out[:total.value]
    return out.numpy().tobytes() if as_bytes else out


def synth(kind: int, first_k: int, n_streams: int, stream_len: int, device="cuda") -> torch.Tensor:
    """Deterministic synthetic streams generated on the device."""
    out = torch.empty((n_streams, stream_len), dtype=torch.uint8, device=device)
    with torch.cuda.device(out.device):
        r = _lib.lib().tamp_b200_synth_device(kind, first_k, n_streams, stream_len, _ptr(out),
                                              _stream_handle(out.device))
    if r != 0:
        raise TampError(r, "synth")
    return out


def set_kernel_mode(mode: int) -> None:
    """Test / benchmark hook: 0 = auto (specialised kernels where available), 1 = general kernels only, 2 = no
    position-parallel compressor (bitmap kernels), 4 = the position-parallel compressor without its lap variant
    (round-1 dispatch).  Applies to both library flavours."""
    _lib.lib().tamp_b200_set_kernel_mode(mode)
    if _lib.LIB_PATH_LAZY.exists():
        _lib.lib(lazy=True).tamp_b200_set_kernel_mode(mode)


def launch_count() -> int:
    return int(_lib.lib().tamp_b200_launch_count())


def copy_bytes() -> tuple[int, int]:
    """Cumulative (host->device, device->host) bytes moved by the host-pointer entry points."""
    a, b = C.c_uint64(0), C.c_uint64(0)
    _lib.lib().tamp_b200_copy_bytes(C.byref(a), C.byref(b))
    return int(a.value), int(b.value)
