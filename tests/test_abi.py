"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/*.h declares, keeps the reference's struct layouts, and its host-side functions (init, sink,
full, read_header, dictionary seeding, window copy) behave like the reference's.  No kernel is
launched here."""
import ctypes as C
import hashlib
import random
import re
import subprocess
from pathlib import Path

import pytest

import oracle
from tamp_b200 import _lib
from tamp_b200.capi import CCompressor, CDecompressor, make_conf, read_header

ROOT = Path(__file__).resolve().parents[1]
REF_INC = Path("/root/reference/tamp/_c_src")


def declared_symbols():
    names = set()
    for h in list((ROOT / "include").glob("*.h")) + list((ROOT / "include" / "tamp").glob("*.h")):
        text = re.sub(r"/\*.*?\*/", "", h.read_text(), flags=re.S)
        text = re.sub(r"static inline[^{;]*\{.*?\n\}", "", text, flags=re.S)
        for m in re.finditer(r"\b(tamp_[a-z0-9_]+)\s*\(", text):
            names.add(m.group(1))
    names.discard("tamp_compressor_compress_poll")  # macro alias
    names.discard("tamp_callback_t")
    return names


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    declared = declared_symbols()
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/ but not exported"
    assert declared <= set(_lib.EXPORTS) | {"tamp_compressor_compress", "tamp_compressor_compress_and_flush",
                                             "tamp_decompressor_decompress"}
    assert b"sm_100a" in L.tamp_b200_version()


def test_library_contains_sm100a_code():
    out = subprocess.run(["cuobjdump", "-lelf", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_struct_sizes():
    assert C.sizeof(_lib.TampConf) == 2
    assert C.sizeof(_lib.TampCompressor) == 48
    assert C.sizeof(_lib.TampDecompressor) == 24


LAYOUT_PROBE = r"""
#include <stddef.h>
#include <stdio.h>
#include "tamp/compressor.h"
#include "tamp/decompressor.h"
#define O(t, f) printf(#t "." #f " %zu\n", offsetof(t, f))
int main(void) {
    printf("sizeof TampConf %zu\nsizeof TampCompressor %zu\nsizeof TampDecompressor %zu\n", sizeof(TampConf),
           sizeof(TampCompressor), sizeof(TampDecompressor));
    O(TampCompressor, window); O(TampCompressor, bit_buffer); O(TampCompressor, window_pos);
    O(TampCompressor, bit_buffer_pos); O(TampCompressor, input_size); O(TampCompressor, input_pos);
    O(TampCompressor, input); O(TampCompressor, min_pattern_size); O(TampCompressor, conf);
    O(TampCompressor, extended_match_position); O(TampCompressor, rle_count);
    O(TampCompressor, extended_match_count); O(TampCompressor, last_was_flush);
    O(TampDecompressor, window); O(TampDecompressor, bit_buffer); O(TampDecompressor, window_pos);
    O(TampDecompressor, pos_and_state); O(TampDecompressor, pending_window_offset);
    O(TampDecompressor, pending_match_size); O(TampDecompressor, skip_bytes);
    TampConf c = {0}; c.window = 15; c.literal = 8; c.extended = 1; c.dictionary_reset = 1;
    printf("conf bits %04x\n", *(unsigned short *)&c);
    TampDecompressor d; unsigned char *p = (unsigned char *)&d;
    for (unsigned i = 0; i < sizeof d; i++) p[i] = 0;
    d.conf_window = 15; d.conf_literal = 5; d.min_pattern_size = 3; d.conf_extended = 1; d.window_bits_max = 9;
    d.configured = 1; d.header_bytes_read = 2; d.last_was_flush = 1;
    for (unsigned i = 20; i < 24; i++) printf("%02x", p[i]);
    printf("\n%d %d %d %d\n", TAMP_OK, TAMP_OUTPUT_FULL, TAMP_INPUT_EXHAUSTED, TAMP_OOB);
    return 0;
}
"""


@pytest.mark.skipif(not REF_INC.exists(), reason="reference headers not present on this box")
@pytest.mark.parametrize("lazy", [0, 1])
def test_header_layout_matches_reference(tmp_path, lazy):
    """Compile the same offsetof probe against our headers and the reference's; outputs must match."""
    src = tmp_path / "probe.c"
    src.write_text(LAYOUT_PROBE)
    outs = []
    for inc in (ROOT / "include", REF_INC):
        exe = tmp_path / f"probe_{inc.name}"
        subprocess.run(["gcc", f"-DTAMP_LAZY_MATCHING={lazy}", "-I", str(inc), str(src), "-o", str(exe)], check=True)
        outs.append(subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout)
    assert outs[0] == outs[1]


def test_dictionary_seed_kat(kats):
    L = _lib.lib()
    for k in kats:
        if k["kind"] == "dictionary":
            buf = C.create_string_buffer(256)
            L.tamp_initialize_dictionary(buf, 256, k["literal"])
            assert buf.raw.hex() == k["expected"]
    for lit in (5, 6, 7, 8):
        for size in (256, 1000, 32768):
            buf = C.create_string_buffer(size)
            L.tamp_initialize_dictionary(buf, size, lit)
            assert buf.raw == oracle.initialize_dictionary(size, lit)


def test_min_pattern_size_and_window_copy():
    L = _lib.lib()
    for w in range(8, 16):
        for lit in range(5, 9):
            assert L.tamp_compute_min_pattern_size(w, lit) == oracle.min_pattern_size(w, lit)
    rng = random.Random(5)
    for _ in range(2000):
        bits = rng.choice([8, 10])
        W = 1 << bits
        win = bytearray(rng.randrange(256) for _ in range(W))
        n = rng.randrange(0, 135)
        src = rng.randrange(0, W - n)
        pos = rng.choice([rng.randrange(W), (src + rng.randrange(-3, n + 3)) % W])
        # the reference's direction rule (common.c:62-79): back-to-front iff 0 < (pos - src) & mask < n
        expect = bytearray(win)
        gap = (pos - src) % W
        order = range(n - 1, -1, -1) if 0 < gap < n else range(n)
        for i in order:
            expect[(pos + i) % W] = expect[src + i]
        if n <= 16 or pos + n <= W:  # every call the codec can make: equals "snapshot, then write"
            snap = bytes(win[src:src + n])
            chk = bytearray(win)
            for i, b in enumerate(snap):
                chk[(pos + i) % W] = b
            assert chk == expect
        buf = (C.c_char * W).from_buffer(win)
        wp = C.c_uint16(pos)
        L.tamp_window_copy(buf, C.byref(wp), src, n, W - 1)
        assert bytes(win) == bytes(expect) and wp.value == (pos + n) % W


def test_compressor_init_sink_full_host_side():
    """init/sink/full run on the host (compressor.c:191-245, :665-679, :77-79): state bytes must equal
    the oracle's view: header queued in the bit buffer, seeded window, ring bookkeeping."""
    for w, lit, ext, dr in [(10, 8, True, False), (8, 5, False, False), (15, 7, True, True), (12, 6, True, False)]:
        c = CCompressor(window=w, literal=lit, extended=ext, dictionary_reset=dr)
        assert c.init_res == 0
        header = ((w - 8) << 5) | ((lit - 5) << 3) | (int(ext) << 1) | int(dr)
        assert c.state.bit_buffer == header << 24 and c.state.bit_buffer_pos == (16 if dr else 8)
        assert c.state.min_pattern_size == oracle.min_pattern_size(w, lit)
        assert c.window.raw == oracle.initialize_dictionary(1 << w, lit if ext else 8)
        assert c.sink(b"0123456789") == 10 and not c.full()
        assert c.sink(b"abcdefghij") == 6 and c.full()
        assert c.sink(b"zz") == 0
        assert bytes(c.state.input) == b"0123456789abcdef"
    assert CCompressor(window=10, default_conf=True).state.conf.extended == 1   # conf == NULL => v2
    # invalid configurations (compressor.c:207-209): append needs dictionary_reset and no custom dictionary
    L = _lib.lib()
    st, win = _lib.TampCompressor(), C.create_string_buffer(1 << 15)
    for conf in (make_conf(7 & 0xF, 8), make_conf(10, 4), make_conf(10, 9 & 0xF), make_conf(10, 8, append=True),
                 make_conf(10, 8, use_custom_dictionary=True, dictionary_reset=True, append=True)):
        assert L.tamp_compressor_init(C.byref(st), C.byref(conf), win) == _lib.INVALID_CONF
    ok = make_conf(10, 8, dictionary_reset=True, append=True)
    assert L.tamp_compressor_init(C.byref(st), C.byref(ok), win) == 0
    assert st.bit_buffer == (0xAB << 23) and st.bit_buffer_pos == 16 and st.last_was_flush == 1


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built")
def test_host_functions_vs_reference_state_bytes():
    ref = oracle.Ref()
    rng = random.Random(3)
    for _ in range(100):
        w, lit = rng.randrange(8, 16), rng.randrange(5, 9)
        ext, dr = rng.random() < .5, rng.random() < .5
        app = dr and rng.random() < .3
        a = CCompressor(window=w, literal=lit, extended=ext, dictionary_reset=dr, append=app)
        b = oracle.RefCompressor(ref, window=w, literal=lit, extended=ext, dictionary_reset=dr, append=app)
        assert a.init_res == b.init_res == 0
        for _ in range(5):
            chunk = bytes(rng.randrange(256) for _ in range(rng.randrange(0, 12)))
            assert a.sink(chunk) == b.sink(chunk)
            assert a.full() == b.full()
        assert a.state_bytes() == b.state.raw[8:]
        assert hashlib.sha256(a.window.raw).digest() == hashlib.sha256(b.window.raw).digest()


def test_read_header_and_decompressor_init():
    conf, used, r = read_header(bytes([0x5A]))
    assert (r, used, conf.window, conf.literal, conf.extended, conf.use_custom_dictionary) == (0, 1, 10, 8, 1, 0)
    conf, used, r = read_header(bytes([0xFD, 0x00]))
    assert (r, used, conf.window, conf.literal, conf.dictionary_reset, conf.use_custom_dictionary) == (0, 2, 15, 8, 1, 1)
    assert read_header(b"")[2] == _lib.INPUT_EXHAUSTED
    assert read_header(bytes([0x59]))[2] == _lib.INPUT_EXHAUSTED        # second byte missing
    assert read_header(bytes([0x59, 0x01]))[2] == _lib.INVALID_CONF     # reserved bits set (decompressor.c:285)
    assert CDecompressor(window_bits=7).init_res == _lib.INVALID_CONF
    d = CDecompressor(window_bits=10)
    assert d.init_res == 0 and d.state.configured == 0 and d.state.window_bits_max == 10
    d = CDecompressor(window_bits=10, conf=make_conf(12, 8))
    assert d.init_res == _lib.INVALID_CONF                                # window > buffer (decompressor.c:311)
    d = CDecompressor(window_bits=12, conf=make_conf(12, 7, extended=True))
    assert d.init_res == 0 and d.state.configured == 1 and d.state.min_pattern_size == 2
    assert d.window.raw == oracle.initialize_dictionary(1 << 12, 7)


def test_codec_calls_fail_loudly_without_gpu():
    """No CPU fallback: on a box without CUDA the codec entry points return TAMP_ERROR with a message."""
    L = _lib.lib()
    if L.tamp_b200_device_count() > 0:
        pytest.skip("CUDA device present")
    c = CCompressor(window=10)
    c.sink(b"0123456789abcdef")
    out, res = c.poll(16)
    assert res == _lib.ERROR and out == b"" and "no CUDA device" in _lib.last_error()
    d = CDecompressor(window_bits=10)
    assert d.decompress(bytes.fromhex("58b3041c8100030000"), 64)[2] == _lib.ERROR
    # batch entry points: host-pointer compress and the device-side compaction
    import ctypes as C
    from tamp_b200.capi import make_conf
    n = 4
    inp = (C.c_ubyte * (n * 64))()
    out = (C.c_ubyte * (n * 96))()
    osz = (C.c_uint32 * n)()
    b = _lib.TampB200Batch(C.cast(inp, C.c_void_p), None, None, 64, C.cast(out, C.c_void_p), 96,
                           C.cast(osz, C.c_void_p), None, n)
    conf = make_conf(10, 8)
    assert L.tamp_b200_compress_batch(C.byref(conf), None, C.byref(b), False) == _lib.ERROR
    offs = (C.c_uint64 * (n + 1))()
    assert L.tamp_b200_compact_batch_device(C.byref(b), C.cast(out, C.c_void_p), 0, C.cast(offs, C.c_void_p), None) == _lib.ERROR
    assert "no CUDA device" in _lib.last_error()


def test_traffic_capture_is_current():
    """profiles/traffic.json (roofline.traffic of bench.py) must come from an ncu capture of the kernel sources as they
    are now: profiles/make_traffic.py stores their hash beside the DRAM bytes."""
    import json
    import sys
    root = Path(__file__).resolve().parents[1]
    sys.path.insert(0, str(root / "profiles"))
    import make_traffic
    tj = json.loads((root / "profiles" / "traffic.json").read_text())
    for kind in ("compress", "decompress"):
        assert tj[kind]["sources_sha256"] == make_traffic.source_hash(kind), \
            f"{kind}: kernel source changed since the capture; re-run ncu and profiles/make_traffic.py"
        assert tj[kind]["dram_bytes_per_launch"] > 0
