"""The CUDA kernel SOURCES stepped on the CPU (tests/emu: a fiber-per-thread SIMT emulator, g++) against the oracle.

What this adds to the GPU parity tests: the kernel's logic — shared-memory regions that alias each other over the
phases, queue indices, token lists, bit packing — is checked wherever the CPU suite runs, and with the lanes of a
warp deliberately NOT in lock step between collectives (seeded shuffles), so a missing ``__syncwarp`` shows up.
The emulator is test infrastructure: nothing in ``tamp_b200`` can reach it, and it says nothing about speed.
"""
import ctypes as C
import random
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle
from conftest import gen_stream

HERE = Path(__file__).resolve().parent
EMU = HERE / "emu"
CUDA_INC = Path("/usr/local/cuda/include")

pytestmark = pytest.mark.timeout(900, method="thread")  # a broken kernel may spin for ever inside the C call

F_EXTENDED, F_DICT_RESET, F_LAZY, F_CUSTOM, F_APPEND, F_APPEND_TAIL = 1, 2, 4, 8, 16, 32


def _append_flags(append):
    """append = 1: every stream starts in append mode; 2: every stream behind the first (the segments of one stream)."""
    return {0: 0, 1: F_APPEND | F_DICT_RESET, 2: F_APPEND | F_APPEND_TAIL | F_DICT_RESET}[append]
DEFERRED = 0xFFFFFFFF


@pytest.fixture(scope="module")
def emu():
    if shutil.which("g++") is None or not (CUDA_INC / "cuda_runtime.h").exists():
        pytest.skip("needs g++ and the CUDA headers")
    out = EMU / "_build" / "libemu_kernels.so"
    out.parent.mkdir(exist_ok=True)
    units = sorted(EMU.glob("emu_*.cpp"))
    srcs = units + [EMU / "cuda_emu.h"] + sorted((HERE.parent / "tamp_b200" / "csrc").rglob("*.cu*"))
    import fcntl
    with open(out.parent / ".lock", "w") as lock:  # (pytest-xdist: one worker builds, the others wait and find it built)
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not out.exists() or out.stat().st_mtime < max(s.stat().st_mtime for s in srcs):
            flags = ["-O1", "-g", "-std=c++17", "-fPIC", "-Wno-attributes", f"-I{CUDA_INC}"]
            objs = [out.parent / (u.stem + ".o") for u in units]
            procs = [subprocess.Popen(["g++"] + flags + ["-c", str(u), "-o", str(o)]) for u, o in zip(units, objs)]  # in parallel
            assert all(p.wait() == 0 for p in procs), "g++ failed on an emulator unit"
            tmp = out.with_suffix(".so.tmp")
            subprocess.run(["g++", "-shared"] + [str(o) for o in objs] + ["-o", str(tmp)], check=True)
            tmp.replace(out)
    lib = C.CDLL(str(out))
    lib.emu_fast_compress.restype = C.c_int
    lib.emu_fast_compress.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint,
                                      C.c_uint64]
    lib.emu_wide_compress.restype = C.c_int
    lib.emu_wide_compress.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64,
                                      C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint, C.c_uint64,
                                      C.c_int]
    lib.emu_compact.restype = C.c_uint64
    lib.emu_compact.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
    lib.emu_wide_decompress.restype = None
    lib.emu_wide_decompress.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                        C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint, C.c_int,
                                        C.c_uint64]
    lib.emu_generic_compress.restype = None
    lib.emu_generic_compress.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64,
                                         C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_uint64]
    lib.emu_generic_decompress.restype = None
    lib.emu_generic_decompress.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                           C.c_uint64, C.c_uint64]
    lib.emu_set_seg_header.restype = None
    lib.emu_set_seg_header.argtypes = [C.c_uint32]
    lib.emu_synth.restype = None
    lib.emu_synth.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
    lib.emu_fast_decompress.restype = None
    lib.emu_fast_decompress.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint,
                                        C.c_uint64]
    lib.emu_ppar_compress.restype = C.c_int
    lib.emu_ppar_compress.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                      C.c_uint64, C.c_uint, C.c_uint64]
    lib.emu_fast_decompress_pickup.restype = None
    lib.emu_fast_decompress_pickup.argtypes = lib.emu_fast_decompress.argtypes
    lib.emu_split_decompress.restype = C.c_int
    lib.emu_split_decompress.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                         C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint, C.c_uint64]
    lib.emu_walk_compress.restype = C.c_int
    lib.emu_walk_compress.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint,
                                      C.c_uint64]
    lib.emu_hwalk_compress.restype = C.c_int
    lib.emu_hwalk_compress.argtypes = [C.c_void_p] + [C.c_int] * 10 + [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                                                      C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint, C.c_uint64]
    lib.emu_cwalk_compress.restype = C.c_int
    lib.emu_cwalk_compress.argtypes = [C.c_void_p] + [C.c_int] * 9 + [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                                                     C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint, C.c_uint64]
    lib.emu_lsplit_decompress.restype = C.c_int
    lib.emu_lsplit_decompress.argtypes = lib.emu_split_decompress.argtypes
    return lib


WALK = 100  # pseudo-mode of ppar(): k_walk_compress<v1> (segment-walk compressor) instead of k_ppar_compress<mode>
WALK_EXT = 101  # k_walk_compress<extended format>


def ppar(lib, mode, streams, *, window, literal=8, dictionary=None, dict_reset=False, write_token=False, append=0,
         max_pairs=8192, grid=1, seed=0):
    """Run k_ppar_compress<mode> over `streams` (bytes objects, each no longer than the window)."""
    W = 1 << window
    stride = max(16, (max((len(s) for s in streams), default=0) + 15) // 16 * 16)
    assert stride <= W or mode in (3, 5)  # modes 3 / 5: the lap variants take streams of any length
    n = len(streams)
    inp = np.zeros((n, stride), np.uint8)
    sizes = np.zeros(n, np.uint32)
    for i, s in enumerate(streams):
        inp[i, :len(s)] = np.frombuffer(s, np.uint8)
        sizes[i] = len(s)
    out_stride = (2 + (stride * (literal + 1) + 7) // 8 + 6 + 3) // 4 * 4
    out = np.full((n, out_stride), 0xEE, np.uint8)
    out_sizes = np.zeros(n, np.uint32)
    status = np.full(n, 99, np.int8)
    d = np.frombuffer(dictionary if dictionary is not None else oracle.initialize_dictionary(W, literal if mode in (2, WALK_EXT) else 8), np.uint8).copy()  # seed table: engine.cu, as compressor.c:209-213
    flags = (F_EXTENDED if mode in (2, WALK_EXT) else 0) | (F_LAZY if mode in (1, 5) else 0) | (F_DICT_RESET if dict_reset else 0) | \
            (F_CUSTOM if dictionary is not None else 0)
    flags |= _append_flags(append)
    if mode in (WALK, WALK_EXT):
        deferred = lib.emu_walk_compress(d.ctypes.data, window, literal, flags, int(write_token), max_pairs,
                                         inp.ctypes.data, sizes.ctypes.data, stride, out.ctypes.data, out_stride,
                                         out_sizes.ctypes.data, status.ctypes.data, n, grid, seed)
    else:
        deferred = lib.emu_ppar_compress(mode, d.ctypes.data, window, literal, flags, int(write_token), max_pairs,
                                         inp.ctypes.data, sizes.ctypes.data, stride, out.ctypes.data, out_stride,
                                         out_sizes.ctypes.data, status.ctypes.data, n, grid, seed)
    res = []
    for i in range(n):
        res.append(None if out_sizes[i] == DEFERRED else (out[i, :out_sizes[i]].tobytes(), int(status[i])))
    assert deferred == sum(r is None for r in res)
    return res


def _cases(harness, window, rng, count):
    W = 1 << window
    out = []
    for i in range(count):
        n = rng.choice([W, W, W - 1, W - 16, W // 2 + 3, 33, 32, 17, 16, 15, 2, 1, 0])
        out.append(gen_stream(harness, i % 6, 100 + i, n))
    return out


@pytest.mark.parametrize("mode", [0, 1, 2, WALK, WALK_EXT])
@pytest.mark.parametrize("window,seed", [(10, 0), (10, 3), (8, 5), (9, 11)])
def test_position_parallel_kernel_source_matches_the_oracle(emu, harness, mode, window, seed):
    rng = random.Random(1000 * window + seed)
    streams = _cases(harness, window, rng, 24)
    got = ppar(emu, mode, streams, window=window, seed=seed, max_pairs=8192 if mode != 1 else 0x7fffffff)
    done = 0
    for s, g in zip(streams, got):
        if g is None:
            continue  # left to the bitmap kernel (long chains / long runs): the GPU tests cover the pick-up pass
        want = oracle.compress(s, window=window, literal=8, extended=mode in (2, WALK_EXT), lazy_matching=mode == 1)
        assert g == (want, 0), (mode, window, len(s))
        done += 1
    assert done >= len(streams) // 2


@pytest.mark.parametrize("mode", [3, 5])  # 3: greedy v1, 5: v1 with lazy matching (cached match carried across laps)
@pytest.mark.parametrize("window,seed", [(8, 1), (9, 2), (10, 3), (10, 4)])
def test_position_parallel_lap_variant_source_matches_the_oracle(emu, harness, window, seed, mode):
    """Modes 3 / 5: v1 streams LONGER than the window, one lap of W offsets at a time (previous lap in place of the
    dictionary, walk entry and partial output word carried over).  Lengths around every lap boundary."""
    rng = random.Random(50 * window + seed)
    W = 1 << window
    dic = bytes(rng.choice(b"abcde \n") for _ in range(W)) if seed % 2 else None
    lit = 7 if seed == 2 else 8
    streams = []
    for i, n in enumerate([0, 1, W - 1, W, W + 1, W + 15, W + 16, W + 17, 2 * W - 1, 2 * W, 2 * W + 1, 3 * W + 100, 5 * W - 3,
                           2500, 4000]):
        s = _crafted(harness, rng, max(n, 1), 10 * seed + i)[:n] if i % 2 else gen_stream(harness, (0, 1, 2, 4)[i % 4], 70 + i, n)
        streams.append(bytes(b & ((1 << lit) - 1) for b in s))
    got = ppar(emu, mode, streams, window=window, literal=lit, dictionary=dic, dict_reset=seed == 3, write_token=seed != 1,
               seed=seed, max_pairs=20000)
    done = 0
    for s, g in zip(streams, got):
        if g is None:
            continue
        want = oracle.compress(s, window=window, literal=lit, extended=False, dictionary=dic, dictionary_reset=seed == 3,
                               write_token=seed != 1, lazy_matching=mode == 5)
        assert g == (want, 0), (window, len(s))
        done += 1
    assert done >= 12
    # a literal that does not fit in a later lap: whole bytes of everything before it, TAMP_EXCESS_BITS
    if lit == 7:
        bad = bytearray(b & 127 for b in gen_stream(harness, 0, 99, 3 * W + 100))
        bad[2 * W + 40] = 0xF0
        g = ppar(emu, mode, [bytes(bad)], window=window, literal=7, dictionary=dic, seed=seed)[0]
        assert g[1] == oracle.EXCESS_BITS
        good = oracle.compress(bytes(bad[:2 * W + 40]), window=window, literal=7, extended=False, dictionary=dic,
                               lazy_matching=mode == 5)
        assert g[0] == good[:len(g[0])] and len(good) - len(g[0]) <= 4


def hwalk(lib, streams, *, window, literal=8, dictionary=None, dict_reset=False, write_token=False, append=0, cbits=0, hbits=0, seg=0,
          threads=0, budget=0, max_pairs=0, grid=1, seed=0):
    """Run k_hwalk_dict + k_hwalk_compress<seg> over `streams` (any length); 0 = the launcher's plan for the window."""
    W = 1 << window
    stride = max(16, (max((len(s) for s in streams), default=0) + 15) // 16 * 16)
    n = len(streams)
    inp = np.zeros((n, stride), np.uint8)
    sizes = np.zeros(n, np.uint32)
    for i, s in enumerate(streams):
        inp[i, :len(s)] = np.frombuffer(s, np.uint8)
        sizes[i] = len(s)
    out_stride = (2 + (stride * (literal + 1) + 7) // 8 + 6 + 3) // 4 * 4
    out = np.full((n, out_stride), 0xEE, np.uint8)
    out_sizes = np.zeros(n, np.uint32)
    status = np.full(n, 99, np.int8)
    d = np.frombuffer(dictionary if dictionary is not None else oracle.initialize_dictionary(W, 8), np.uint8).copy()
    flags = (F_DICT_RESET if dict_reset else 0) | (F_CUSTOM if dictionary is not None else 0)
    flags |= _append_flags(append)
    deferred = lib.emu_hwalk_compress(d.ctypes.data, window, literal, flags, int(write_token), cbits, hbits, seg, threads, budget,
                                      max_pairs, inp.ctypes.data, sizes.ctypes.data, stride, out.ctypes.data, out_stride,
                                      out_sizes.ctypes.data, status.ctypes.data, n, grid, seed)
    assert deferred >= 0, "layout does not fit shared memory"
    res = [None if out_sizes[i] == DEFERRED else (out[i, :out_sizes[i]].tobytes(), int(status[i])) for i in range(n)]
    assert deferred == sum(r is None for r in res)
    return res


# (window, plan overrides): small chunks put many chunk boundaries, ring wraps and head-table turns into short streams
HWALK_CASES = [(8, {}), (8, dict(cbits=8, seg=16, threads=32)), (9, dict(cbits=9, seg=32, threads=32)), (10, {}),
               (10, dict(cbits=9, seg=16, threads=64)), (11, dict(cbits=8, seg=16, threads=32)), (11, {}),
               (12, dict(cbits=10, seg=32, threads=64)), (13, dict(cbits=9, seg=16, threads=96)), (15, dict(cbits=10, seg=32, threads=32))]


@pytest.mark.parametrize("case", range(len(HWALK_CASES)))
def test_history_walk_kernel_source_matches_the_oracle(emu, harness, case):
    """k_hwalk_compress (v1, any window, any length; the default for everything the segment-walk kernel does not take):
    lengths around every chunk and lap boundary, crafted streams (short periods, runs: forced through the candidate
    walk with the give-up thresholds off), custom dictionary, narrow literals, dictionary_reset + FLUSH token."""
    window, plan = HWALK_CASES[case]
    rng = random.Random(31 * window + case)
    W = 1 << window
    Cc = 1 << plan.get("cbits", 10)
    dic = bytes(rng.choice(b"abcde \n") for _ in range(W)) if case % 2 else None
    lit = 7 if case % 3 == 2 else 8
    lens = [0, 1, 2, 17, W - 1, W, W + 1, W + 15, W + 16, W + 17, Cc - 1, Cc, Cc + 1, 2 * Cc + 5, W + Cc, W + 2 * Cc + 3,
            2 * W + 100, 3 * W + 1] if window <= 12 else [0, 1, 31, Cc, Cc + 1, 3 * Cc + 7, W - 1, W + 1, W + Cc + 9]
    streams = []
    for i, n in enumerate(lens):
        n = min(n, 9000 if window <= 12 else 40000)
        srcs = _crafted(harness, rng, max(n, 1), 10 * case + i)[:n] if i % 3 == 1 else gen_stream(harness, (0, 1, 3, 4)[i % 4], 70 + i, n)
        streams.append(bytes(b & ((1 << lit) - 1) for b in srcs))
    kw = dict(window=window, literal=lit, dictionary=dic, dict_reset=case % 4 == 3, write_token=case % 2 == 0)
    got = hwalk(emu, streams, seed=case, max_pairs=1 << 24, budget=1 << 28, **kw, **plan)
    for s, g in zip(streams, got):
        want = oracle.compress(s, window=window, literal=lit, extended=False, dictionary=dic, dictionary_reset=kw["dict_reset"],
                               write_token=kw["write_token"])
        assert g == (want, 0), (window, plan, len(s))
    # with the thresholds on, runs are left to the bitmap kernels; everything it does finish is still the oracle's
    runs = [gen_stream(harness, oracle.RUNS, 5 + i, min(3 * W + 5, 6000)) for i in range(3)] + streams[-2:]
    got = hwalk(emu, runs, seed=case + 1, **kw, **plan)
    assert any(g is None for g in got[:3])
    for s, g in zip(runs, got):
        if g is not None:
            assert g == (oracle.compress(bytes(b & ((1 << lit) - 1) for b in s), window=window, literal=lit, extended=False,
                                         dictionary=dic, dictionary_reset=kw["dict_reset"], write_token=kw["write_token"]), 0) \
                or lit < 8
    # a literal that does not fit, in the first chunk and in a later one: whole bytes of everything before it
    if lit == 7:
        for at in (5, Cc + 40, min(2 * W + 33, 8000)):
            bad = bytearray(b & 127 for b in gen_stream(harness, 0, 99, at + 300))
            bad[at] = 0xF0
            g = hwalk(emu, [bytes(bad)], seed=case, **dict(kw, write_token=False), **plan)[0]
            assert g[1] == oracle.EXCESS_BITS
            good = oracle.compress(bytes(bad[:at]), window=window, literal=7, extended=False, dictionary=dic,
                                   dictionary_reset=kw["dict_reset"])
            assert g[0] == good[:len(g[0])] and len(good) - len(g[0]) <= 4, at


def test_history_walk_kernel_source_lanes_out_of_lock_step(emu, harness):
    """Several CTAs, several scheduler seeds: the queue, the head-table turns and the staging line tolerate any order."""
    streams = [gen_stream(harness, g, 40 + g, n) for g, n in ((0, 5000), (1, 3000), (3, 4097), (4, 2500), (0, 700))]
    want = [(oracle.compress(s, window=10, extended=False), 0) for s in streams]
    for seed in range(4):
        assert hwalk(emu, streams, window=10, cbits=10, seg=16, threads=64, grid=2, seed=seed) == want
        assert hwalk(emu, streams, window=11, cbits=9, seg=32, threads=96, grid=3, seed=seed) == \
            [(oracle.compress(s, window=11, extended=False), 0) for s in streams]


def cwalk(lib, streams, *, window, literal=8, dictionary=None, dict_reset=False, write_token=False, append=0, cbits=0, hbits=0, threads=0,
          gl=32, budget=0, grid=1, seed=0):
    """Run k_cwalk_compress over `streams` (any length); 0 = the launcher's plan for the window."""
    W = 1 << window
    stride = max(16, (max((len(s) for s in streams), default=0) + 15) // 16 * 16)
    n = len(streams)
    inp = np.zeros((n, stride), np.uint8)
    sizes = np.zeros(n, np.uint32)
    for i, s in enumerate(streams):
        inp[i, :len(s)] = np.frombuffer(s, np.uint8)
        sizes[i] = len(s)
    out_stride = (2 + (stride * (literal + 1) + 7) // 8 + 6 + 3) // 4 * 4
    out = np.full((n, out_stride), 0xEE, np.uint8)
    out_sizes = np.zeros(n, np.uint32)
    status = np.full(n, 99, np.int8)
    d = np.frombuffer(dictionary if dictionary is not None else oracle.initialize_dictionary(W, 8), np.uint8).copy()
    flags = (F_DICT_RESET if dict_reset else 0) | (F_CUSTOM if dictionary is not None else 0)
    flags |= _append_flags(append)
    deferred = lib.emu_cwalk_compress(d.ctypes.data, window, literal, flags, int(write_token), cbits, hbits, threads, gl, budget,
                                      inp.ctypes.data, sizes.ctypes.data, stride, out.ctypes.data, out_stride,
                                      out_sizes.ctypes.data, status.ctypes.data, n, grid, seed)
    assert deferred >= 0, "layout does not fit shared memory"
    res = [None if out_sizes[i] == DEFERRED else (out[i, :out_sizes[i]].tobytes(), int(status[i])) for i in range(n)]
    assert deferred == sum(r is None for r in res)
    return res


# (window, plan overrides): small chunks put many chunk boundaries and ring wraps into short streams
CWALK_CASES = [(11, dict(cbits=10, hbits=10, threads=64)), (11, {}), (12, dict(cbits=10, threads=128)), (12, dict(threads=64)),
               (13, dict(cbits=10, hbits=11, threads=32)), (14, dict(cbits=11, hbits=11, threads=64)),
               (15, dict(cbits=10, hbits=10, threads=32)), (9, dict(cbits=10, hbits=10, threads=64)),
               (11, dict(cbits=10, hbits=10, threads=64, gl=16)), (12, dict(cbits=11, hbits=11, threads=128, gl=16)),
               (13, dict(cbits=10, hbits=11, threads=32, gl=16)), (15, dict(cbits=11, hbits=11, threads=256, gl=16))]


@pytest.mark.parametrize("case", range(len(CWALK_CASES)))
def test_cooperative_history_walk_kernel_source_matches_the_oracle(emu, harness, case):
    """k_cwalk_compress (v1, windows 11..15 by default; the source itself takes any window): lengths around every chunk,
    lap and walker-segment boundary, crafted streams (short periods, runs) with the give-up budget off, custom
    dictionary, narrow literals, dictionary_reset + FLUSH token; then the give-up path and a literal that does not fit."""
    window, plan = CWALK_CASES[case]
    rng = random.Random(17 * window + case)
    W = 1 << window
    Cc = 1 << plan.get("cbits", 11)
    dic = bytes(rng.choice(b"abcde \n") for _ in range(W)) if case % 2 else None
    lit = 7 if case % 3 == 2 else 8
    lens = [0, 1, 2, 17, W - 1, W, W + 1, W + 16, Cc - 1, Cc, Cc + 1, 2 * Cc + 5, W + Cc, W + 2 * Cc + 3, 2 * W + 100] \
        if window <= 12 else [0, 1, 31, Cc, Cc + 1, 3 * Cc + 7, W - 1, W + 1, W + Cc + 9]
    streams = []
    for i, n in enumerate(lens):
        n = min(n, 9000 if window <= 12 else 36000)
        srcs = _crafted(harness, rng, max(n, 1), 10 * case + i)[:n] if i % 3 == 1 else gen_stream(harness, (0, 1, 3, 4)[i % 4], 70 + i, n)
        streams.append(bytes(b & ((1 << lit) - 1) for b in srcs))
    kw = dict(window=window, literal=lit, dictionary=dic, dict_reset=case % 4 == 3, write_token=case % 2 == 0)
    okw = dict(window=window, literal=lit, extended=False, dictionary=dic, dictionary_reset=kw["dict_reset"],
               write_token=kw["write_token"])
    got = cwalk(emu, streams, seed=case, budget=1 << 30, **kw, **plan)
    for s, g in zip(streams, got):
        assert g == (oracle.compress(s, **okw), 0), (window, plan, len(s))
    # a tiny budget: long streams are given up (left to the bitmap kernel), what is finished is still the oracle's
    got = cwalk(emu, streams, seed=case + 1, budget=3, **kw, **plan)
    assert any(g is None for g in got)
    for s, g in zip(streams, got):
        assert g is None or g == (oracle.compress(s, **okw), 0)
    if lit == 7:
        for at in (5, Cc + 40, min(2 * W + 33, 8000)):
            bad = bytearray(b & 127 for b in gen_stream(harness, 0, 99, at + 300))
            bad[at] = 0xF0
            g = cwalk(emu, [bytes(bad)], seed=case, **dict(kw, write_token=False), **plan)[0]
            assert g[1] == oracle.EXCESS_BITS
            good = oracle.compress(bytes(bad[:at]), **dict(okw, write_token=False))
            assert g[0] == good[:len(g[0])] and len(good) - len(g[0]) <= 4, at


def test_cooperative_history_walk_kernel_source_lanes_out_of_lock_step(emu, harness):
    streams = [gen_stream(harness, g, 40 + g, n) for g, n in ((0, 5000), (1, 3000), (3, 4097), (4, 2500), (0, 700))]
    want = [(oracle.compress(s, window=11, extended=False), 0) for s in streams]
    for seed in range(4):
        assert cwalk(emu, streams, window=11, cbits=10, hbits=10, threads=128, grid=2, seed=seed) == want


@pytest.mark.parametrize("mode", [0, 2, WALK, WALK_EXT])
def test_position_parallel_kernel_source_options(emu, harness, mode):
    """Custom dictionary, dictionary_reset header, FLUSH token, narrow literals with excess bits, several CTAs."""
    rng = random.Random(77 + mode)
    window, W = 9, 512
    dic = bytes(rng.choice(b"etaoin shrdlu\n") for _ in range(W))
    streams = _cases(harness, window, rng, 40)  # more streams than one CTA has warps
    got = ppar(emu, mode, streams, window=window, dictionary=dic, dict_reset=True, write_token=True, grid=2, seed=9)
    for s, g in zip(streams, got):
        if g is None:
            continue
        want = oracle.compress(s, window=window, extended=mode in (2, WALK_EXT), dictionary=dic, dictionary_reset=True, write_token=True)
        assert g == (want, 0)
    # literal = 6: bytes >= 64 end the stream with TAMP_EXCESS_BITS after the whole bytes written so far
    texts = [bytes(b & 63 for b in gen_stream(harness, 0, 300 + i, 200)) for i in range(6)]
    bad = [t[:k] + b"\xf0" + t[k + 1:] for t, k in zip(texts, (0, 1, 57, 120, 198, 199))]
    got = ppar(emu, mode, texts + bad, window=8, literal=6, seed=4)
    for s, g in zip(texts, got[:6]):
        if g is not None:
            assert g == (oracle.compress(s, window=8, literal=6, extended=mode in (2, WALK_EXT)), 0)
    for s, g in zip(bad, got[6:]):
        if g is not None:
            assert g[1] == oracle.EXCESS_BITS


def _crafted(harness, rng, W, i):
    """Text with the things the extended format reacts to: short runs (some losing to a match), runs that reach the end
    of the input, a lone repeated byte at the end, repeats of 14..140 bytes (extended matches, capped at 133)."""
    n = rng.choice([W, W, W - 1, W - 7, W // 2, 200, 64, 40, 17, 5, 2, 1])
    base = bytearray(gen_stream(harness, rng.choice([0, 0, 0, 2, 3, 5]), 9000 + i, n))
    for _ in range(rng.randrange(0, 8)):
        if n < 4:
            break
        at, ln = rng.randrange(0, n - 1), rng.randrange(2, 10)
        base[at:at + ln] = bytes([base[at]]) * len(base[at:at + ln])
    for _ in range(rng.randrange(0, 4)):
        ln = rng.choice([14, 15, 16, 17, 30, 60, 133, 134, 140])
        if n < 64 or 2 * ln + 2 > n:
            continue
        src = rng.randrange(0, n - 2 * ln)
        dst = rng.randrange(src + ln, n - ln + 1)
        base[dst:dst + ln] = base[src:src + ln]
    if rng.random() < 0.3 and n >= 3:
        k = rng.randrange(1, min(n, 20))
        base[n - k:] = bytes([base[n - k - 1]]) * k
    if rng.random() < 0.1 and n >= 2:
        base[n - 1] = base[n - 2]
    return bytes(base[:n])


@pytest.mark.parametrize("mode,round_", [(2, r) for r in range(9)] + [(WALK_EXT, r) for r in range(12)])
def test_extended_format_parse_on_crafted_streams(emu, harness, round_, mode):
    rng = random.Random(4242 + round_)
    window = rng.choice([8, 9, 10, 10])
    W = 1 << window
    dic = None if round_ % 2 else bytes(rng.choice(b"abcde \n") for _ in range(W))
    streams = [_crafted(harness, rng, W, 100 * round_ + i) for i in range(30)]
    got = ppar(emu, mode, streams, window=window, dictionary=dic, seed=round_, max_pairs=30000)
    done = 0
    for s, g in zip(streams, got):
        if g is None:
            continue  # a run longer than 8 bytes before the end: the window becomes parse-dependent (bitmap kernel)
        assert g == (oracle.compress(s, window=window, extended=True, dictionary=dic), 0), (window, len(s))
        done += 1
    assert done >= 15


# ---- k_fast_decompress (lane per stream) ------------------------------------------------------------------------------

def _seed_tables():
    t = np.zeros(3 * 32768, np.uint8)
    for k, lit in enumerate((5, 6, 8)):
        t[k * 32768:(k + 1) * 32768] = np.frombuffer(oracle.initialize_dictionary(32768, lit), np.uint8)
    return t


def fdec(lib, frames, cap, *, wmaxbits=10, window_bits_max=None, dictionary=None, packed=False, grid=1, seed=0, split=False):
    """Run k_fast_decompress<wmaxbits> over `frames` (bytes objects) with `cap` output bytes per row.  split: the default
    dispatch for rows no longer than the window — k_split_decompress first, k_fast_decompress picks up what it deferred;
    returns (results, number of deferred streams)."""
    n = len(frames)
    sizes = np.array([len(f) for f in frames], np.uint32)
    if packed:
        blob = np.frombuffer(b"".join(frames) + b"\0" * 16, np.uint8).copy()
        offsets = np.concatenate([[0], np.cumsum(sizes[:-1], dtype=np.uint64)]).astype(np.uint64)
        in_stride, in_ptr, off_ptr = 0, blob.ctypes.data, offsets.ctypes.data
    else:
        in_stride = (max(int(sizes.max()), 1) + 15) // 16 * 16
        blob = np.zeros((n, in_stride), np.uint8)
        for i, f in enumerate(frames):
            blob[i, :len(f)] = np.frombuffer(f, np.uint8)
        in_ptr, off_ptr = blob.ctypes.data, None
    out = np.full((n, cap), 0xEE, np.uint8)
    out_sizes = np.zeros(n, np.uint32)
    status = np.full(n, 99, np.int8)
    tables = _seed_tables()
    d = np.frombuffer(dictionary, np.uint8).copy() if dictionary is not None else None
    wmax = window_bits_max if window_bits_max is not None else wmaxbits
    if split:
        deferred = lib.emu_split_decompress(tables.ctypes.data, d.ctypes.data if d is not None else None, wmax, in_ptr, off_ptr,
                                            sizes.ctypes.data, in_stride, out.ctypes.data, cap, out_sizes.ctypes.data,
                                            status.ctypes.data, n, grid, seed)
        assert deferred == int((out_sizes == DEFERRED).sum())
        lib.emu_fast_decompress_pickup(wmaxbits, tables.ctypes.data, d.ctypes.data if d is not None else None, wmax, in_ptr,
                                       off_ptr, sizes.ctypes.data, in_stride, out.ctypes.data, cap, out_sizes.ctypes.data,
                                       status.ctypes.data, n, grid, seed)
        return [(out[i, :out_sizes[i]].tobytes(), int(status[i])) for i in range(n)], deferred
    lib.emu_fast_decompress(wmaxbits, tables.ctypes.data, d.ctypes.data if d is not None else None, wmax, in_ptr, off_ptr,
                            sizes.ctypes.data, in_stride, out.ctypes.data, cap, out_sizes.ctypes.data,
                            status.ctypes.data, n, grid, seed)
    return [(out[i, :out_sizes[i]].tobytes(), int(status[i])) for i in range(n)]


@pytest.mark.parametrize("wmaxbits,extended,seed", [(10, False, 0), (10, True, 1), (8, True, 2), (9, False, 3)])
def test_lane_per_stream_decompressor_source_matches_the_oracle(emu, harness, wmaxbits, extended, seed):
    rng = random.Random(31 * wmaxbits + seed)
    W = 1 << wmaxbits
    plain, frames = [], []
    for i in range(40):
        window = rng.choice([8, wmaxbits, wmaxbits])
        n = rng.choice([0, 1, 15, 16, 17, 100, W - 1, W, W + 1, 2 * W + 37, 3000])
        s = _crafted(harness, rng, max(n, 1), 500 + i)[:n] if n and i % 2 else gen_stream(harness, i % 6, 700 + i, n)
        lit = 8 if i % 5 else 7
        s = bytes(b & 127 for b in s) if lit == 7 else s
        plain.append(s)
        frames.append(oracle.compress(s, window=window, literal=lit, extended=extended, dictionary_reset=i % 7 == 0,
                                      write_token=i % 3 == 0))
    cap = 3008
    got = fdec(emu, frames, cap, wmaxbits=wmaxbits, packed=seed % 2 == 1, grid=2, seed=seed)
    for s, f, g in zip(plain, frames, got):
        assert g == (s, oracle.INPUT_EXHAUSTED)
        assert g == oracle.decompress(f, window_bits_max=wmaxbits, cap=cap)


def test_lane_per_stream_decompressor_source_hostile_frames(emu, harness):
    """Truncated and corrupted frames, rows that are too small: output bytes and status as the reference has them."""
    rng = random.Random(99)
    frames = []
    for i in range(60):
        s = _crafted(harness, rng, 1024, 900 + i)
        f = bytearray(oracle.compress(s, window=rng.choice([8, 9, 10]), extended=i % 2 == 0))
        kind = i % 4
        if kind == 0:
            f = f[:rng.randrange(0, len(f))]
        elif kind == 1:
            for _ in range(3):
                f[rng.randrange(0, len(f))] ^= 1 << rng.randrange(8)
        elif kind == 2:
            f[0] = rng.randrange(256)
        frames.append(bytes(f))
    for cap in (1040, 256, 16):
        got = fdec(emu, frames, cap, wmaxbits=10, grid=3, seed=cap)
        for f, g in zip(frames, got):
            want = oracle.decompress(f, window_bits_max=10, cap=cap)
            if want[1] == oracle.INVALID_CONF and (f[0] & 4):
                assert g[1] == oracle.INVALID_CONF  # custom-dictionary header without a dictionary: rejected up front
                continue
            assert g == want, (cap, f[:4].hex(), len(f))


@pytest.mark.parametrize("window,extended,cap_kind,seed", [(10, False, "exact", 0), (10, False, "roomy", 1), (8, False, "exact", 2),
                                                           (9, True, "roomy", 3), (10, True, "short", 4), (10, False, "short", 5),
                                                           (10, True, "roomy", 6), (8, True, "exact", 7), (10, True, "exact", 8)])
def test_split_decompressor_source_matches_the_oracle(emu, harness, window, extended, cap_kind, seed):
    """k_split_decompress (parse / copy split) + the pick-up pass of k_fast_decompress: frames no longer than the window,
    rows of exactly the stream length (OUTPUT_FULL vs INPUT_EXHAUSTED at the last bit), with room, and too short; narrow
    literals, custom dictionary, packed frames.  v1 frames with room to spare must not be deferred."""
    rng = random.Random(77 * window + seed)
    W = 1 << window
    n = W if cap_kind != "short" else W - 16
    cap = {"exact": n, "roomy": W, "short": n - 48}[cap_kind]
    dic = bytes(rng.choice(b"abcdefgh \n") for _ in range(W)) if seed % 3 == 2 else None
    plain, frames = [], []
    for i in range(70):
        ln = n if i % 4 else rng.choice([0, 1, 2, 17, n - 1, n // 2])
        s = _crafted(harness, rng, max(ln, 1), 300 + i)[:ln] if i % 3 == 0 else gen_stream(harness, i % 6, 40 + i, ln)
        lit = 8 if i % 5 else 7
        s = bytes(b & 127 for b in s) if lit == 7 else s
        plain.append(s)
        frames.append(oracle.compress(s, window=window, literal=lit, extended=extended, dictionary=dic, write_token=i % 7 == 0))
    got, deferred = fdec(emu, frames, cap, wmaxbits=window, dictionary=dic, packed=seed % 2 == 1, grid=2, seed=seed, split=True)
    for s, f, g in zip(plain, frames, got):
        assert g == oracle.decompress(f, window_bits_max=window, cap=cap, dictionary=dic), (len(s), len(f))
        if cap >= len(s):
            assert g[0] == s
    if not extended and cap_kind == "roomy":
        assert deferred <= 70 // 7 + 1  # only the frames that end with a FLUSH token
    if extended and cap_kind == "roomy":
        # run / extended-match tokens are decoded in the split kernel too: besides the FLUSH frames only the streams
        # with a match behind a run of more than 8 bytes (run-heavy generator, some crafted streams) are left over
        print("deferred (extended, roomy):", deferred)
        assert deferred <= 35


def lsdec(lib, frames, cap, *, wmax, dictionary=None, packed=False, grid=1, seed=0):
    """Run k_lsplit_decompress over `frames`; None for the streams it left to the pick-up pass."""
    n = len(frames)
    sizes = np.array([len(f) for f in frames], np.uint32)
    if packed:
        blob = np.frombuffer(b"".join(frames) + b"\0" * 16, np.uint8).copy()
        offsets = np.concatenate([[0], np.cumsum(sizes[:-1], dtype=np.uint64)]).astype(np.uint64)
        in_stride, in_ptr, off_ptr = 0, blob.ctypes.data, offsets.ctypes.data
    else:
        in_stride = (max(int(sizes.max()), 1) + 15) // 16 * 16
        blob = np.zeros((n, in_stride), np.uint8)
        for i, f in enumerate(frames):
            blob[i, :len(f)] = np.frombuffer(f, np.uint8)
        in_ptr, off_ptr = blob.ctypes.data, None
    out = np.full((n, cap), 0xEE, np.uint8)
    out_sizes = np.zeros(n, np.uint32)
    status = np.full(n, 99, np.int8)
    tables = _seed_tables()
    d = np.frombuffer(dictionary, np.uint8).copy() if dictionary is not None else None
    deferred = lib.emu_lsplit_decompress(tables.ctypes.data, d.ctypes.data if d is not None else None, wmax, in_ptr, off_ptr,
                                         sizes.ctypes.data, in_stride, out.ctypes.data, cap, out_sizes.ctypes.data,
                                         status.ctypes.data, n, grid, seed)
    assert deferred == int((out_sizes == DEFERRED).sum())
    return [None if out_sizes[i] == DEFERRED else (out[i, :out_sizes[i]].tobytes(), int(status[i])) for i in range(n)]


@pytest.mark.parametrize("window,n,extended,cap_kind", [(10, 4096, False, "exact"), (10, 4096, True, "roomy"), (8, 1024, False, "exact"),
                                                        (12, 9000, False, "roomy"), (12, 9000, True, "exact"), (15, 36000, False, "exact"),
                                                        (13, 20000, True, "short"), (9, 300, False, "roomy"), (11, 5000, False, "short")])
def test_long_split_decompressor_source_matches_the_oracle(emu, harness, window, n, extended, cap_kind):
    """k_lsplit_decompress (any window, any row length; the frame's own output row is the history): rows that wrap the
    window many times, exact / roomy / short rows, crafted runs and long repeats (extended tokens, partly written
    tokens -> pick-up pass), narrow literals, custom dictionary, packed frames, truncated and corrupted frames.  What
    the kernel finishes must be the reference's bytes and status; the rest is marked for the window-keeping kernels."""
    rng = random.Random(13 * window + n)
    W = 1 << window
    dic = bytes(rng.choice(b"abcdefgh \n") for _ in range(W)) if window % 3 == 0 else None
    plain, frames = [], []
    for i in range(36):
        ln = min(n, n if i % 4 else rng.choice([0, 1, 2, 17, W - 1, W, W + 1, n // 2]))
        s = _crafted(harness, rng, max(ln, 1), 300 + i)[:ln] if i % 3 == 0 else gen_stream(harness, i % 6, 40 + i, ln)
        lit = 8 if i % 5 else 7
        s = bytes(b & 127 for b in s) if lit == 7 else s
        plain.append(s)
        frames.append(oracle.compress(s, window=window, literal=lit, extended=extended, dictionary=dic, write_token=i % 7 == 0))
    for k in range(8):  # hostile: truncated, bit-flipped
        f = bytearray(frames[3 + k])
        if k % 2 and len(f) > 8:
            f = f[:rng.randrange(1, len(f))]
        elif len(f) > 8:
            f[rng.randrange(1, len(f))] ^= 1 << rng.randrange(8)
        frames.append(bytes(f))
        plain.append(None)
    cap = {"exact": (n + 3) // 4 * 4, "roomy": (n + 64) // 4 * 4, "short": (n - 48) // 4 * 4}[cap_kind]
    got = lsdec(emu, frames, cap, wmax=window, dictionary=dic, packed=window % 2 == 1, grid=2, seed=window)
    done = 0
    for s, f, g in zip(plain, frames, got):
        if g is None:
            continue
        assert g == oracle.decompress(f, window_bits_max=window, cap=cap, dictionary=dic), (None if s is None else len(s), len(f))
        done += 1
    assert done >= (20 if not extended else 12)


def test_split_decompressor_source_hostile_frames(emu, harness):
    """Truncated / corrupted frames and random headers through the split kernel + pick-up: bytes and status as the
    reference has them (whatever the split kernel cannot decide is deferred, never guessed)."""
    rng = random.Random(4321)
    frames = []
    for i in range(90):
        s = _crafted(harness, rng, 1024, 1900 + i)
        f = bytearray(oracle.compress(s, window=rng.choice([8, 9, 10]), extended=i % 2 == 0))
        kind = i % 4
        if kind == 0:
            f = f[:rng.randrange(0, len(f))]
        elif kind == 1:
            for _ in range(3):
                f[rng.randrange(0, len(f))] ^= 1 << rng.randrange(8)
        elif kind == 2:
            f[0] = rng.randrange(256)
        frames.append(bytes(f))
    for cap in (1024, 256, 16):
        got, _ = fdec(emu, frames, cap, wmaxbits=10, grid=2, seed=cap, split=True)
        for f, g in zip(frames, got):
            want = oracle.decompress(f, window_bits_max=10, cap=cap)
            if want[1] == oracle.INVALID_CONF and (f[0] & 4):
                assert g[1] == oracle.INVALID_CONF
                continue
            assert g == want, (cap, f[:4].hex(), len(f))


# ---- k_fast_compress (bitmap compressor: streams of any length, pick-up pass behind the position-parallel kernel) ------

def fcomp(lib, streams, *, window, extended, literal=8, dictionary=None, dict_reset=False, write_token=False, append=0, grid=2,
          seed=0, pickup_of=None):
    """Run k_fast_compress.  pickup_of: results of ppar() — only its deferred streams are compressed (pick-up pass)."""
    W = 1 << window
    n = len(streams)
    stride = max(16, (max((len(s) for s in streams), default=0) + 15) // 16 * 16)
    inp = np.zeros((n, stride), np.uint8)
    sizes = np.zeros(n, np.uint32)
    for i, s in enumerate(streams):
        inp[i, :len(s)] = np.frombuffer(s, np.uint8)
        sizes[i] = len(s)
    out_stride = (2 + (stride * (literal + 1) + 7) // 8 + 6 + 3) // 4 * 4
    out = np.full((n, out_stride), 0xEE, np.uint8)
    out_sizes = np.zeros(n, np.uint32)
    if pickup_of is not None:
        for i, r in enumerate(pickup_of):
            out_sizes[i] = DEFERRED if r is None else 7
    status = np.full(n, 99, np.int8)
    d = np.frombuffer(dictionary if dictionary is not None else oracle.initialize_dictionary(W, literal if extended else 8),
                      np.uint8).copy()
    flags = (F_EXTENDED if extended else 0) | (F_DICT_RESET if dict_reset else 0) | (F_CUSTOM if dictionary is not None else 0)
    flags |= _append_flags(append)
    rc = lib.emu_fast_compress(d.ctypes.data, window, literal, flags, int(write_token), int(pickup_of is not None),
                               inp.ctypes.data, sizes.ctypes.data, stride, out.ctypes.data, out_stride,
                               out_sizes.ctypes.data, status.ctypes.data, n, grid, seed)
    assert rc == 0
    return [(out[i, :out_sizes[i]].tobytes(), int(status[i])) if status[i] != 99 else None for i in range(n)]


@pytest.mark.parametrize("window,extended,seed", [(10, False, 0), (10, True, 1), (8, True, 2), (9, False, 3), (8, False, 4)])
def test_bitmap_compressor_source_matches_the_oracle(emu, harness, window, extended, seed):
    rng = random.Random(17 * window + seed)
    W = 1 << window
    streams = []
    for i in range(20):
        n = rng.choice([0, 1, 15, 16, 17, W - 1, W, W + 1, 2 * W + 5, 1500, 2600])
        streams.append(_crafted(harness, rng, max(n, 1), 300 + i)[:n] if i % 2 else gen_stream(harness, i % 6, 400 + i, n))
    dic = bytes(rng.choice(b"etaoin shrdlu") for _ in range(W)) if seed % 2 else None
    got = fcomp(emu, streams, window=window, extended=extended, dictionary=dic, dict_reset=seed == 1,
                write_token=seed % 2 == 0, seed=seed)
    for s, g in zip(streams, got):
        want = oracle.compress(s, window=window, extended=extended, dictionary=dic, dictionary_reset=seed == 1,
                               write_token=seed % 2 == 0)
        assert g == (want, 0), (window, extended, len(s))


@pytest.mark.parametrize("extended", [False, True])
def test_pick_up_pass_completes_what_the_position_parallel_kernel_defers(emu, harness, extended):
    """Run-heavy and periodic streams are deferred (chain population / long runs); the bitmap kernel's pick-up launch
    compresses exactly those and leaves the others alone."""
    rng = random.Random(5 + extended)
    streams = [gen_stream(harness, (5, 3, 0)[i % 3], 40 + i, rng.choice([1024, 1000, 512])) for i in range(24)]
    first = ppar(emu, 2 if extended else 0, streams, window=10, seed=3, max_pairs=3000)
    assert 4 <= sum(r is None for r in first) < len(streams)
    second = fcomp(emu, streams, window=10, extended=extended, pickup_of=first, grid=1, seed=4)
    for s, a, b in zip(streams, first, second):
        want = (oracle.compress(s, window=10, extended=extended), 0)
        if a is None:
            assert b == want
        else:
            assert a == want and b is None


# ---- general kernels (every window 8..15, every option) and the synthetic generator --------------------------------------

def gcomp(lib, streams, *, window, literal=8, extended=True, lazy=False, dictionary=None, dict_reset=False, write_token=False, append=0,
          out_stride=None, wpc=2, seed=0):
    W = 1 << window
    n = len(streams)
    stride = max(16, (max((len(s) for s in streams), default=0) + 15) // 16 * 16)
    inp = np.zeros((n, stride), np.uint8)
    sizes = np.zeros(n, np.uint32)
    for i, s in enumerate(streams):
        inp[i, :len(s)] = np.frombuffer(s, np.uint8)
        sizes[i] = len(s)
    out_stride = out_stride or 2 + (stride * (literal + 1) + 7) // 8 + 8
    out = np.full((n, out_stride), 0xEE, np.uint8)
    out_sizes = np.zeros(n, np.uint32)
    status = np.full(n, 99, np.int8)
    d = np.frombuffer(dictionary if dictionary is not None else oracle.initialize_dictionary(W, literal if extended else 8),
                      np.uint8).copy()
    flags = (F_EXTENDED if extended else 0) | (F_LAZY if lazy else 0) | (F_DICT_RESET if dict_reset else 0) | \
            (F_CUSTOM if dictionary is not None else 0)
    flags |= _append_flags(append)
    lib.emu_generic_compress(d.ctypes.data, window, literal, flags, int(write_token), inp.ctypes.data, sizes.ctypes.data,
                             stride, out.ctypes.data, out_stride, out_sizes.ctypes.data, status.ctypes.data, n, wpc, seed)
    return [(out[i, :out_sizes[i]].tobytes(), int(status[i])) for i in range(n)]


def gdec(lib, frames, cap, *, window_bits_max=15, dictionary=None, seed=0):
    n = len(frames)
    sizes = np.array([len(f) for f in frames], np.uint32)
    in_stride = (max(int(sizes.max()), 1) + 15) // 16 * 16
    blob = np.zeros((n, in_stride), np.uint8)
    for i, f in enumerate(frames):
        blob[i, :len(f)] = np.frombuffer(f, np.uint8)
    out = np.full((n, cap), 0xEE, np.uint8)
    out_sizes = np.zeros(n, np.uint32)
    status = np.full(n, 99, np.int8)
    tables = _seed_tables()
    d = np.frombuffer(dictionary, np.uint8).copy() if dictionary is not None else None
    scratch = np.zeros(128 << window_bits_max, np.uint8)
    lib.emu_generic_decompress(tables.ctypes.data, d.ctypes.data if d is not None else None, window_bits_max,
                               scratch.ctypes.data, 128, blob.ctypes.data, None, sizes.ctypes.data, in_stride,
                               out.ctypes.data, cap, out_sizes.ctypes.data, status.ctypes.data, n, seed)
    return [(out[i, :out_sizes[i]].tobytes(), int(status[i])) for i in range(n)]


@pytest.mark.parametrize("window,literal,extended,lazy", [(12, 8, True, False), (15, 8, False, False), (11, 7, True, False),
                                                          (13, 5, False, False), (10, 8, True, True), (9, 8, False, True),
                                                          (14, 6, True, False)])
def test_general_kernels_source_round_trip_and_parity(emu, harness, window, literal, extended, lazy):
    rng = random.Random(window * 100 + literal)
    streams = []
    for i in range(6):
        n = rng.choice([0, 1, 17, 300, 1500, 2500])
        s = _crafted(harness, rng, max(n, 1), 60 + i)[:n] if i % 2 else gen_stream(harness, i % 6, 80 + i, n)
        streams.append(bytes(b & ((1 << literal) - 1) for b in s))
    dic = bytes(rng.choice(b"abc de") & ((1 << literal) - 1) for _ in range(1 << window)) if window in (11, 13) else None
    got = gcomp(emu, streams, window=window, literal=literal, extended=extended, lazy=lazy, dictionary=dic,
                dict_reset=window == 12, write_token=window % 2 == 0, seed=window)
    want = [oracle.compress(s, window=window, literal=literal, extended=extended, lazy_matching=lazy, dictionary=dic,
                            dictionary_reset=window == 12, write_token=window % 2 == 0) for s in streams]
    for s, g, w in zip(streams, got, want):
        assert g == (w, 0), (window, len(s))
    back = gdec(emu, want, 2512, window_bits_max=window, dictionary=dic, seed=literal)
    for s, w, b in zip(streams, want, back):
        assert b == (s, oracle.INPUT_EXHAUSTED)
        assert b == oracle.decompress(w, window_bits_max=window, dictionary=dic, cap=2512)
    # rows that are too small: OUTPUT_FULL with the bytes the reference delivers
    tight = gdec(emu, want, 96, window_bits_max=window, dictionary=dic)
    for w, t in zip(want, tight):
        assert t == oracle.decompress(w, window_bits_max=window, dictionary=dic, cap=96)


def test_synthetic_generator_source_matches_the_cpu_harness(emu, harness):
    """bench.py's inputs come from k_synth; the CPU baseline's from the harness: they must be the same bytes."""
    for kind in range(6):
        out = np.zeros((5, 700), np.uint8)
        emu.emu_synth(kind, 1000, 5, 700, out.ctypes.data)
        assert out.tobytes() == harness.generate(kind, 1000, 5, 700).tobytes(), kind


# ---- k_wide_compress (windows 11..15, one CTA per stream) ---------------------------------------------------------------

def wcomp(lib, streams, *, window, literal=8, extended=True, dictionary=None, dict_reset=False, write_token=False, append=0, grid=2,
          seed=0, multi=False):
    W = 1 << window
    n = len(streams)
    stride = max(16, (max((len(s) for s in streams), default=0) + 15) // 16 * 16)
    inp = np.zeros((n, stride), np.uint8)
    sizes = np.zeros(n, np.uint32)
    for i, s in enumerate(streams):
        inp[i, :len(s)] = np.frombuffer(s, np.uint8)
        sizes[i] = len(s)
    out_stride = (2 + (stride * (literal + 1) + 7) // 8 + 6 + 3) // 4 * 4
    out = np.full((n, out_stride), 0xEE, np.uint8)
    out_sizes = np.zeros(n, np.uint32)
    status = np.full(n, 99, np.int8)
    d = np.frombuffer(dictionary if dictionary is not None else oracle.initialize_dictionary(W, literal if extended else 8),
                      np.uint8).copy()
    flags = (F_EXTENDED if extended else 0) | (F_DICT_RESET if dict_reset else 0) | (F_CUSTOM if dictionary is not None else 0)
    flags |= _append_flags(append)
    assert lib.emu_wide_compress(d.ctypes.data, window, literal, flags, int(write_token), inp.ctypes.data, sizes.ctypes.data,
                                 stride, out.ctypes.data, out_stride, out_sizes.ctypes.data, status.ctypes.data, n, grid,
                                 seed, int(multi)) == 0
    return [(out[i, :out_sizes[i]].tobytes(), int(status[i])) for i in range(n)]


@pytest.mark.parametrize("window,literal,extended,multi", [
    (11, 8, True, False), (12, 8, False, False), (13, 7, True, False), (14, 8, False, False), (15, 8, True, False),
    (15, 6, False, False)])
def test_cta_per_stream_compressor_source_matches_the_oracle(emu, harness, window, literal, extended, multi):
    rng = random.Random(window * 10 + literal)
    W = 1 << window
    lengths = [0, 1, 17, 700, 2100] + ([W + 300] if window <= 12 else [3000])
    streams = []
    for i, n in enumerate(lengths):
        s = _crafted(harness, rng, max(n, 1), 20 + i)[:n] if i % 2 else gen_stream(harness, (0, 5, 3)[i % 3], 30 + i, n)
        streams.append(bytes(b & ((1 << literal) - 1) for b in s))
    dic = bytes(rng.choice(b"abc de") & ((1 << literal) - 1) for _ in range(W)) if window == 13 else None
    got = wcomp(emu, streams, window=window, literal=literal, extended=extended, dictionary=dic, dict_reset=window == 14,
                write_token=window % 2 == 1, seed=window, multi=multi)
    for s, g in zip(streams, got):
        want = oracle.compress(s, window=window, literal=literal, extended=extended, dictionary=dic,
                               dictionary_reset=window == 14, write_token=window % 2 == 1)
        assert g == (want, 0), (window, len(s))


# ---- output compaction ---------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("n,stride", [(1, 64), (1000, 48), (1024, 32), (2500, 40), (0, 16), (37, 5000), (300, 2049), (9, 70001)])
def test_compaction_kernels_source(emu, n, stride):
    rng = np.random.default_rng(n + stride)
    rows = rng.integers(0, 256, (max(n, 1), stride), dtype=np.uint8)
    sizes = rng.integers(0, stride + 1, max(n, 1), dtype=np.uint32)
    want = b"".join(rows[i, :sizes[i]].tobytes() for i in range(n))
    packed = np.full(len(want) + 16, 0xEE, np.uint8)
    offsets = np.zeros(n + 1, np.uint64)
    total = emu.emu_compact(rows.ctypes.data, stride, sizes.ctypes.data, n, packed.ctypes.data, len(want), offsets.ctypes.data, 3)
    assert total == len(want) and packed[:len(want)].tobytes() == want and packed[len(want)] == 0xEE
    assert offsets.tolist() == [0] + np.cumsum(sizes[:n], dtype=np.uint64).tolist()
    if n > 1:  # a buffer that is too small: the total says so, nothing is written past the capacity
        small = np.full(len(want), 0xEE, np.uint8)
        cap = len(want) // 2
        total = emu.emu_compact(rows.ctypes.data, stride, sizes.ctypes.data, n, small.ctypes.data, cap, offsets.ctypes.data, 4)
        assert total == len(want) and (small[cap:] == 0xEE).all()


# ---- k_wide_decompress (one warp per stream, window in shared memory; kernel mode 4) ---------------------------------------

def wdec(lib, frames, cap, *, window_bits_max, dictionary=None, packed=False, grid=2, seed=0):
    n = len(frames)
    sizes = np.array([len(f) for f in frames], np.uint32)
    if packed:
        blob = np.frombuffer(b"".join(frames) + b"\0" * 16, np.uint8).copy()
        offsets = np.concatenate([[0], np.cumsum(sizes[:-1], dtype=np.uint64)]).astype(np.uint64)
        in_stride, in_ptr, off_ptr = 0, blob.ctypes.data, offsets.ctypes.data
    else:
        in_stride = (max(int(sizes.max()), 1) + 15) // 16 * 16
        blob = np.zeros((n, in_stride), np.uint8)
        for i, f in enumerate(frames):
            blob[i, :len(f)] = np.frombuffer(f, np.uint8)
        in_ptr, off_ptr = blob.ctypes.data, None
    out = np.full((n, cap), 0xEE, np.uint8)
    out_sizes = np.zeros(n, np.uint32)
    status = np.full(n, 99, np.int8)
    tables = _seed_tables()
    d = np.frombuffer(dictionary, np.uint8).copy() if dictionary is not None else None
    wpc = min(16, (224 * 1024) >> window_bits_max)
    lib.emu_wide_decompress(tables.ctypes.data, d.ctypes.data if d is not None else None, window_bits_max, in_ptr, off_ptr,
                            sizes.ctypes.data, in_stride, out.ctypes.data, cap, out_sizes.ctypes.data, status.ctypes.data,
                            n, grid, wpc, seed)
    return [(out[i, :out_sizes[i]].tobytes(), int(status[i])) for i in range(n)]


@pytest.mark.parametrize("wmax,extended,seed", [(15, True, 0), (15, False, 1), (12, True, 2), (11, False, 3), (10, True, 4),
                                                (13, True, 5)])
def test_warp_per_stream_decompressor_source_matches_the_oracle(emu, harness, wmax, extended, seed):
    rng = random.Random(7 * wmax + seed)
    plain, frames = [], []
    for i in range(24):
        window = rng.choice([8, 10, wmax, wmax])
        window = min(window, wmax)
        W = 1 << window
        n = rng.choice([0, 1, 15, 16, 17, 100, 700, 2000, 3000] + ([W + 50] if W <= 2048 else []))
        s = _crafted(harness, rng, max(n, 1), 500 + i)[:n] if n and i % 2 else gen_stream(harness, i % 6, 700 + i, n)
        lit = 8 if i % 5 else 7
        s = bytes(b & 127 for b in s) if lit == 7 else s
        plain.append(s)
        frames.append(oracle.compress(s, window=window, literal=lit, extended=extended, dictionary_reset=i % 7 == 0,
                                      write_token=i % 3 == 0))
    for cap in (3104, 128):
        got = wdec(emu, frames, cap, window_bits_max=wmax, packed=seed % 2 == 1, seed=seed)
        for s, f, g in zip(plain, frames, got):
            assert g == oracle.decompress(f, window_bits_max=wmax, cap=cap), (wmax, len(s), cap)
            if cap > len(s):
                assert g == (s, oracle.INPUT_EXHAUSTED)


def test_warp_per_stream_decompressor_source_hostile_frames(emu, harness):
    rng = random.Random(123)
    frames = []
    for i in range(60):
        s = _crafted(harness, rng, 1024, 900 + i)
        f = bytearray(oracle.compress(s, window=rng.choice([8, 10, 12]), extended=i % 2 == 0))
        kind = i % 4
        if kind == 0:
            f = f[:rng.randrange(0, len(f))]
        elif kind == 1:
            for _ in range(3):
                f[rng.randrange(0, len(f))] ^= 1 << rng.randrange(8)
        elif kind == 2:
            f[0] = rng.randrange(256)
        frames.append(bytes(f))
    for cap in (1040, 256, 16):
        got = wdec(emu, frames, cap, window_bits_max=12, grid=3, seed=cap)
        for f, g in zip(frames, got):
            want = oracle.decompress(f, window_bits_max=12, cap=cap)
            if want[1] == oracle.INVALID_CONF and (f[0] & 4):
                assert g[1] == oracle.INVALID_CONF
                continue
            assert g == want, (cap, f[:4].hex(), len(f))


# ---- append mode in the batch kernels / the segments of ONE stream (tamp_b200_compress_segmented) -------

def _ref():
    if not oracle.ref_available():
        pytest.skip("oracle/_ref (the reference C) is not built")
    return oracle.Ref()


def _ref_frame(ref, data, *, append, window, literal=8, extended=False, write_token=True):
    """The unmodified reference: init(dictionary_reset, append) + compress_and_flush(write_token)."""
    c = oracle.RefCompressor(ref, window=window, literal=literal, extended=extended, dictionary_reset=True, append=append)
    assert c.init_res == 0
    out, consumed, res = c.compress_and_flush(bytes(data), len(data) * 9 // 8 + 64, write_token)
    assert res == 0 and consumed == len(data)
    return out


def _ref_segmented_stream(ref, segs, *, window, literal=8, extended=False):
    """ONE reference compressor: compress(segment) ; reset_dictionary() between segments ; flush(write_token=True)."""
    c = oracle.RefCompressor(ref, window=window, literal=literal, extended=extended, dictionary_reset=True)
    out = b""
    for i, s in enumerate(segs):
        if i:
            o, res = c.reset_dictionary(64)
            assert res == 0
            out += o
        o, consumed, res = c.compress(bytes(s), len(s) * 9 // 8 + 64)
        assert res == 0 and consumed == len(s)
        out += o
    o, res = c.flush(64, True)
    assert res == 0
    return out + o


APPEND_KERNELS = ["walk", "walk_ext", "ppar", "ppar_lazy", "ppar_ext", "ppar_laps", "hwalk", "cwalk", "fcomp", "fcomp_ext", "wcomp", "gcomp"]


@pytest.mark.parametrize("kernel", APPEND_KERNELS)
@pytest.mark.parametrize("append", [1, 2])
def test_append_mode_frames_of_every_compressor_source_match_the_reference(emu, harness, kernel, append):
    """conf.append in a batch (compressor.c:227-234): every stream (append = 1) or every stream behind the first (2: the
    segments of one stream) starts with FLUSH padded to 16 bits instead of a header; an empty append-mode stream does not
    get a second FLUSH (:784-794).  Frame by frame against the unmodified reference, and — for the segments — the
    concatenation against ONE reference compressor with tamp_compressor_reset_dictionary() between the segments."""
    ref = _ref()
    window = 12 if kernel in ("cwalk", "wcomp", "gcomp") else 10 if kernel != "hwalk" else 9
    W = 1 << window
    short = kernel in ("walk", "walk_ext", "ppar", "ppar_lazy", "ppar_ext")  # streams no longer than the window
    lens = [W, 700, 0, 1, W - 16, 333] if short else [2 * W + 48, W, 0, 5, 3 * W - 7]
    segs = [gen_stream(harness, (0, 3, 5, 0, 2, 1)[i % 6], 40 + i, n) for i, n in enumerate(lens)]
    extended = kernel in ("walk_ext", "ppar_ext", "fcomp_ext", "gcomp")
    lazy = kernel == "ppar_lazy"
    for write_token in (True, False):
        kw = dict(window=window, write_token=write_token, append=append)
        if kernel in ("walk", "walk_ext"):
            got = ppar(emu, WALK_EXT if extended else WALK, segs, max_pairs=1 << 30, **kw)
        elif kernel.startswith("ppar"):
            got = ppar(emu, {"ppar": 0, "ppar_lazy": 1, "ppar_ext": 2, "ppar_laps": 3}[kernel], segs, max_pairs=1 << 30, **kw)
        elif kernel == "hwalk":
            got = hwalk(emu, segs, budget=1 << 30, max_pairs=1 << 30, **kw)
        elif kernel == "cwalk":
            got = cwalk(emu, segs, budget=1 << 30, **kw)
        elif kernel in ("fcomp", "fcomp_ext"):
            got = fcomp(emu, segs, extended=extended, **kw)
        elif kernel == "wcomp":
            got = wcomp(emu, segs, extended=False, **kw)
        else:
            got = gcomp(emu, segs, extended=extended, **kw)
        if lazy and not oracle.ref_available(lazy=True):
            pytest.skip("lazy reference not built")
        r = oracle.Ref(lazy=True) if lazy else ref
        for i, (s, g) in enumerate(zip(segs, got)):
            c = oracle.RefCompressor(r, window=window, extended=extended, dictionary_reset=True, append=append == 1 or i > 0,
                                     lazy_matching=lazy)
            want, consumed, res = c.compress_and_flush(bytes(s), len(s) * 9 // 8 + 64, write_token)
            assert res == 0 and g == (want, 0), (kernel, append, write_token, i, len(s))
        if append == 2 and write_token and not lazy:
            # (without the empty segment: the segmented call never makes one, and reset_dictionary() on a compressor that
            # wrote nothing since the last reset repeats the double FLUSH where the append-mode frame stays at one)
            assert b"".join(g[0] for g, s in zip(got, segs) if s) == \
                _ref_segmented_stream(ref, [s for s in segs if s], window=window, extended=extended)


@pytest.mark.parametrize("decoder", ["fdec", "split", "lsdec", "wdec", "gdec"])
def test_segment_frames_through_every_decompressor_source(emu, harness, decoder):
    """BatchArgs::seg_header: the frames behind segment 0 start with the append-mode marker (55 80) and take their
    configuration from segment 0's header.  Rows of exactly the segment size (OUTPUT_FULL in front of the closing FLUSH,
    decompressor.c:433-463) and roomy rows; a frame that does not start with the marker is INVALID_CONF.  Expected
    bytes and status: the reference decompressor initialised with the configuration (no header to read)."""
    ref = _ref()
    window = 10 if decoder in ("fdec", "split") else 12
    W = 1 << window
    seg = W if decoder == "split" else 2 * W + 32
    extended = decoder in ("fdec", "gdec")
    segs = [gen_stream(harness, (0, 3, 0, 2, 0)[i], 90 + i, n) for i, n in enumerate([seg, seg, seg, seg, 777])]
    frames = [_ref_frame(ref, s, append=i > 0, window=window, extended=extended) for i, s in enumerate(segs)]
    frames.append(b"\x55\x81" + frames[1][2:])  # not the marker
    conf = oracle.pack_conf(window, 8, False, extended, True)
    for cap in (seg, seg + 64):
        want = []
        for i, f in enumerate(frames):
            if i == 0:
                d = oracle.RefDecompressor(ref, window_bits=window)
            else:
                d = oracle.RefDecompressor(ref, window_bits=window, conf=conf)
            out, consumed, res = d.decompress(f, cap)
            want.append((out, res))
        want[-1] = (b"", -3)
        for header_mode in (0x100, 0x300):
            fr = frames if header_mode == 0x100 else frames[1:]
            wt = want if header_mode == 0x100 else want[1:]
            emu.emu_set_seg_header(header_mode | frames[0][0])
            try:
                if decoder == "fdec":
                    got = fdec(emu, fr, cap, wmaxbits=window)
                elif decoder == "split":
                    got, deferred = fdec(emu, fr, cap, wmaxbits=window, split=True)
                    assert deferred <= 2 or cap > seg  # exact rows stop in front of the closing FLUSH: nothing to defer but the short / bad frames
                elif decoder == "lsdec":
                    got = lsdec(emu, fr, cap, wmax=window, packed=True)
                    pick = wdec(emu, fr, cap, window_bits_max=window, packed=True)
                    if cap == seg:
                        assert sum(g is None for g in got) <= 2
                    got = [p if g is None else g for g, p in zip(got, pick)]
                elif decoder == "wdec":
                    got = wdec(emu, fr, cap, window_bits_max=window)
                else:
                    got = gdec(emu, fr, cap, window_bits_max=window)
            finally:
                emu.emu_set_seg_header(0)
            for i, (g, w) in enumerate(zip(got, wt)):
                assert g == w, (decoder, cap, hex(header_mode), i, g[1], w[1], len(g[0]), len(w[0]))
            assert all(g[0] == s for g, s in zip(got, segs if header_mode == 0x100 else segs[1:]))


@pytest.mark.parametrize("decoder,window,n", [("split", 10, 1000), ("split", 8, 256), ("lsdec", 10, 4000), ("lsdec", 12, 9000)])
def test_split_decompressors_finish_frames_that_close_with_a_flush_token(emu, harness, decoder, window, n):
    """Frames written with flush(write_token = true) — what tamp.Compressor.flush() does by default — end with a FLUSH
    token and padding (compressor.c:784-794).  The split decompressors finish them themselves (INPUT_EXHAUSTED, like the
    reference: decompressor.c:501-514) instead of leaving them to the pick-up pass; a FLUSH in the middle of a frame, or
    one followed by another byte, still goes there.  Both formats, dictionary_reset headers, exact and roomy rows."""
    rng = random.Random(5 * window + n)
    frames, want = [], []
    for i in range(40):
        ext, dr = i % 2 == 1, i % 3 == 0
        m = n if i % 4 else rng.randrange(1, n)
        data = gen_stream(harness, (0, 3, 0, 1)[i % 4], 500 + i, m)
        f = oracle.compress(data, window=window, extended=ext, dictionary_reset=dr, write_token=True)
        if i == 7:
            f += b"\x00"          # a byte behind the closing FLUSH: not the end of the frame
        if i == 9:                # a FLUSH in the middle: two frames' tokens in one (the second without its header)
            g = oracle.compress(data, window=window, extended=ext, dictionary_reset=dr, write_token=False)
            f = f + g[2 if dr else 1:]
        frames.append(f)
    for cap in (n, n + 40):
        want = [oracle.decompress(f, window_bits_max=window, cap=cap) for f in frames]
        if decoder == "split":
            got, deferred = fdec(emu, frames, cap, wmaxbits=window, split=True, packed=True)
            assert deferred <= (2 if cap > n else 12), deferred
        else:
            first = lsdec(emu, frames, cap, wmax=window, packed=True)
            # (extended frames longer than the window may go to the pick-up pass for their partly written tokens)
            if cap > n:
                assert not any(g is None for i, g in enumerate(first) if i % 2 == 0 and i not in (7, 9))
            pick = wdec(emu, frames, cap, window_bits_max=window, packed=True)
            got = [p if g is None else g for g, p in zip(first, pick)]
        for i, (g, w) in enumerate(zip(got, want)):
            assert g == w, (decoder, window, cap, i, g[1], w[1], len(g[0]), len(w[0]))
