"""GPU parity of the batch entry points (include/tamp_b200.h) against the oracle, the committed
reference fixtures, and size-independent properties at BASELINE.json scale."""
import hashlib
import os
import random
from collections import defaultdict

import numpy as np
import pytest
import torch

import oracle
from conftest import gen_stream
from tamp_b200 import batch

pytestmark = pytest.mark.gpu

MODES = [0, 1, 2, 4]  # 0 = auto (specialised kernels), 1 = general kernels, 2 = no position-parallel compressor,
                       # 4 = position-parallel compressor without its lap variant (round-1 dispatch)
DMODES = MODES + [6]   # 6 = the long split decompressor for every batch (mode 0 takes it for large wide-window batches only)


@pytest.fixture(autouse=True)
def _reset_mode():
    yield
    batch.set_kernel_mode(0)


def _oracle_many(rows, **kw):
    """oracle.compress over many streams on all host cores (the C restatement runs outside the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        return list(ex.map(lambda row: oracle.compress(row, **kw), rows))


def _rows(buf: torch.Tensor, sizes: torch.Tensor):
    b = buf.cpu().numpy()
    s = sizes.cpu().numpy().astype(np.int64)
    return [b[i, :s[i]].tobytes() for i in range(b.shape[0])]


@pytest.mark.parametrize("mode", MODES)
def test_reference_fixtures_through_batch_api(ref_fixtures, harness, mode):
    """Digests recorded from the unmodified reference C; ragged lengths incl. 0/1/15/16/17."""
    batch.set_kernel_mode(mode)
    groups = defaultdict(list)
    for f in ref_fixtures:
        conf = dict(f["conf"])
        if conf.get("lazy_matching"):
            continue  # lazy matching is exercised through the per-object path
        key = (conf["window"], conf.get("literal", 8), bool(conf.get("extended")), bool(conf.get("dictionary_reset")),
               bool(conf.get("write_token")))
        groups[key].append(f)
    checked = 0
    for (w, lit, ext, dr, wt), fs in groups.items():
        stride = max(16, max(f["n"] for f in fs))
        host = np.zeros((len(fs), stride), dtype=np.uint8)
        sizes = np.array([f["n"] for f in fs], dtype=np.int32)
        for i, f in enumerate(fs):
            d = gen_stream(harness, f["gen"], f["k"], f["n"], lit)
            host[i, :len(d)] = np.frombuffer(d, dtype=np.uint8)
        x = torch.from_numpy(host).cuda()
        r = batch.compress_batch(x, window=w, literal=lit, extended=ext, dictionary_reset=dr, write_token=wt,
                                 sizes=torch.from_numpy(sizes).cuda())
        torch.cuda.synchronize()
        assert (r.status == 0).all()
        for f, row in zip(fs, _rows(r.data, r.sizes)):
            assert len(row) == f["size"] and hashlib.sha256(row).hexdigest() == f["sha"], (w, lit, ext, f["gen"], f["n"])
            checked += 1
        # and back: decompress our bytes into n+16 bytes of room -> INPUT_EXHAUSTED; exact room -> recorded status
        d = batch.decompress_batch(r.data, r.sizes, stride + 16, window_bits_max=w)
        torch.cuda.synchronize()
        assert (d.sizes.cpu().numpy() == sizes).all() and (d.status == 2).all()
        assert (d.data[:, :stride].cpu().numpy()[np.arange(stride)[None, :] < sizes[:, None]] ==
                host[np.arange(stride)[None, :] < sizes[:, None]]).all()
    assert checked > 700


@pytest.mark.parametrize("mode", DMODES)
@pytest.mark.parametrize("window,n,ext", [(10, 1024, False), (10, 1024, True), (8, 1024, True), (10, 4096, False),
                                          (10, 4096, True), (12, 4096, True), (9, 600, False), (11, 2000, True),
                                          (15, 8192, True), (13, 5000, False)])
def test_differential_vs_oracle(harness, window, n, ext, mode):
    """Same seeded bytes through the oracle harness (CPU) and the CUDA batch path; memcmp every stream,
    then cross-decompress both ways."""
    batch.set_kernel_mode(mode)
    n_streams = 192
    for gen in (oracle.TEXT, oracle.RUNS, oracle.RAND, oracle.PERIODIC, oracle.BINARY):
        first_k = 1000 * gen + window
        host = harness.generate(gen, first_k, n_streams, n)
        exp, esz, est, _ = harness.compress(host, window=window, extended=ext)
        assert (est == 0).all()
        x = batch.synth(gen, first_k, n_streams, n)
        assert (x.cpu().numpy() == host).all(), "device generator drifted from the oracle's"
        r = batch.compress_batch(x, window=window, extended=ext, out_stride=exp.shape[1])
        torch.cuda.synchronize()
        got, gsz = r.data.cpu().numpy(), r.sizes.cpu().numpy().astype(np.uint32)
        assert (r.status == 0).all()
        assert (gsz == esz).all(), (gen, np.nonzero(gsz != esz)[0][:5])
        mask = np.arange(exp.shape[1])[None, :] < esz[:, None]
        assert (got[mask] == exp[mask]).all(), gen
        # CUDA decompress of the oracle's bytes
        d = batch.decompress_batch(torch.from_numpy(exp).cuda(), torch.from_numpy(esz.astype(np.int32)).cuda(), n + 32,
                                   window_bits_max=window)
        torch.cuda.synchronize()
        assert (d.sizes.cpu().numpy() == n).all() and (d.status == 2).all()
        assert (d.data[:, :n].cpu().numpy() == host).all()
        # oracle decompress of CUDA's bytes
        back, bsz, bst, _ = harness.decompress(got, gsz, n + 32, window_bits_max=window)
        assert (bsz == n).all() and (back[:, :n] == host).all()


@pytest.mark.parametrize("lazy", [False, True])
@pytest.mark.parametrize("window,n", [(8, 1024), (9, 1500), (10, 4096), (10, 1040), (8, 300)])
def test_lap_variant_streams_longer_than_the_window(harness, window, n, lazy):
    """v1 streams longer than the window through k_ppar_compress<kModeLaps / kModeLazyLaps> (ragged sizes, every
    generator) — the default dispatch since round 2 (measured 2.4x the bitmap kernel at window 8)."""
    batch.set_kernel_mode(0)
    n_streams = 128
    rng = random.Random(window + n)
    stride = (n + 15) // 16 * 16
    sizes = np.array([n, n - 1, 0, 1, (1 << window), (1 << window) + 1, (1 << window) - 1] +
                     [rng.randrange(0, n + 1) for _ in range(n_streams - 7)], dtype=np.int32)
    for gen in (oracle.TEXT, oracle.RUNS, oracle.RAND, oracle.PERIODIC, oracle.BINARY, 2):
        host = harness.generate(gen, 900 * gen + window, n_streams, stride)
        exp = [oracle.compress(host[i, :sizes[i]].tobytes(), window=window, extended=False, write_token=gen % 2 == 0,
                               lazy_matching=lazy) for i in range(n_streams)]
        r = batch.compress_batch(torch.from_numpy(host).cuda(), window=window, extended=False, write_token=gen % 2 == 0,
                                 sizes=torch.from_numpy(sizes).cuda(), **({"lazy_matching": True} if lazy else {}))
        torch.cuda.synchronize()
        got, gsz = r.data.cpu().numpy(), r.sizes.cpu().numpy()
        assert (r.status == 0).all()
        for i in range(n_streams):
            assert got[i, :gsz[i]].tobytes() == exp[i], (gen, i, int(sizes[i]))
    batch.set_kernel_mode(0)


@pytest.mark.parametrize("mode", [0, 6])
@pytest.mark.parametrize("window,n,ext", [(12, 4096, True), (15, 8192, True), (13, 5000, False), (11, 2000, True), (15, 70000, False),
                                          (10, 4096, False)])
def test_warp_per_stream_decompressor_wide_windows(harness, window, n, ext, mode):
    """Frames with windows 11..15 (and rows longer than a small window) through k_wide_decompress / k_fast_decompress
    (mode 0 at this batch size) and through k_lsplit_decompress + its pick-up pass (mode 6: what mode 0 takes for large
    batches); rows with room and rows that are too small."""
    batch.set_kernel_mode(mode)
    n_streams = 96 if n <= 8192 else 32
    for gen in (oracle.TEXT, oracle.RUNS, oracle.PERIODIC, oracle.BINARY):
        host = harness.generate(gen, 300 * gen + window, n_streams, n)
        exp, esz, est, _ = harness.compress(host, window=window, extended=ext)
        for cap in (n + 32, n - 100):
            d = batch.decompress_batch(torch.from_numpy(exp).cuda(), torch.from_numpy(esz.astype(np.int32)).cuda(), cap,
                                       window_bits_max=window)
            torch.cuda.synchronize()
            back, bsz, bst, _ = harness.decompress(exp, esz, cap, window_bits_max=window)
            assert (d.sizes.cpu().numpy() == bsz).all() and (d.status.cpu().numpy() == bst).all()
            got = d.data.cpu().numpy()
            m = np.arange(cap)[None, :] < bsz[:, None]
            assert (got[:, :cap][m] == back[:, :cap][m]).all()
    batch.set_kernel_mode(0)


@pytest.mark.parametrize("mode", DMODES)
def test_exact_capacity_and_truncated_output(harness, mode):
    """Config 4 shape: frames decoded into exactly-n-byte rows.  Status/size per stream must equal the
    reference semantics restated by the oracle (OUTPUT_FULL when pad bits remain, partial tokens cut)."""
    batch.set_kernel_mode(mode)
    n, n_streams = 4096, 64
    host = harness.generate(oracle.TEXT, 77, n_streams, n)
    comp, csz, _, _ = harness.compress(host, window=10, extended=True)
    for cap in (n, n - 1, 100, 1):
        d = batch.decompress_batch(torch.from_numpy(comp).cuda(), torch.from_numpy(csz.astype(np.int32)).cuda(), cap,
                                   window_bits_max=10)
        torch.cuda.synchronize()
        exp, esz, est, _ = harness.decompress(comp, csz, cap, window_bits_max=10)
        assert (d.sizes.cpu().numpy() == esz).all()
        assert (d.status.cpu().numpy() == est).all()
        assert (d.data.cpu().numpy() == exp).all()


@pytest.mark.parametrize("mode", DMODES)
def test_hostile_and_truncated_frames(harness, mode):
    """Fuzz-style: corrupted / truncated frames never crash and report the oracle's status and bytes
    (fuzz/fuzz_decompressor.c, ctests/test_decompressor.c:79-97, devices/vectors/*)."""
    batch.set_kernel_mode(mode)
    rng = random.Random(123)
    n, n_streams = 700, 256
    host = harness.generate(oracle.RUNS, 9, n_streams, n)
    for ext in (False, True):
        comp, csz, _, _ = harness.compress(host, window=10, extended=ext)
        comp = comp.copy()
        csz = csz.copy()
        for i in range(n_streams):
            kind = i % 4
            if kind == 0:
                comp[i, rng.randrange(1, csz[i])] ^= 1 << rng.randrange(8)
            elif kind == 1:
                csz[i] = rng.randrange(0, csz[i])
            elif kind == 2:
                comp[i, :csz[i]] = np.frombuffer(rng.randbytes(int(csz[i])), dtype=np.uint8)
                comp[i, 0] = 0x58 | (2 if ext else 0)
            else:
                comp[i, 0] = rng.randrange(256)  # random header: window/literal/custom/extended/reserved
        cap = 2000
        exp, esz, est, _ = harness.decompress(comp, csz, cap, window_bits_max=10)
        d = batch.decompress_batch(torch.from_numpy(comp).cuda(), torch.from_numpy(csz.astype(np.int32)).cuda(), cap,
                                   window_bits_max=10)
        torch.cuda.synchronize()
        gst, gsz, got = d.status.cpu().numpy(), d.sizes.cpu().numpy(), d.data.cpu().numpy()
        assert (gst == est).all(), np.nonzero(gst != est)[0][:8]
        assert (gsz == esz).all()
        mask = np.arange(cap)[None, :] < esz[:, None]
        assert (got[mask] == exp[mask]).all()


@pytest.mark.parametrize("mode", DMODES)
def test_custom_dictionary_and_literal_widths(harness, mode):
    batch.set_kernel_mode(mode)
    rng = random.Random(5)
    for w, lit, ext in [(8, 8, False), (10, 7, True), (10, 5, True), (12, 6, False), (11, 5, True), (15, 7, False)]:
        n, n_streams = 1500, 48
        host = harness.generate(oracle.TEXT, 31 * w + lit, n_streams, n) & ((1 << lit) - 1)
        dic = bytes(rng.choice(host[0].tobytes()) for _ in range(1 << w))
        for dictionary in (None, dic):
            exp = [oracle.compress(host[i].tobytes(), window=w, literal=lit, extended=ext, dictionary=dictionary)
                   for i in range(n_streams)]
            dt = None if dictionary is None else torch.frombuffer(bytearray(dictionary), dtype=torch.uint8).cuda()
            r = batch.compress_batch(torch.from_numpy(host).cuda(), window=w, literal=lit, extended=ext, dictionary=dt)
            torch.cuda.synchronize()
            assert _rows(r.data, r.sizes) == exp, (w, lit, ext, dictionary is not None)
            d = batch.decompress_batch(r.data, r.sizes, n + 8, window_bits_max=w, dictionary=dt)
            torch.cuda.synchronize()
            assert (d.status == 2).all() and (d.data[:, :n].cpu().numpy() == host).all()
    # excess bits are reported per stream (compressor.c:629-631)
    bad = harness.generate(oracle.TEXT, 1, 4, 256)
    bad[2, 100] = 0xF0
    r = batch.compress_batch(torch.from_numpy(bad).cuda(), window=10, literal=7, extended=False)
    torch.cuda.synchronize()
    assert r.status.cpu().tolist() == [0, 0, oracle.EXCESS_BITS, 0]


@pytest.mark.parametrize("mode", [0, 2, 4])
@pytest.mark.parametrize("window", [8, 9, 10])
def test_streams_no_longer_than_the_window(harness, window, mode):
    """Streams with N <= W (mode 0: the position-parallel kernel, v1 and extended): every generator, ragged lengths
    from 0 to W, all literal widths, custom dictionary, dictionary_reset + flush token — memcmp against the oracle."""
    batch.set_kernel_mode(mode)
    W = 1 << window
    n_streams = 96
    rng = random.Random(window)
    sizes = np.array([W, W - 1, W - 15, W - 16, W - 17, 0, 1, 2, 3, 15, 16, 17, 31, 32, 33] +
                     [rng.randrange(0, W + 1) for _ in range(n_streams - 15)], dtype=np.int32)
    for gen in (oracle.TEXT, oracle.RUNS, oracle.RAND, oracle.PERIODIC, oracle.BINARY, 2):
        host = harness.generate(gen, 500 * gen + window, n_streams, W)
        if gen == oracle.TEXT:  # runs of 2..12 equal bytes sprinkled in: short-run rule, RLE tokens, the 8-byte limit
            host = host.copy()
            for i in range(n_streams):
                for _ in range(6):
                    at, ln = rng.randrange(0, W - 12), rng.randrange(2, 13)
                    host[i, at:at + ln] = host[i, at]
        for lit, dictionary, dr, wt in [(8, None, False, False), (8, None, True, True), (8, "custom", False, True),
                                        (7, None, False, False), (6, "custom", False, False), (5, None, False, False)]:
            data = host & ((1 << lit) - 1) if lit < 8 else host
            if dictionary == "custom":
                src = data[rng.randrange(n_streams)].tobytes()
                dic = bytes(src[(7 * i) % W] if i % 3 else src[i] for i in range(W))
            else:
                dic = None
            dt = None if dic is None else torch.frombuffer(bytearray(dic), dtype=torch.uint8).cuda()
            for ext in (False, True):
                exp = [oracle.compress(data[i, :sizes[i]].tobytes(), window=window, literal=lit, extended=ext,
                                       dictionary=dic, dictionary_reset=dr, write_token=wt) for i in range(n_streams)]
                r = batch.compress_batch(torch.from_numpy(np.ascontiguousarray(data)).cuda(), window=window,
                                         literal=lit, extended=ext, dictionary=dt, dictionary_reset=dr, write_token=wt,
                                         sizes=torch.from_numpy(sizes).cuda())
                torch.cuda.synchronize()
                assert (r.status == 0).all()
                got = _rows(r.data, r.sizes)
                bad = [i for i in range(n_streams) if got[i] != exp[i]]
                assert not bad, (window, gen, lit, dictionary, dr, wt, ext, bad[:5], [int(sizes[i]) for i in bad[:5]])
            if mode in (0, 2) and lit in (8, 6):
                # lazy matching (compressor.c:576-619) through the same kernel: second match table + serial walk
                expl = [oracle.compress(data[i, :sizes[i]].tobytes(), window=window, literal=lit, extended=False,
                                        dictionary=dic, dictionary_reset=dr, write_token=wt, lazy_matching=True)
                        for i in range(n_streams)]
                r = batch.compress_batch(torch.from_numpy(np.ascontiguousarray(data)).cuda(), window=window,
                                         literal=lit, extended=False, dictionary=dt, dictionary_reset=dr,
                                         write_token=wt, sizes=torch.from_numpy(sizes).cuda(), lazy_matching=True)
                torch.cuda.synchronize()
                assert (r.status == 0).all()
                got = _rows(r.data, r.sizes)
                bad = [i for i in range(n_streams) if got[i] != expl[i]]
                assert not bad, ("lazy", window, gen, lit, dictionary, dr, wt, bad[:5], [int(sizes[i]) for i in bad[:5]])
    # excess bits on a literal end the stream with whole bytes only (compressor.c:629-631)
    bad = harness.generate(oracle.TEXT, 1, 4, 256)
    bad[2, 100] = 0xF0
    r = batch.compress_batch(torch.from_numpy(bad).cuda(), window=window, literal=7, extended=False)
    torch.cuda.synchronize()
    assert r.status.cpu().tolist() == [0, 0, oracle.EXCESS_BITS, 0]
    batch.set_kernel_mode(1)
    r1 = batch.compress_batch(torch.from_numpy(bad).cuda(), window=window, literal=7, extended=False)
    torch.cuda.synchronize()
    assert _rows(r.data, r.sizes) == _rows(r1.data, r1.sizes)


@pytest.mark.parametrize("window,n,n_streams", [(8, 1024, 96), (9, 3000, 64), (10, 4096, 96), (10, 1040, 64), (11, 5000, 48),
                                                (12, 9000, 48), (13, 20000, 24), (14, 40000, 16), (15, 40000, 16)])
def test_history_walk_any_window_any_length(harness, window, n, n_streams):
    """v1 batches the segment-walk kernel does not take (streams longer than the window, windows 11..15) through
    k_hwalk_compress, the default dispatch for them: ragged sizes around every chunk / lap boundary, every generator
    (runs and short periods leave through the pick-up pass of the bitmap kernels), literal widths incl. a literal that
    does not fit in a later chunk, custom dictionary, dictionary_reset + FLUSH token — memcmp against the oracle, and
    against the general kernels (mode 1) for the failing streams' partial output."""
    W = 1 << window
    rng = random.Random(window * 7 + n)
    stride = (n + 15) // 16 * 16
    edges = [n, n - 1, 0, 1, W, W + 1, W - 1, 2 * W, 2 * W + 1, 2 * W - 1, W + 15, W + 16, W + 17, 1024, 1025, 2048, 2047,
             4096, 4097, 8192, 8191, 8193]
    sizes = np.array([min(e, n) for e in edges][:n_streams] +
                     [rng.randrange(0, n + 1) for _ in range(max(0, n_streams - len(edges)))], dtype=np.int32)
    for gen in (oracle.TEXT, oracle.RUNS, oracle.RAND, oracle.PERIODIC, oracle.BINARY, 2):
        host = harness.generate(gen, 700 * gen + window, n_streams, stride)
        confs = [(8, None, False, False), (8, "custom", True, True)]
        if gen in (oracle.TEXT, oracle.BINARY):
            confs += [(7, None, False, True), (6, "custom", False, False), (5, None, True, False)]
        for lit, dictionary, dr, wt in confs:
            data = host & ((1 << lit) - 1) if lit < 8 else host
            if dictionary == "custom":
                src = data[rng.randrange(n_streams)].tobytes()
                dic = bytes(src[(7 * i) % len(src)] if i % 3 else src[i % len(src)] for i in range(W))
            else:
                dic = None
            dt = None if dic is None else torch.frombuffer(bytearray(dic), dtype=torch.uint8).cuda()
            exp = _oracle_many([data[i, :sizes[i]].tobytes() for i in range(n_streams)], window=window, literal=lit,
                               extended=False, dictionary=dic, dictionary_reset=dr, write_token=wt)
            batch.set_kernel_mode(0)
            r = batch.compress_batch(torch.from_numpy(np.ascontiguousarray(data)).cuda(), window=window, literal=lit,
                                     extended=False, dictionary=dt, dictionary_reset=dr, write_token=wt,
                                     sizes=torch.from_numpy(sizes).cuda())
            torch.cuda.synchronize()
            assert (r.status == 0).all()
            got = _rows(r.data, r.sizes)
            bad = [i for i in range(n_streams) if got[i] != exp[i]]
            assert not bad, (window, gen, lit, dictionary, dr, wt, bad[:5], [int(sizes[i]) for i in bad[:5]])
    # a literal that does not fit ends the stream with whole bytes only (compressor.c:629-631): first chunk, later chunk
    bad = harness.generate(oracle.TEXT, 77, 6, stride) & 0x7F
    bad[1, min(100, n - 1)] = 0xF0
    bad[3, n - 1] = 0x80
    bad[4, n // 2] = 0xFF
    bad[5, 0] = 0x80
    outs = []
    for mode in (0, 1):
        batch.set_kernel_mode(mode)
        r = batch.compress_batch(torch.from_numpy(bad).cuda(), window=window, literal=7, extended=False,
                                 sizes=torch.full((6,), n, dtype=torch.int32).cuda())
        torch.cuda.synchronize()
        assert r.status.cpu().tolist() == [0, oracle.EXCESS_BITS, 0, oracle.EXCESS_BITS, oracle.EXCESS_BITS, oracle.EXCESS_BITS]
        outs.append(_rows(r.data, r.sizes))
    assert outs[0] == outs[1]
    batch.set_kernel_mode(0)


@pytest.mark.parametrize("window,n,n_streams", [(12, 16384, 64), (15, 65536, 64)])
def test_full_length_differentials_of_the_wide_classes(harness, window, n, n_streams):
    """BASELINE.json config 3 / the wide classes of config 5 at their full stream length: n_streams streams per generator
    against the C restatement (memcmp), decoded back by the CUDA decoder."""
    for gen, ns in ((oracle.TEXT, n_streams), (oracle.BINARY, 16), (oracle.PERIODIC, 8), (oracle.RUNS, 8)):
        host = harness.generate(gen, 31 * gen + window, ns, n)
        exp, esz, est, _ = harness.compress(host, window=window, extended=False)
        assert (est == 0).all()
        r = batch.compress_batch(torch.from_numpy(host).cuda(), window=window, extended=False)
        torch.cuda.synchronize()
        assert (r.status == 0).all()
        got, gsz = r.data.cpu().numpy(), r.sizes.cpu().numpy().astype(np.uint32)
        assert (gsz == esz).all(), (gen, np.nonzero(gsz != esz)[0][:5])
        wid = min(exp.shape[1], got.shape[1])
        assert esz.max() <= wid
        mask = np.arange(wid)[None, :] < esz[:, None]
        assert (got[:, :wid][mask] == exp[:, :wid][mask]).all(), gen
        d = batch.decompress_batch(r.data, r.sizes, n, window_bits_max=window)
        torch.cuda.synchronize()
        assert (d.sizes.cpu().numpy() == n).all() and (d.data.cpu().numpy() == host).all()


def test_calls_on_two_streams_do_not_share_scratch(harness):
    """The *_device entry points only enqueue: calls issued on different CUDA streams may overlap, so every scratch
    buffer (the aligned copy of a custom dictionary, the general decompressor's windows, compaction sums) is private
    to its call.  Two streams, two different dictionaries / window classes, interleaved, no synchronisation between."""
    n_streams = 4096
    jobs = []
    for k, (window, n) in enumerate([(10, 1024), (9, 512)]):
        W = 1 << window
        host = harness.generate(oracle.TEXT, 50 + k, n_streams, n)
        dic = bytes(host[k].tobytes()[(5 * i + k) % n] for i in range(W))
        exp = _oracle_many([host[i].tobytes() for i in range(64)], window=window, extended=False, dictionary=dic)
        jobs.append((window, n, torch.from_numpy(host).cuda(), torch.frombuffer(bytearray(dic), dtype=torch.uint8).cuda(), exp, host))
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    for rep in range(3):
        res = []
        for (window, n, x, dt, exp, host), st in zip(jobs, streams):
            with torch.cuda.stream(st):
                r = batch.compress_batch(x, window=window, extended=False, dictionary=dt)
                d = batch.decompress_batch(r.data, r.sizes, n, dictionary=dt)                       # specialised kernels
                d15 = batch.decompress_batch(r.data, r.sizes, n, window_bits_max=window, dictionary=dt)
                packed, offsets = batch.compact(r)
                res.append((r, d, d15, packed, offsets))
        torch.cuda.synchronize()
        for (window, n, x, dt, exp, host), (r, d, d15, packed, offsets) in zip(jobs, res):
            rows = _rows(r.data[:64], r.sizes[:64])
            assert rows == exp, (rep, window)
            assert torch.equal(d.data, x) and torch.equal(d15.data, x)
            off = offsets.cpu().numpy()
            assert bytes(packed[off[5]:off[6]].cpu().numpy()) == exp[5]


def test_decompress_window_bound_and_dictionary_size(harness):
    """batch.decompress_batch: the default window bound comes from the frame headers (windows <= 10 reach the split /
    lane-per-stream kernels), a custom dictionary fixes it, and a dictionary shorter than 1 << window_bits_max is an
    error instead of an out-of-bounds read."""
    host = harness.generate(oracle.TEXT, 9, 256, 1024)
    x = torch.from_numpy(host).cuda()
    for window in (8, 10, 12):
        r = batch.compress_batch(x, window=window, extended=True)
        assert batch._window_bits_max(r.data[:, 0], None, None) == window
        d = batch.decompress_batch(r.data, r.sizes, 1040)
        torch.cuda.synchronize()
        assert torch.equal(d.data[:, :1024], x) and (d.status == 2).all() and (d.sizes == 1024).all()
        packed, offsets = batch.compact(r)
        dp = batch.decompress_packed(packed, offsets[:-1], r.sizes, 1040)
        torch.cuda.synchronize()
        assert torch.equal(dp.data[:, :1024], x) and (dp.status == 2).all()
        # a bound below the frames' window: TAMP_INVALID_CONF per stream, as tamp_decompressor_init would say
        if window > 8:
            bad = batch.decompress_batch(r.data, r.sizes, 1024, window_bits_max=window - 1)
            torch.cuda.synchronize()
            assert (bad.status == oracle.INVALID_CONF).all()
    dic = torch.frombuffer(bytearray(host[0].tobytes()), dtype=torch.uint8).cuda()   # 1 KiB
    r = batch.compress_batch(x, window=10, extended=False, dictionary=dic)
    assert torch.equal(batch.decompress_batch(r.data, r.sizes, 1024, dictionary=dic).data, x)
    with pytest.raises(ValueError):
        batch.decompress_batch(r.data, r.sizes, 1024, window_bits_max=15, dictionary=dic)
    with pytest.raises(ValueError):
        batch.decompress_batch(r.data, r.sizes, 1024, dictionary=dic[:1000])


def test_host_pointer_entry_points(harness):
    """tamp_b200_compress_batch / decompress_batch with HOST buffers (the e2e path of bench.py)."""
    n, n_streams = 1024, 300
    host = harness.generate(oracle.TEXT, 4242, n_streams, n)
    exp, esz, _, _ = harness.compress(host, window=10, extended=True)
    x = torch.from_numpy(host).pin_memory()
    r = batch.compress_batch(x, window=10, extended=True, out_stride=exp.shape[1])
    assert r.data.device.type == "cpu"
    assert (r.sizes.numpy().astype(np.uint32) == esz).all()
    mask = np.arange(exp.shape[1])[None, :] < esz[:, None]
    assert (r.data.numpy()[mask] == exp[mask]).all()
    d = batch.decompress_batch(r.data, r.sizes, n, window_bits_max=10)
    assert (d.data.numpy() == host).all()


def test_host_pointer_calls_from_two_threads(harness):
    """A host-pointer compress call and a host-pointer decompress call may run concurrently (separate staging slots per
    direction; bench.py's e2e leg overlaps step k's decompress with step k+1's compress this way)."""
    import threading
    n, n_streams = 1024, 20000  # >= 4096 streams: the pipelined path
    a = harness.generate(oracle.TEXT, 600, n_streams, n)
    b_ = harness.generate(oracle.BINARY, 601, n_streams, n)
    xa, xb = torch.from_numpy(a).pin_memory(), torch.from_numpy(b_).pin_memory()
    ra = batch.compress_batch(xa, window=10, extended=True)
    rb = batch.compress_batch(xb, window=10, extended=True)
    exp_a, esz_a, _, _ = harness.compress(a[:512], window=10, extended=True)
    got = {}

    def comp():
        for i in range(3):
            got["c%d" % i] = batch.compress_batch(xb, window=10, extended=True)

    def dec():
        for i in range(3):
            got["d%d" % i] = batch.decompress_batch(ra.data, ra.sizes, n + 16, window_bits_max=10)

    t1, t2 = threading.Thread(target=comp), threading.Thread(target=dec)
    t1.start(); t2.start(); t1.join(); t2.join()
    for i in range(3):
        assert torch.equal(got["c%d" % i].sizes, rb.sizes)
        m = torch.arange(rb.data.shape[1])[None, :] < rb.sizes[:, None]
        assert torch.equal(got["c%d" % i].data[m], rb.data[m])
        assert torch.equal(got["d%d" % i].data[:, :n], xa) and (got["d%d" % i].status == 2).all()
    rows = _rows(ra.data[:512], ra.sizes[:512])
    assert rows == [exp_a[i, :esz_a[i]].tobytes() for i in range(512)]


def test_host_pointer_packed_frames(harness):
    """tamp_b200_compress_batch_packed (contiguous frames + offsets in host memory) and the packed-input pipelined path of
    tamp_b200_decompress_batch: frames equal the reference's bytes, offsets are the prefix sum of the sizes, ragged
    inputs, small batches, a buffer that is too small, and the round trip."""
    for n, n_streams, ext in [(1024, 9000, False), (1024, 9000, True), (700, 300, False), (4096, 5000, False)]:
        host = harness.generate(oracle.TEXT, 77 + n_streams, n_streams, (n + 15) // 16 * 16)
        rng = random.Random(n)
        sizes = np.array([rng.randrange(0, n + 1) if i % 5 == 0 else n for i in range(n_streams)], dtype=np.int32)
        x = torch.from_numpy(host).pin_memory()
        packed, offsets, osz, st = batch.compress_batch_packed(x, window=10, extended=ext, sizes=torch.from_numpy(sizes))
        exp, esz, est, _ = harness.compress(host, window=10, extended=ext, sizes=sizes)
        assert (st == 0).all() and (osz.numpy().astype(np.uint32) == esz).all()
        off = offsets.numpy()
        assert off[0] == 0 and (np.diff(off) == esz).all()
        pk = packed.numpy()
        for i in list(range(0, n_streams, 97)) + [n_streams - 1]:
            assert pk[off[i]:off[i] + esz[i]].tobytes() == exp[i, :esz[i]].tobytes(), i
        d = batch.decompress_packed(packed, offsets, osz, host.shape[1] + 16, window_bits_max=10)
        assert (d.sizes.numpy() == sizes).all() and (d.status == 2).all()
        got = d.data.numpy()
        m = np.arange(host.shape[1] + 16)[None, :] < sizes[:, None]
        assert (got[m] == np.pad(host, ((0, 0), (0, 16)))[m]).all()
    small = torch.empty(1000, dtype=torch.uint8, pin_memory=True)
    with pytest.raises(batch.TampError):
        batch.compress_batch_packed(x, window=10, extended=False, packed=small)


def test_host_pointer_pipelined_path_ragged(harness):
    """>= 4096 strided streams take the chunked, stream-overlapped host path; rows travel as 2-D copies
    only as wide as the longest row.  Ragged lengths, both formats, byte-exact against the oracle."""
    n, n_streams = 1024, 70000
    host = harness.generate(oracle.TEXT, 99, n_streams, n)
    sizes = (np.arange(n_streams) * 37 % (n + 1)).astype(np.int32)
    for ext in (False, True):
        exp, esz, _, _ = harness.compress(host, window=10, extended=ext, sizes=sizes)
        x = torch.from_numpy(host).pin_memory()
        before = batch.copy_bytes()
        r = batch.compress_batch(x, window=10, extended=ext, sizes=torch.from_numpy(sizes), out_stride=exp.shape[1])
        after = batch.copy_bytes()
        assert after[0] > before[0] and after[1] > before[1]
        assert (r.sizes.numpy().astype(np.uint32) == esz).all() and (r.status.numpy() == 0).all()
        mask = np.arange(exp.shape[1])[None, :] < esz[:, None]
        assert (r.data.numpy()[mask] == exp[mask]).all()
        d = batch.decompress_batch(r.data, r.sizes, n, window_bits_max=10)
        assert (d.sizes.numpy() == sizes).all()
        m2 = np.arange(n)[None, :] < sizes[:, None]
        assert (d.data.numpy()[m2] == host[m2]).all()


def test_compaction_into_contiguous_frames(harness):
    """tamp_b200_compact_batch_device: offsets are the exclusive prefix sum of the sizes, the packed bytes are the
    rows' bytes back to back, and the packed layout decompresses through in_offsets / in_sizes."""
    for n_streams, n, w in [(1, 1024, 10), (1000, 700, 10), (3000, 1024, 10), (70000, 256, 8)]:
        x = batch.synth(oracle.TEXT, 11, n_streams, (n + 15) // 16 * 16)
        r = batch.compress_batch(x, window=w, extended=True)
        packed, offsets = batch.compact(r)
        torch.cuda.synchronize()
        sizes = r.sizes.cpu().numpy().astype(np.int64)
        exp_off = np.concatenate([[0], np.cumsum(sizes)])
        assert (offsets.cpu().numpy() == exp_off).all()
        rows = r.data.cpu().numpy()
        exp = np.concatenate([rows[i, :sizes[i]] for i in range(n_streams)])
        assert packed.numel() == exp.size and (packed.cpu().numpy() == exp).all()
        d = batch.decompress_packed(packed, offsets[:-1], r.sizes, x.shape[1] + 16, window_bits_max=w)
        torch.cuda.synchronize()
        assert torch.equal(d.data[:, :x.shape[1]], x) and (d.status == 2).all() and (d.sizes == x.shape[1]).all()
        # too little room: frames that do not fit are left out, the total is still reported
        small, off2 = batch.compact(r, capacity=int(exp_off[-1]) // 2)
        torch.cuda.synchronize()
        assert int(off2[-1].item()) == int(exp_off[-1])
        k = int(np.searchsorted(exp_off, int(exp_off[-1]) // 2, side="right")) - 1  # frames 0..k-1 fit entirely
        if k > 0:
            assert (small[:int(exp_off[k])].cpu().numpy() == exp[:int(exp_off[k])]).all()


def test_round_trip_at_baseline_scale():
    """BASELINE.json config 2 shape at a size the CPU cannot check stream by stream: 2^18 x 1 KiB through
    compress -> decompress must reproduce the input exactly; spot streams are memcmp'd with the oracle."""
    n_streams, n = 1 << 18, 1024
    x = batch.synth(oracle.TEXT, 0, n_streams, n)
    for ext in (False, True):
        r = batch.compress_batch(x, window=10, extended=ext)
        d = batch.decompress_batch(r.data, r.sizes, n, window_bits_max=10)
        torch.cuda.synchronize()
        assert (r.status == 0).all()
        assert torch.equal(d.data, x)
        assert (d.sizes == n).all()
        ratio = r.sizes.double().mean().item() / n
        assert 0.5 < ratio < 0.65, ratio
        idx = [0, 1, 12345, n_streams - 1]
        rows = _rows(r.data[idx], r.sizes[idx])
        host = x[idx].cpu().numpy()
        for j, i in enumerate(idx):
            assert rows[j] == oracle.compress(host[j].tobytes(), window=10, extended=ext), i
