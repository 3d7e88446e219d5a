"""GPU parity of the segmented single-stream path (tamp_b200_compress_segmented / tamp_b200_decompress_segmented,
SURVEY.md 8f rank 2) and of append mode in the batch entry points: against golden digests recorded from the unmodified
reference C (tests/golden/ref_segmented.json, make_segmented_fixtures.py) and, where oracle/_ref is present, against
the reference itself."""
import hashlib
import json

import numpy as np
import pytest
import torch

import oracle
from conftest import GOLDEN, gen_stream
from tamp_b200 import batch

pytestmark = pytest.mark.gpu

CASES = json.loads((GOLDEN / "ref_segmented.json").read_text())


@pytest.fixture(autouse=True)
def _reset_mode():
    yield
    batch.set_kernel_mode(0)


def _input(harness, c):
    data = gen_stream(harness, c["gen"], c["k"], c["n"], c["literal"])
    assert hashlib.sha256(data).hexdigest() == c["input_sha256"]
    return data


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("where", ["device", "host"])
def test_segmented_stream_is_the_reference_stream(harness, case, where):
    """One long input -> ONE Tamp stream, segment-parallel: the bytes are those of one reference compressor that resets
    its dictionary between the segments (digest recorded from the reference); the offsets are where its segments
    start; the segment-parallel decompressor gives the input back; a short output buffer is OUTPUT_FULL."""
    c = CASES[case]
    data = _input(harness, c)
    kw = dict(window=c["window"], literal=c["literal"], extended=c["extended"])
    seg = c["segment_size"]
    if where == "device":
        t = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda() if data else torch.empty(0, dtype=torch.uint8, device="cuda")
        stream, offs = batch.compress_segmented(t, seg, **kw)
        torch.cuda.synchronize()
        got, offsets = stream.cpu().numpy().tobytes(), offs.cpu().tolist()
    else:
        got, offsets = batch.compress_segmented(data, seg, **kw)
    assert len(got) == c["size"] and hashlib.sha256(got).hexdigest() == c["sha256"]
    assert hashlib.sha256(json.dumps(offsets).encode()).hexdigest() == c["offsets_sha256"]
    # the oracle (C restatement of the reference decoder) reads it front to back as one stream
    back, res = oracle.decompress(got, window_bits_max=c["window"], cap=len(data) + 64)
    assert back == data and res == oracle.INPUT_EXHAUSTED
    # segment-parallel
    if where == "device":
        out = batch.decompress_segmented(stream, offs, seg)
        assert out.cpu().numpy().tobytes() == data
        if len(data) > 1:
            with pytest.raises(batch.TampError) as e:
                batch.decompress_segmented(stream, offs, seg, out_size=len(data) - 1)
            assert e.value.status == 1
        exact = batch.decompress_segmented(stream, offs, seg, out_size=len(data))
        assert exact.cpu().numpy().tobytes() == data
    else:
        assert batch.decompress_segmented(got, offsets, seg) == data


@pytest.mark.parametrize("mode", [0, 1, 2, 4, 7])
@pytest.mark.parametrize("window,seg,extended", [(10, 1024, False), (10, 4096, False), (10, 4096, True), (12, 8192, False),
                                                 (8, 2048, True)])
def test_segmented_stream_through_every_kernel_mode(harness, mode, window, seg, extended):
    """Same stream whichever kernels run (walk / history walk / position-parallel / bitmap / general), checked against
    the oracle segment by segment (an append-mode frame = FLUSH padded to 16 bits + the frame body of a
    dictionary_reset stream) and decoded back through every decompressor."""
    batch.set_kernel_mode(mode)
    data = gen_stream(harness, 0, 900 + window, 20 * seg + 777)
    got, offsets = batch.compress_segmented(torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda(), seg, window=window,
                                            extended=extended)
    raw = got.cpu().numpy().tobytes()
    offsets = offsets.cpu().tolist()
    want = b""
    for i in range(0, len(data), seg):
        f = oracle.compress(data[i:i + seg], window=window, extended=extended, dictionary_reset=True, write_token=True)
        want += f if i == 0 else b"\x55\x80" + f[2:]
    assert raw == want
    for dmode in (0, 1, 2, 6):
        batch.set_kernel_mode(dmode)
        out = batch.decompress_segmented(got, torch.tensor(offsets), seg)
        assert out.cpu().numpy().tobytes() == data


def test_append_mode_batch_matches_the_reference(harness):
    """conf.append through tamp_b200_compress_batch_device: every stream starts with a FLUSH instead of a header
    (compressor.c:227-234); frame by frame against the unmodified reference."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref (the reference C) is not built")
    ref = oracle.Ref()
    for window, n, extended in ((10, 1024, False), (10, 1024, True), (10, 3000, False), (12, 5000, False), (9, 512, True)):
        stride = (n + 15) // 16 * 16
        rows = [gen_stream(harness, g, 300 + i, n if i % 3 else n // 2) for i, g in enumerate((0, 1, 2, 3, 5, 0, 0, 4))]
        rows.append(b"")
        host = np.zeros((len(rows), stride), np.uint8)
        for i, r in enumerate(rows):
            host[i, :len(r)] = np.frombuffer(r, np.uint8)
        sizes = torch.tensor([len(r) for r in rows], dtype=torch.int32)
        for write_token in (False, True):
            r = batch.compress_batch(torch.from_numpy(host).cuda(), window=window, extended=extended, dictionary_reset=True,
                                     append=True, write_token=write_token, sizes=sizes)
            torch.cuda.synchronize()
            out, osz, st = r.data.cpu().numpy(), r.sizes.cpu().numpy(), r.status.cpu().numpy()
            for i, row in enumerate(rows):
                c = oracle.RefCompressor(ref, window=window, extended=extended, dictionary_reset=True, append=True)
                want, consumed, res = c.compress_and_flush(row, len(row) * 9 // 8 + 64, write_token)
                assert res == 0 and st[i] == 0 and out[i, :osz[i]].tobytes() == want, (window, n, extended, write_token, i)
    with pytest.raises(batch.TampError):  # append needs dictionary_reset (compressor.c:209)
        batch.compress_batch(torch.zeros((1, 16), dtype=torch.uint8).cuda(), append=True)


def test_segments_written_by_the_reference_decode_in_parallel(harness):
    """A stream written by ONE reference compressor with resets every segment_size bytes, its segment offsets taken from
    where the resets fell: the segment-parallel decompressor reads it."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref (the reference C) is not built")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_segmented_fixtures", GOLDEN / "make_segmented_fixtures.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    ref = oracle.Ref()
    for window, seg, extended, n in ((10, 1024, True, 40_000), (10, 8192, False, 100_000), (13, 16384, True, 200_000)):
        data = gen_stream(harness, 0, 77 + window, n)
        stream, offsets = m.ref_segmented(ref, data, seg, window=window, literal=8, extended=extended)
        out = batch.decompress_segmented(torch.frombuffer(bytearray(stream), dtype=torch.uint8).cuda(), offsets, seg)
        assert out.cpu().numpy().tobytes() == data
        assert batch.decompress_segmented(stream, offsets, seg) == data


@pytest.mark.parametrize("mode", [0, 6])
@pytest.mark.parametrize("window,n,extended", [(10, 1024, False), (10, 1024, True), (10, 4096, False), (12, 16384, False)])
def test_frames_that_close_with_a_flush_token(harness, mode, window, n, extended):
    """Frames written with write_token = true (tamp.Compressor.flush()'s default): the split decompressors finish them
    themselves; bytes and status are the oracle's for exact and roomy rows."""
    batch.set_kernel_mode(mode)
    ns = 600
    x = batch.synth(0, 4242, ns, n)
    r = batch.compress_batch(x, window=window, extended=extended, write_token=True)
    sizes = r.sizes.cpu().numpy()
    rows = r.data.cpu().numpy()
    for cap in (n, n + 64):
        d = batch.decompress_batch(r.data, r.sizes, cap, window_bits_max=window)
        torch.cuda.synchronize()
        assert torch.equal(d.data[:, :n], x)
        st, osz = d.status.cpu().numpy(), d.sizes.cpu().numpy()
        for i in range(0, ns, 37):
            out, res = oracle.decompress(rows[i, :sizes[i]].tobytes(), window_bits_max=window, cap=cap)
            assert (osz[i], st[i]) == (len(out), res), (mode, window, n, extended, cap, i)


def test_host_pointer_segmented_compress_is_pipelined_and_equal(harness, monkeypatch):
    """tamp_b200_compress_segmented with 64 full segments or more goes through the pipelined packed path (chunks of
    segments: only the chunk that holds segment 0 writes a header) with the last, shorter segment behind it: same stream
    and offsets as the device-pointer call, chunk by chunk (TAMP_B200_CHUNK_MIB = 1: five chunks of 1024 segments)."""
    data = b"".join(gen_stream(harness, 0, 7000 + i, 65536) for i in range(80))[: 5000 * 1024 + 333]
    for window, seg, extended in ((10, 1024, False), (10, 1024, True), (10, 4096, False), (12, 16384, False)):
        dev_stream, dev_offs = batch.compress_segmented(torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda(), seg,
                                                        window=window, extended=extended)
        want, want_offs = dev_stream.cpu().numpy().tobytes(), dev_offs.cpu().tolist()
        for chunk_mib in ("1", "48"):
            monkeypatch.setenv("TAMP_B200_CHUNK_MIB", chunk_mib)
            got, offs = batch.compress_segmented(data, seg, window=window, extended=extended)
            assert offs == want_offs and got == want, (window, seg, extended, chunk_mib)
        assert batch.decompress_segmented(got, offs, seg) == data
    # the stream is what the oracle composes (segment 0 with its header, the rest with the append-mode marker)
    head = got[:offs[3]]
    exp = b""
    for i in range(3):
        f = oracle.compress(data[i * 16384:(i + 1) * 16384], window=12, extended=False, dictionary_reset=True, write_token=True)
        exp += f if i == 0 else b"\x55\x80" + f[2:]
    assert head == exp
