#!/usr/bin/env python
"""Benchmark of the Tamp batch codec hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference C on the host cores

Workload (BASELINE.json configs[1]): 2^20 independent 1 KiB synthetic-ASCII
streams (G_text) per GPU, window=10 literal=8 (one warp per stream in the compressor, one lane per stream
in the decompressor).  A step = compress the whole batch, then decompress it.
`value` = uncompressed MB/s of that step with everything resident in HBM; `e2e` = the same step
through the host-pointer C-ABI entry points with pinned host buffers (H2D/D2H inside the timed region).
Weak scaling: each rank owns its own 2^20-stream shard (streams are independent; no data-path
collective — DESIGN.md "Multi-GPU").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_STREAMS = 1 << 20
STREAM_LEN = 1024
WINDOW, LITERAL = 10, 8
METRIC = "uncompressed MB/s (compress+decompress) at 1/2/4/8 B200 vs C ref on host"
UNIT = "MB/s"


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i] == "Active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path on this box's host cores (oracle/_ref when it
    was built from /root/reference, else the oracle port), all host threads, bounded sample per step."""
    if rank != 0:
        return
    import numpy as np

    import oracle
    h = oracle.Harness("auto")
    cores = os.cpu_count() or 1
    n_sample = args.ref_streams or min(N_STREAMS, max(1 << 13, cores << 12))  # ~2 s of host work per step
    data = h.generate(oracle.TEXT, 0, n_sample, STREAM_LEN, threads=cores)
    stride = (STREAM_LEN * 9 + 7) // 8 + 32
    times, tc, td = [], [], []
    for it in range(args.warmup + args.steps):
        comp, csz, st, t_c = h.compress(data, window=WINDOW, literal=LITERAL, extended=bool(args.extended),
                                        out_stride=stride, threads=cores)
        back, bsz, st2, t_d = h.decompress(comp, csz, STREAM_LEN, window_bits_max=WINDOW, threads=cores)
        if it >= args.warmup:
            times.append(t_c + t_d)
            tc.append(t_c)
            td.append(t_d)
    assert (back == data).all()
    ms = 1e3 * sum(times) / len(times)
    value = n_sample * STREAM_LEN / 1e6 / (ms / 1e3)
    sample = f"{n_sample} x {STREAM_LEN} B G_text streams (first {n_sample} of the {N_STREAMS}-stream workload)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{N_STREAMS} x {STREAM_LEN} B synthetic ASCII streams (G_text), window={WINDOW} "
                               f"literal={LITERAL} extended={args.extended}", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": h.kind, "sample": sample,
                         "compress_MBps": n_sample * STREAM_LEN / 1e6 / (sum(tc) / len(tc)),
                         "decompress_MBps": n_sample * STREAM_LEN / 1e6 / (sum(td) / len(td))},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(extended: int, budget_streams: int):
    import oracle
    h = oracle.Harness("auto")
    cores = os.cpu_count() or 1
    n_sample = budget_streams
    data = h.generate(oracle.TEXT, 0, n_sample, STREAM_LEN, threads=cores)
    comp, csz, st, t_c = h.compress(data, window=WINDOW, literal=LITERAL, extended=bool(extended), threads=cores)
    back, bsz, st2, t_d = h.decompress(comp, csz, STREAM_LEN, window_bits_max=WINDOW, threads=cores)
    mb = n_sample * STREAM_LEN / 1e6
    return {"value": mb / (t_c + t_d), "unit": UNIT, "cores": cores, "kind": h.kind,
            "sample": f"first {n_sample} of the {N_STREAMS} G_text streams ({mb:.0f} MB), one pass, all host threads",
            "compress_MBps": mb / t_c, "decompress_MBps": mb / t_d}, comp, csz


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--extended", type=int, default=0, help="0 = v1 format (TampConf{.window,.literal}), 1 = v2")
    ap.add_argument("--streams", type=int, default=N_STREAMS, help="streams per GPU (default = the named config)")
    ap.add_argument("--ref-streams", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-format", action="store_true")
    ap.add_argument("--kernel-mode", type=int, default=0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import oracle  # checker only: cpu_baseline leg + spot parity check of the measured bytes
    from tamp_b200 import batch

    assert torch.cuda.is_available(), "bench.py needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    batch.set_kernel_mode(args.kernel_mode)

    n_streams = args.streams
    out_stride = (batch.compress_bound(STREAM_LEN, LITERAL) + 15) // 16 * 16
    ext = bool(args.extended)
    # each rank owns its own shard of the (weak-scaled) job: streams [rank * n_streams, (rank+1) * n_streams)
    x = batch.synth(oracle.TEXT, rank * n_streams, n_streams, STREAM_LEN, device=dev)
    comp = torch.empty((n_streams, out_stride), dtype=torch.uint8, device=dev)
    back = torch.empty((n_streams, STREAM_LEN), dtype=torch.uint8, device=dev)

    def step():
        r = batch.compress_batch(x, window=WINDOW, literal=LITERAL, extended=ext, out=comp)
        d = batch.decompress_batch(comp, r.sizes, STREAM_LEN, window_bits_max=WINDOW, out=back)
        return r, d

    for _ in range(max(args.warmup, 3)):
        r, d = step()
    torch.cuda.synchronize()
    assert torch.equal(back, x), "round trip failed"
    comp_bytes = int(r.sizes.sum().item())

    # ---- timed region: K steps, device events on the launching (current) stream --------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = batch.launch_count()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(args.steps):
        ev[i][0].record()
        r = batch.compress_batch(x, window=WINDOW, literal=LITERAL, extended=ext, out=comp)
        ev[i][1].record()
        d = batch.decompress_batch(comp, r.sizes, STREAM_LEN, window_bits_max=WINDOW, out=back)
        ev[i][2].record()
    t_end.record()
    torch.cuda.synchronize()
    launches = batch.launch_count() - launches0
    elapsed_ms = t_start.elapsed_time(t_end)
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
        dist.barrier()
    clocks = sampler.stop()
    ms_step = elapsed_ms / args.steps
    c_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    d_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps
    total_mb = world * n_streams * STREAM_LEN / 1e6
    value = total_mb / (ms_step / 1e3)

    peak, peak_src = measured_peak_gbs()
    algo_bytes = n_streams * STREAM_LEN + comp_bytes  # N + C per stream, both directions
    achieved = algo_bytes / 1e9 / (c_ms / 1e3)
    roofline = {"bound": "hbm", "kernel": "compress", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": c_ms,
                "hbm_read_frac": n_streams * STREAM_LEN / 1e9 / (c_ms / 1e3) / peak,
                "decompress": {"achieved": algo_bytes / 1e9 / (d_ms / 1e3), "kernel_ms": d_ms,
                               "frac": algo_bytes / 1e9 / (d_ms / 1e3) / peak}}
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        try:
            roofline["traffic"] = json.loads(tf.read_text()).get("compress_dram_bytes_per_launch")
        except Exception:
            pass

    # ---- the other format (report both; v1 is what a C caller's TampConf{.window,.literal} selects, v2 the
    # conf == NULL / Python default): same workload, same timing rules, reported beside the headline, not in it ----
    other = None
    if not args.no_other_format:
        oext = not ext
        for _ in range(3):
            ro = batch.compress_batch(x, window=WINDOW, literal=LITERAL, extended=oext, out=comp)
            do = batch.decompress_batch(comp, ro.sizes, STREAM_LEN, window_bits_max=WINDOW, out=back)
        torch.cuda.synchronize()
        assert torch.equal(back, x), "round trip failed (other format)"
        o_steps = max(2, min(args.steps, 5))
        oe = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        oc = od = 0.0
        for _ in range(o_steps):
            oe[0].record()
            ro = batch.compress_batch(x, window=WINDOW, literal=LITERAL, extended=oext, out=comp)
            oe[1].record()
            do = batch.decompress_batch(comp, ro.sizes, STREAM_LEN, window_bits_max=WINDOW, out=back)
            oe[2].record()
            torch.cuda.synchronize()
            oc += oe[0].elapsed_time(oe[1])
            od += oe[1].elapsed_time(oe[2])
        oc, od = oc / o_steps, od / o_steps
        mb1 = n_streams * STREAM_LEN / 1e6
        other = {"extended": int(oext), "compress_ms": oc, "decompress_ms": od, "MBps": mb1 / ((oc + od) / 1e3),
                 "compressed_ratio": int(ro.sizes.sum().item()) / (n_streams * STREAM_LEN), "n_gpus": 1,
                 "note": "per GPU, rank 0"}
        # leave the buffers as the headline format left them (the CPU parity check below reads them)
        r = batch.compress_batch(x, window=WINDOW, literal=LITERAL, extended=ext, out=comp)
        d = batch.decompress_batch(comp, r.sizes, STREAM_LEN, window_bits_max=WINDOW, out=back)
        torch.cuda.synchronize()

    # ---- e2e: host buffers through the host-pointer C-ABI entry points ----------------------------------
    e2e = None
    if not args.no_e2e:
        hx = torch.empty((n_streams, STREAM_LEN), dtype=torch.uint8, pin_memory=True)
        hx.copy_(x)
        hcomp = torch.empty((n_streams, out_stride), dtype=torch.uint8, pin_memory=True)
        hback = torch.empty((n_streams, STREAM_LEN), dtype=torch.uint8, pin_memory=True)
        e_steps = max(2, min(args.steps, 5))
        for it in range(1 + e_steps):
            if it == 1:
                if world > 1:
                    dist.barrier()
                bytes0 = batch.copy_bytes()
                t0 = time.perf_counter()
            hr = batch.compress_batch(hx, window=WINDOW, literal=LITERAL, extended=ext, out=hcomp)
            hd = batch.decompress_batch(hcomp, hr.sizes, STREAM_LEN, window_bits_max=WINDOW, out=hback)
        e_ms = (time.perf_counter() - t0) * 1e3 / e_steps
        if world > 1:
            t = torch.tensor([e_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t.item())
        assert torch.equal(hback, hx)
        bytes1 = batch.copy_bytes()
        e2e = {"value": total_mb / (e_ms / 1e3), "unit": UNIT, "ms_per_step": e_ms,
               "h2d_bytes_per_step": (bytes1[0] - bytes0[0]) // e_steps,
               "d2h_bytes_per_step": (bytes1[1] - bytes0[1]) // e_steps,
               "api": "tamp_b200_compress_batch + tamp_b200_decompress_batch (host pointers, pinned)"}

    # ---- CPU baseline beside it (rank 0, N = 1 only) + parity spot check of the measured bytes --------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cpu, ref_comp, ref_sz = cpu_baseline(args.extended, min(n_streams, max(1 << 14, cores << 15)))  # ~10-20 s of host work
        k = ref_comp.shape[0]
        got = comp[:k].cpu().numpy()
        gsz = r.sizes[:k].cpu().numpy().astype(np.uint32)
        assert (gsz == ref_sz).all(), "compressed sizes differ from the CPU reference"
        w = min(got.shape[1], ref_comp.shape[1])
        mask = np.arange(w)[None, :] < ref_sz[:, None]
        assert (got[:, :w][mask] == ref_comp[:, :w][mask]).all(), "bitstreams differ from the CPU reference"
        cpu["parity"] = f"bit-exact on the {k}-stream sample"

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"{n_streams} x {STREAM_LEN} B synthetic ASCII streams (G_text) per GPU, window="
                                   f"{WINDOW} literal={LITERAL} extended={args.extended}, one stream per warp (compress) / per lane (decompress)",
                       "l2": "inputs (1 GiB per GPU) exceed the 126 MB L2; no flush needed",
                       "compressed_ratio": comp_bytes / (n_streams * STREAM_LEN),
                       "kernel_mode": args.kernel_mode},
            "compress_MBps": total_mb / (c_ms / 1e3), "decompress_MBps": total_mb / (d_ms / 1e3),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "other_format": other,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
