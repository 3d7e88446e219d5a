#!/usr/bin/env python
"""Summarise an ncu report: headline metrics + hot SASS (instructions executed per source line)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 2e7
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__average_warps_issue_stalled', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed_pipe_lsu', 'sm__inst_executed_pipe', 'lts__t_bytes.sum', 'smsp__warps_eligible.avg']
for h, u, v in zip(hdr, units, vals):
    if any(w in h for w in want) and not h.endswith(('.min', '.max', 'peak_sustained_elapsed')):
        try:
            if float(v.replace(',', '')) == 0: continue
        except ValueError:
            pass
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ia = hdr.index("Instructions Executed"); isrc = hdr.index("Source"); ist = hdr.index("Warp Stall Sampling (All Samples)")
data = rows[2:]
tot = sum(int(r[ia]) for r in data)
print("total inst", tot, "n sass", len(data))
if thr > 0:
    for i, r in enumerate(data):
        c = int(r[ia])
        if c >= thr:
            print(i, f"{c/1e6:8.1f}M", r[ist].rjust(6), r[isrc].strip()[:100])
