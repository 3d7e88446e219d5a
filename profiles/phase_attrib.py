#!/usr/bin/env python
"""Per-phase instruction attribution of one kernel from an `ncu --set full --import-source on` capture.

    ncu -i X.ncu-rep --page source --csv --print-source sass > sass.csv
    cuobjdump -xelf all tamp_b200/_build/<file>.cu.o ; nvdisasm -g -c <file>.sm_100a.cubin > dis.txt
    python profiles/phase_attrib.py sass.csv dis.txt <mangled-kernel-substring> <phases.json|builtin-name>

The ncu SASS page carries "Instructions Executed" (warp instructions) and stall samples per SASS instruction but no
line numbers in its CSV; nvdisasm -g carries `//## File ..., line N` markers for the same instruction sequence.  The
two are joined by instruction index inside the kernel, and source lines are bucketed into the phases given as
{name: [first_line, last_line]} (inlined helpers are attributed to their own lines: give them their own bucket).
"""
import csv
import json
import re
import sys
from collections import defaultdict


def read_ncu(path):
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {n: hdr.index(n) for n in ("Source", "Instructions Executed", "# Samples", "Thread Instructions Executed")}
    out = []
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr):
            continue
        out.append((r[col["Source"]].strip(), int(r[col["Instructions Executed"]] or 0), int(r[col["# Samples"]] or 0),
                    int(r[col["Thread Instructions Executed"]] or 0)))
    return out


def read_dis(path, kernel_sub):
    insts, line, inside = [], 0, False
    for ln in open(path):
        if ln.startswith(".text."):
            inside = kernel_sub in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File ".*?", line (\d+)', ln)
        if m:
            line = int(m.group(1))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            insts.append((line, m.group(2).strip()))
    return insts


def main():
    sass_csv, dis_txt, ksub, phases_arg = sys.argv[1:5]
    phases = json.loads(open(phases_arg).read()) if phases_arg.endswith(".json") else json.loads(phases_arg)
    ncu = read_ncu(sass_csv)
    dis = read_dis(dis_txt, ksub)
    assert len(ncu) == len(dis), (len(ncu), len(dis))
    per_line = defaultdict(lambda: [0, 0, 0, 0])  # warp instr, samples, static count, thread instr
    for (src, n, s, t), (line, text) in zip(ncu, dis):
        op_a, op_b = src.split()[0].lstrip("@!P0123456789UT "), text.split()[0]
        per_line[line][0] += n
        per_line[line][1] += s
        per_line[line][2] += 1
        per_line[line][3] += t
    tot_i = sum(v[0] for v in per_line.values())
    tot_s = sum(v[1] for v in per_line.values())
    print(f"kernel {ksub}: {len(ncu)} SASS instructions, {tot_i} warp instructions executed, {tot_s} stall samples")
    print(f"{'phase':34s} {'warp-instr':>14s} {'share':>7s} {'samples':>9s} {'share':>7s} {'lanes':>6s} {'static':>7s}")
    seen = 0
    for name, ranges in phases.items():
        if ranges and not isinstance(ranges[0], list):
            ranges = [ranges]
        i = s = st = t = 0
        for lo, hi in ranges:
            for l, v in per_line.items():
                if lo <= l <= hi:
                    i += v[0]; s += v[1]; st += v[2]; t += v[3]
        seen += i
        lanes = t / i if i else 0
        print(f"{name:34s} {i:14d} {100 * i / tot_i:6.1f}% {s:9d} {100 * s / max(tot_s, 1):6.1f}% {lanes:6.1f} {st:7d}")
    print(f"{'(unbucketed)':34s} {tot_i - seen:14d} {100 * (tot_i - seen) / tot_i:6.1f}%")
    if len(sys.argv) > 5:  # per-line dump of the hottest lines
        print("\nhottest source lines:")
        for l, v in sorted(per_line.items(), key=lambda kv: -kv[1][0])[: int(sys.argv[5])]:
            print(f"  line {l:4d}: {v[0]:12d} warp-instr ({100 * v[0] / tot_i:4.1f}%), {v[2]:4d} SASS, {v[1]:6d} samples")


if __name__ == "__main__":
    main()
