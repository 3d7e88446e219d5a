#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:  kernel, launches, total ms, share."""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    tot[r[ik]] += float(r[iv].replace(",", "")) * scale.get(r[iu], 1e-6)
    cnt[r[ik]] += 1
total = sum(tot.values())
w = csv.writer(sys.stdout)
w.writerow(["kernel", "launches", "total_ms", "share"])
for k in sorted(tot, key=tot.get, reverse=True):
    w.writerow([k, cnt[k], f"{tot[k]:.3f}", f"{tot[k] / total:.4f}"])
