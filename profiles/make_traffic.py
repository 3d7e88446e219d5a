#!/usr/bin/env python
"""Write profiles/traffic.json from `ncu --set full` captures of the two kernels of the bench workload.

    python profiles/make_traffic.py <compress.ncu-rep> <decompress.ncu-rep> [note]

Records dram__bytes_read.sum + dram__bytes_write.sum of the single captured launch of each kernel, and the SHA-256 of
the kernel sources the capture belongs to.  bench.py reports `roofline.traffic` only while those hashes still match
the sources (a changed kernel needs a new capture); tests/test_abi.py::test_traffic_capture_is_current fails otherwise.
"""
import csv
import hashlib
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CUDA = ROOT / "tamp_b200" / "csrc" / "cuda"
SOURCES = {"compress": ["walk_compress.cu", "tb_device_common.cuh"], "decompress": ["split_decompress.cu", "tb_device_common.cuh"]}


def source_hash(kind):
    h = hashlib.sha256()
    for name in SOURCES[kind]:
        h.update((CUDA / name).read_bytes())
    return h.hexdigest()


def dram_bytes(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(name)
        tot += float(vals[i].replace(",", "")) * scale[units[i]]
    k = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
    t = float(vals[hdr.index("gpu__time_duration.sum")].replace(",", ""))
    return int(tot), (vals[k] if k is not None else ""), t, units[hdr.index("gpu__time_duration.sum")]


if __name__ == "__main__":
    c, cname, ct, cu = dram_bytes(sys.argv[1])
    d, dname, dt, du = dram_bytes(sys.argv[2])
    out = {"note": sys.argv[3] if len(sys.argv) > 3 else "",
           "source": "ncu --set full --clock-control none, one launch each at the bench workload (2^20 x 1 KiB, window=10, extended=0)",
           "compress": {"kernel": cname, "dram_bytes_per_launch": c, "duration": [ct, cu], "capture": Path(sys.argv[1]).name,
                        "sources": SOURCES["compress"], "sources_sha256": source_hash("compress")},
           "decompress": {"kernel": dname, "dram_bytes_per_launch": d, "duration": [dt, du], "capture": Path(sys.argv[2]).name,
                          "sources": SOURCES["decompress"], "sources_sha256": source_hash("decompress")}}
    (ROOT / "profiles" / "traffic.json").write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps(out, indent=1))
