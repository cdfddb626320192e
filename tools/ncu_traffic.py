#!/usr/bin/env python3
"""Turn ncu CSV logs of tools/merge_range.py into profiles/r02_merge_traffic.json, the file bench.py's
`roofline.traffic` is read from (no pasted constants: the entry carries its provenance).

    python tools/ncu_traffic.py --mode range --csv gpurun_out/r02_merge_range_bf16.csv --launches 12 --bm 8 --n 1568 --r 100 --dtype bf16
"""
import argparse
import csv
import json
import os
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--csv", required=True)
ap.add_argument("--mode", default="range", choices=["range", "kernel"])
ap.add_argument("--launches", type=int, required=True, help="kernel launches the figures in the CSV cover (range mode) ")
ap.add_argument("--bm", type=int, default=8)
ap.add_argument("--n", type=int, default=1568)
ap.add_argument("--r", type=int, default=100)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_merge_traffic.json"))
a = ap.parse_args()

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "usecond": 1e3, "nsecond": 1, "msecond": 1e6}
rows = [r for r in csv.reader(l for l in open(a.csv) if l.startswith('"'))]
hdr = rows[0]
ix = {k: hdr.index(k) for k in ("Metric Name", "Metric Unit", "Metric Value")}
tot = {}
for r in rows[1:]:
    name, unit, val = r[ix["Metric Name"]], r[ix["Metric Unit"]], float(r[ix["Metric Value"]].replace(",", ""))
    tot[name] = tot.get(name, 0.0) + val * UNIT.get(unit, 1)
rd, wr = tot.get("dram__bytes_read.sum", 0.0), tot.get("dram__bytes_write.sum", 0.0)
entry = {"bm": a.bm, "n": a.n, "r": a.r, "dtype": a.dtype, "mode": a.mode, "launches": a.launches,
         "dram_bytes_read": rd / a.launches, "dram_bytes_write": wr / a.launches, "dram_bytes_per_launch": (rd + wr) / a.launches,
         "gpu_time_ns_per_launch": tot.get("gpu__time_duration.sum", 0.0) / a.launches,
         "source": os.path.relpath(a.csv, ROOT), "captured": time.strftime("%Y-%m-%d")}
doc = {"entries": []}
if os.path.exists(a.out):
    doc = json.load(open(a.out))
doc["entries"] = [e for e in doc["entries"] if (e["bm"], e["n"], e["r"], e["dtype"]) != (a.bm, a.n, a.r, a.dtype)] + [entry]
json.dump(doc, open(a.out, "w"), indent=1)
print(json.dumps(entry))
