#!/usr/bin/env python
"""ncu CSV (--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:<kernels>) -> the JSON that
bench.py's `roofline.traffic` reads: DRAM bytes per launch of the dominant kernel family over the LAST `--launches` profiled
launches (= one denoising step).   usage: ncu_dram_to_json.py in.csv out.json --launches N [--source "..."]"""
import argparse
import csv
import json

ap = argparse.ArgumentParser()
ap.add_argument("csv_in")
ap.add_argument("json_out")
ap.add_argument("--launches", type=int, required=True)
ap.add_argument("--source", default="")
a = ap.parse_args()
rows = []
with open(a.csv_in, newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    rows.append(r)
per = {}
for r in rows:
    d = per.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"].lower()
    mul = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1)
    d[r["Metric Name"]] = v * mul
ids = sorted(per)[-a.launches:]
rd = sum(per[i].get("dram__bytes_read.sum", 0.0) for i in ids)
wr = sum(per[i].get("dram__bytes_write.sum", 0.0) for i in ids)
ms = sum(per[i].get("gpu__time_duration.sum", 0.0) for i in ids)
names = {}
for i in ids:
    names[per[i]["name"][:60]] = names.get(per[i]["name"][:60], 0) + 1
out = {"kernel": "conv3x3_halo_t_kernel family (all geometries)", "launches": len(ids), "dram_bytes_read": rd,
       "dram_bytes_write": wr, "dram_bytes_per_launch": (rd + wr) / max(1, len(ids)), "ncu_time_ms": ms, "by_kernel": names,
       "profiled_launches_total": len(per), "source": a.source}
json.dump(out, open(a.json_out, "w"), indent=1)
print(json.dumps(out)[:400])
