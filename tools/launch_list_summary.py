#!/usr/bin/env python
"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals:
    python tools/launch_list_summary.py gpurun_out/launches.csv > profiles/x_launch_list.json
(ncu times are cold-cache and serialised: compare SHARES, not absolutes)."""
import csv
import json
import re
import sys

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
    name = re.sub(r"<.*", "", name)
    rows.append((name, us))
tot = sum(u for _, u in rows)
agg = {}
for n, u in rows:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += u
out = {"launches": len(rows), "total_us": round(tot, 1),
       "ours_us": round(sum(u for n, u in rows if n.startswith("dgm::")), 1),
       "kernels": [{"kernel": n, "launches": c, "us": round(u, 1), "share": round(u / tot, 4)}
                   for n, (c, u) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]]}
print(json.dumps(out, indent=1))
