#!/usr/bin/env python3
"""profiles/traffic.json from an `ncu -i X.ncu-rep --page raw --csv` dump: per kernel the DRAM bytes per launch
(dram__bytes_read.sum + dram__bytes_write.sum), the ncu duration, the issue-slot utilisation
(sm__inst_issued.avg.pct_of_peak_sustained_active) and the share of stall samples spent at barriers.

    python tools/ncu_traffic_json.py gpurun_out/r2z_raw.csv "profiles/r2_ncu_full_summary.csv (...)" > profiles/traffic.json
"""
import collections
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def num(r, name):
    try:
        return float(r[col[name]].replace(",", ""))
    except (KeyError, ValueError):
        return None


def to_bytes(r, name):
    v = num(r, name)
    if v is None:
        return None
    u = units[col[name]].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def to_us(r, name):
    v = num(r, name)
    u = units[col[name]].lower()
    return None if v is None else v * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "second": 1e6, "nsecond": 1e-3}.get(u, 1)


agg = collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    key = re.sub(r"^void ", "", name).split("(")[0]
    agg[key]["dram"].append((to_bytes(r, "dram__bytes_read.sum") or 0) + (to_bytes(r, "dram__bytes_write.sum") or 0))
    agg[key]["us"].append(to_us(r, "gpu__time_duration.sum"))
    v = num(r, "sm__inst_issued.avg.pct_of_peak_sustained_active")
    if v is not None:
        agg[key]["issue"].append(v / 100.0)
out = {"source": sys.argv[2] if len(sys.argv) > 2 else sys.argv[1], "kernels": {}}
for k, d in agg.items():
    n = len(d["dram"])
    out["kernels"][k] = {"dram_bytes_per_launch": sum(d["dram"]) / n, "ncu_duration_us": sum(d["us"]) / n, "launches": n}
    if d["issue"]:
        out["kernels"][k]["issue_slot_util"] = round(sum(d["issue"]) / len(d["issue"]), 4)
print(json.dumps(out, indent=1))
