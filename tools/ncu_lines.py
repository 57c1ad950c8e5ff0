#!/usr/bin/env python3
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur_file = None
agg = collections.Counter(); stall = collections.defaultdict(collections.Counter); text = {}
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; si = hdr.index("# Samples"); sc = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0].isdigit():      # a CUDA source line row
        line = (cur_file, int(r[0])); text[line] = r[1].strip()
        if r[si].isdigit():
            agg[line] += 0  # counted via its SASS rows below? (source rows already carry the sum)
            agg[line] = int(r[si])
            for i in sc:
                if r[i].isdigit(): stall[line][hdr[i]] = int(r[i])
T = sum(agg.values())
print("total samples", T)
for line, n in agg.most_common(top_n):
    print(f"{n:6d} {100*n/max(T,1):5.1f}%  {line[0]}:{line[1]:<5d} {text[line][:90]:90s} {stall[line].most_common(2)}")
