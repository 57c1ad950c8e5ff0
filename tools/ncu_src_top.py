#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` dump: stall-reason totals and the hottest SASS lines."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = next(r for r in rows if "Source" in r and "# Samples" in r)
si, src = hdr.index("# Samples"), hdr.index("Source")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = {hdr[i]: 0 for i in stall_cols}
data = []
for r in rows:
    if len(r) < len(hdr) or r is hdr or not r[si].isdigit():
        continue
    n = int(r[si]); data.append((n, r[src].strip(), r))
    for i in stall_cols:
        tot[hdr[i]] += int(r[i] or 0)
T = sum(n for n, _, _ in data)
print("total samples", T)
print("stall totals:", [(k, v, f"{100*v/max(T,1):.1f}%") for k, v in sorted(tot.items(), key=lambda x: -x[1])[:8]])
for n, s, r in sorted(data, key=lambda x: -x[0])[:top_n]:
    top = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{n:7d} {100*n/max(T,1):5.1f}%  {s[:72]:72s} {top}")
