#!/usr/bin/env python
"""Print the hottest SASS instructions (warp-stall samples) of one kernel from `ncu --page source --csv` output."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(r for r in rows if r and r[0] == "Address")
data = []
for r in rows[rows.index(hdr) + 1:]:
    if len(r) != len(hdr) or r[0] == "Address":
        break
    data.append(r)
i_src, i_samp = hdr.index("Source"), hdr.index("# Samples")
stall_cols = [(j, h) for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(d[i_samp]) for d in data)
print("total samples", tot, "instructions", len(data))
agg = {}
for d in data:
    for j, h in stall_cols:
        agg[h] = agg.get(h, 0) + int(d[j])
print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
for k, d in sorted(enumerate(data), key=lambda t: -int(t[1][i_samp]))[:n]:
    st = sorted(((int(d[j]), h) for j, h in stall_cols), reverse=True)[:2]
    print(f"{k:5d} {int(d[i_samp]):7d} {100 * int(d[i_samp]) / tot:5.1f}%  {d[i_src].strip()[:64]:64s} {st}")
