#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: top SASS lines by stall samples, with reasons."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
tot = 0
for r in rows[2:]:
    try:
        s = int(r[ix["# Samples"]])
    except (ValueError, IndexError):
        continue
    tot += s
    data.append((s, r))
print("total samples", tot)
agg = {}
for s, r in data:
    for c in stall_cols:
        try:
            agg[c] = agg.get(c, 0) + int(r[ix[c]])
        except ValueError:
            pass
print("by reason:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:10])
for s, r in sorted(data, key=lambda x: -x[0])[:n]:
    reasons = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols if r[ix[c]] not in ("", "0")), reverse=True)[:3]
    print(f"{s:7d} {100.0 * s / tot:5.1f}%  {r[ix['Address']][-5:]}  {r[ix['Source']][:90]:90s} {reasons}")
