#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total time and share per kernel."""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(r[iu], 1e-6)
    tot[r[ik]] += v
    cnt[r[ik]] += 1
lib = sum(v for k, v in tot.items() if "fa2" in k)
print(sys.argv[2] if len(sys.argv) > 2 else "")
for k, v in tot.most_common(14):
    name = re.sub(r"\(.*$", "", k)[-62:]
    print(f"{name:62s} launches {cnt[k]:4d}  total {v:9.3f} ms  share of library kernels {100 * v / lib:5.1f} %")
