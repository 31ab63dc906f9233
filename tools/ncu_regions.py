#!/usr/bin/env python
"""Split an `ncu --page source --csv` dump into regions delimited by mbarrier waits / named barriers and
print the sample count of each region in address order (a poor man's per-phase timeline)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
region, acc, n_inst, start = "entry", 0, 0, data[0][ix["Address"]][-5:]
mufu = 0
for r in data:
    src = r[ix["Source"]].strip()
    s = int(r[ix["# Samples"]] or 0)
    if "SYNCS.PHASECHK" in src or "BAR.SYNC" in src or "EXIT" in src or "UTCBAR" in src and False:
        if acc > tot * 0.004:
            print(f"{start}  {100.0 * acc / tot:5.1f}%  insts {n_inst:4d} mufu {mufu:3d}  before: {src[:70]}")
        acc, n_inst, mufu, start = 0, 0, 0, r[ix["Address"]][-5:]
    acc += s
    n_inst += 1
    mufu += "MUFU" in src
print("total", tot)
