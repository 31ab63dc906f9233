#!/usr/bin/env python
"""Debug: CTA 0's per-role timeline of the forward kernel (-DFA2_TIMELINE build, `make timeline`)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.environ.get("FA2_TL_LIB") or os.path.join(ROOT, "cuda-flash-attention_b200", "build", "libfa2_b200_tl.so"))
B, H, S, D = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (1, 8, 4096, 128)))
q, k, v = (torch.randn(B, H, S, D, device="cuda") for _ in range(3))
o = torch.empty_like(q); l = torch.empty(B, H, S, device="cuda")
n_cta = B * H * ((S + 255) // 256)
tl = torch.zeros(32 * 32 + 8 * n_cta, dtype=torch.int64, device="cuda")
P = lambda t: ctypes.c_void_p(t.data_ptr())
for _ in range(2):
    lib.fa2_forward(P(q), P(k), P(v), P(o), P(l), B, H, S, D, 1, None)
lib.fa2_debug_set_timeline(P(tl))
if len(sys.argv) > 5 and sys.argv[5] == "fused":      # the fused forward+backward variant of the forward kernel
    os.environ["FA2_TL_FWD_ONLY"] = "1"
    g = torch.randn_like(q); dq, dk, dv = (torch.empty_like(q) for _ in range(3))
    lib.fa2_forward_backward(P(q), P(k), P(v), P(g), P(o), P(l), P(dq), P(dk), P(dv), B, H, S, D, 1, None)
else:
    lib.fa2_forward(P(q), P(k), P(v), P(o), P(l), B, H, S, D, 1, None)
torch.cuda.synchronize()
life = tl[1024:].cpu().view(n_cta, 8)
t = tl[:1024].cpu().view(32, 32)
names = {0: "MMA p0a seen -> issue PV0 half0", 1: "MMA p0b seen -> issue PV0 half1 + S0(j+1)", 2: "MMA p1a seen -> issue PV1 half0",
         3: "MMA p1b seen -> issue PV1 half1 + S1(j+1)", 8: "SM0 s_full seen", 9: "SM0 P half0 arrive", 10: "SM0 P half1 arrive",
         12: "SM1 s_full seen", 13: "SM1 P half0 arrive", 14: "SM1 P half1 arrive"}
for j in range(8, 12):
    print(f"--- kv step {j} (period vs previous: {int(t[j,0]-t[j-1,0])} cycles)")
    for c, nme in sorted((int(t[j, s]), names[s]) for s in names if int(t[j, s]) > 0):
        print(f"   {c - int(t[j,0]):7d}  {nme}")

# per-CTA lifetimes: how much of an SM's time is prologue / steady KV loop / epilogue / gap between CTAs
import collections
by_sm = collections.defaultdict(list)
for row in life.tolist():
    by_sm[row[4]].append(row)
pro, loop, epi, gap = [], [], [], []
for sm, rows in by_sm.items():
    rows.sort()
    for i, r in enumerate(rows):
        pro.append(r[1] - r[0]); loop.append(r[2] - r[1]); epi.append(r[3] - r[2])
        if i:
            gap.append(r[0] - rows[i - 1][3])
med = lambda x: sorted(x)[len(x) // 2] if x else 0
if any(r[5] for r in life.tolist()):
    rows = life.tolist()
    print("epilogue chunks (cycles after last O seen): store 0 issued %d, store 1 %d, store 2 %d, done %d" % tuple(
        med([r[k] - r[2] for r in rows]) for k in (5, 6, 7, 3)))
print(f"CTAs {n_cta} on {len(by_sm)} SMs: median cycles  prologue(start->first S) {med(pro)}  KV loop {med(loop)}  "
      f"epilogue(last O->end) {med(epi)}  gap between CTAs on an SM {med(gap)}")
tot = [rows[-1][3] - rows[0][0] for rows in by_sm.values()]
print(f"per-SM busy span: median {med(tot)} cycles; sum of loop parts / span = {sum(loop) / sum(tot):.3f}")
