#!/usr/bin/env python
"""Debug: run the -DFA2_TIMELINE build of the backward kernel and print CTA 0's per-role timeline
(clock64 stamps relative to the start of each iteration).  `make -C cuda-flash-attention_b200 timeline` first."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.environ.get("FA2_TL_LIB") or os.path.join(ROOT, "cuda-flash-attention_b200", "build", "libfa2_b200_tl.so"))
B, H, S, D = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (1, 8, 4096, 128)))
q, k, v, g = (torch.randn(B, H, S, D, device="cuda") for _ in range(4))
o = torch.empty_like(q); l = torch.empty(B, H, S, device="cuda")
dq, dk, dv = (torch.empty_like(q) for _ in range(3))
n_cta = B * H * ((S + 127) // 128)
tl = torch.zeros(32 * 32 + 8 * n_cta, dtype=torch.int64, device="cuda")
vp = ctypes.c_void_p
P = lambda t: vp(t.data_ptr())
lib.fa2_forward(P(q), P(k), P(v), P(o), P(l), B, H, S, D, 1, None)
for _ in range(2):
    lib.fa2_backward(P(q), P(k), P(v), P(o), P(g), P(l), P(dq), P(dk), P(dv), B, H, S, D, 1, None)
lib.fa2_debug_set_timeline(P(tl))
lib.fa2_backward(P(q), P(k), P(v), P(o), P(g), P(l), P(dq), P(dk), P(dv), B, H, S, D, 1, None)
torch.cuda.synchronize()
life = tl[1024:].cpu().view(n_cta, 8)
t = tl[:1024].cpu().view(32, 32)
names = {0: "MMA issue S", 1: "MMA ds_full(i-1) seen", 2: "MMA dq_empty(i-1) seen", 3: "MMA issue dP", 4: "MMA p_full seen/issue dV",
         5: "MMA peer atoms landed / issue dQ(i-1) [pair kernel]", 6: "MMA dK(i-1) issued [pair kernel]",
         8: "C  s_full seen", 9: "C  p_full arrive", 10: "C  dp_full seen", 11: "C  ds_full arrive",
         12: "C  dS math starts [pair kernel]",
         20: "    follower C s_full seen (follower clock)", 21: "    follower C p_full arrive", 22: "    follower C dp_full seen",
         23: "    follower C ds_full arrive", 24: "C  warpgroup 1 s_full seen", 25: "C  warpgroup 1 p_full arrive",
         26: "    follower C warpgroup 1 s_full seen", 27: "    follower C warpgroup 1 p_full arrive",
         15: "Dr dq_full seen", 16: "Dr dq_empty arrive", 17: "Dr staging done",
         18: "Dr chunk0 issued", 19: "Dr chunk1 issued"}
base = int(t[8, 0])
n_it = min(32, (S + 127) // 128)
for i in range(8, min(n_it, 13)):
    print(f"--- iteration {i} (period vs previous: {int(t[i,0]-t[i-1,0])} cycles)")
    ev = sorted((int(t[i, s]), names[s]) for s in names if int(t[i, s]) > 0)
    for c, nme in ev:
        print(f"   {c - int(t[i,0]):7d}  {nme}")

# per-CTA (work item) lifetimes: prologue / Q loop / epilogue / gap between consecutive items on an SM
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
print(f"work items {n_cta} on {len(by_sm)} SMs: median cycles  prologue(start->first S) {med(pro)}  Q loop {med(loop)}  "
      f"epilogue(dK/dV complete->end) {med(epi)}  gap between items on an SM {med(gap)}")
tot = [rows[-1][3] - rows[0][0] for rows in by_sm.values()]
print(f"per-SM busy span: median {med(tot)} cycles; sum of loop parts / span = {sum(loop) / sum(tot):.3f}")
