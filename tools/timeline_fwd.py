#!/usr/bin/env python
"""Debug: CTA 0's per-role timeline of the forward kernel (-DFA2_TIMELINE build, `make timeline`)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, "cuda-flash-attention_b200", "build", "libfa2_b200_tl.so"))
B, H, S, D = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (1, 8, 4096, 128)))
q, k, v = (torch.randn(B, H, S, D, device="cuda") for _ in range(3))
o = torch.empty_like(q); l = torch.empty(B, H, S, device="cuda")
tl = torch.zeros(32 * 32, dtype=torch.int64, device="cuda")
P = lambda t: ctypes.c_void_p(t.data_ptr())
for _ in range(2):
    lib.fa2_forward(P(q), P(k), P(v), P(o), P(l), B, H, S, D, 1, None)
lib.fa2_debug_set_timeline(P(tl))
lib.fa2_forward(P(q), P(k), P(v), P(o), P(l), B, H, S, D, 1, None)
torch.cuda.synchronize()
t = tl.cpu().view(32, 32)
names = {0: "MMA p0a seen -> issue PV0 half0", 1: "MMA p0b seen -> issue PV0 half1 + S0(j+1)", 2: "MMA p1a seen -> issue PV1 half0",
         3: "MMA p1b seen -> issue PV1 half1 + S1(j+1)", 8: "SM0 s_full seen", 9: "SM0 P half0 arrive", 10: "SM0 P half1 arrive",
         12: "SM1 s_full seen", 13: "SM1 P half0 arrive", 14: "SM1 P half1 arrive"}
for j in range(8, 12):
    print(f"--- kv step {j} (period vs previous: {int(t[j,0]-t[j-1,0])} cycles)")
    for c, nme in sorted((int(t[j, s]), names[s]) for s in names if int(t[j, s]) > 0):
        print(f"   {c - int(t[j,0]):7d}  {nme}")
