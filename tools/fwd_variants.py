#!/usr/bin/env python
"""Time the forward kernel of several library builds (e.g. different polynomial-exp fractions) back to back."""
import ctypes
import sys

import numpy as np
import torch

B, H, S, D = 8, 32, 4096, 128
q, k, v = (torch.randn(B, H, S, D, device="cuda") for _ in range(3))
o = torch.empty_like(q); l = torch.empty(B, H, S, device="cuda")
P = lambda t: ctypes.c_void_p(t.data_ptr())
libs = [(path, ctypes.CDLL(path)) for path in sys.argv[1:]]
best = {path: 1e9 for path, _ in libs}
for path, lib in libs:
    lib.fa2_profile_enable(1)
    for _ in range(3):
        lib.fa2_forward(P(q), P(k), P(v), P(o), P(l), B, H, S, D, 1, None)
torch.cuda.synchronize()
for rnd in range(6):                       # round-robin so that clock / thermal drift hits every build alike
    for path, lib in libs:
        ms = (ctypes.c_float * 4)(); n = (ctypes.c_int * 4)()
        lib.fa2_profile_read(ms, n)
        ms = (ctypes.c_float * 4)(); n = (ctypes.c_int * 4)()
        for _ in range(5):
            lib.fa2_forward(P(q), P(k), P(v), P(o), P(l), B, H, S, D, 1, None)
        torch.cuda.synchronize()
        lib.fa2_profile_read(ms, n)
        best[path] = min(best[path], ms[1] / n[1])
    print("round", rnd, " ".join(f"{best[p]:.3f}" for p, _ in libs), flush=True)
for path, _ in libs:
    t = best[path]
    print(f"{path.split('/')[-1]:20s} best fwd kernel {t:.3f} ms  {4.0*B*H*S*S*D/t/1e9:.0f} TFLOP/s")
