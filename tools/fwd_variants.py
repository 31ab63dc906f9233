#!/usr/bin/env python
"""Time the forward kernel of several library builds (tools/build_variant.sh) back to back.
usage: fwd_variants.py B H S D lib1.so lib2.so ...   (order alternates per round; mean over rounds: the box is
power-capped, so whoever runs first after an idle gap looks faster)"""
import ctypes
import sys

import torch

B, H, S, D = (int(x) for x in sys.argv[1:5])
paths = sys.argv[5:]
q, k, v = (torch.randn(B, H, S, D, device="cuda") for _ in range(3))
o = torch.empty_like(q); l = torch.empty(B, H, S, device="cuda")
P = lambda t: ctypes.c_void_p(t.data_ptr())
libs = [(path, ctypes.CDLL(path)) for path in paths]


def run(lib, n):
    for _ in range(n):
        lib.fa2_forward(P(q), P(k), P(v), P(o), P(l), B, H, S, D, 1, None)
    torch.cuda.synchronize()


acc = {p: [0.0, 0] for p in paths}
for path, lib in libs:
    lib.fa2_profile_enable(1)
    run(lib, 3)
reps = max(5, int(40e-3 / (4.0 * B * H * S * S * D / 1e15)))     # ~40 ms of work per measurement
for rnd in range(6):
    for path, lib in (libs if rnd % 2 == 0 else libs[::-1]):
        ms = (ctypes.c_float * 4)(); n = (ctypes.c_int * 4)()
        run(lib, max(2, reps // 4))
        lib.fa2_profile_read(ms, n)
        ms = (ctypes.c_float * 4)(); n = (ctypes.c_int * 4)()
        run(lib, reps)
        lib.fa2_profile_read(ms, n)
        acc[path][0] += ms[1] / n[1]; acc[path][1] += 1
for path in paths:
    t = acc[path][0] / acc[path][1]
    print(f"{path.split('/')[-1]:20s} mean fwd kernel {t:.4f} ms  {4.0*B*H*S*S*D/t/1e9:.0f} TFLOP/s")
