#!/usr/bin/env python
"""CTA-pair backward (FA2_BWD_PAIR=1) against float64 on D = 128 shapes, then timing of both backward kernels."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cuda-flash-attention_b200"))
import fa2_b200  # noqa: E402
from oracle import fa2_oracle as orc  # noqa: E402

print("FA2_BWD_PAIR =", os.environ.get("FA2_BWD_PAIR"), flush=True)
shapes = [(1, 1, 128, 128), (1, 1, 256, 128), (1, 2, 384, 128), (1, 2, 100, 128), (2, 3, 1000, 128), (1, 40, 1536, 128)]
if len(sys.argv) > 1 and sys.argv[1] == "time":
    shapes = []
for shp in shapes:
    rng = np.random.default_rng(0)
    Q, K, V, dO = (rng.standard_normal(shp).astype(np.float32) for _ in range(4))
    t = orc.attention_fp64(Q, K, V, dO)
    q, k, v, g = (torch.from_numpy(x).cuda() for x in (Q, K, V, dO))
    o, l = (torch.from_numpy(np.ascontiguousarray(x, np.float32)).cuda() for x in t[:2])
    dQ, dK, dV = fa2_b200.backward(q, k, v, o, g, l)
    torch.cuda.synchronize()
    print("shape %s: dQ %.3e dK %.3e dV %.3e" % ((shp,) + tuple(np.abs(x.cpu().numpy() - r).max() for x, r in zip((dQ, dK, dV), t[2:]))), flush=True)
lib = fa2_b200.load()
for (B, H, S, D) in [(8, 32, 4096, 128), (1, 16, 16384, 128)]:
    q, k, v, g = (torch.randn(B, H, S, D, device="cuda") for _ in range(4))
    out = (torch.empty_like(q), torch.empty(B, H, S, device="cuda"), torch.empty_like(q), torch.empty_like(q), torch.empty_like(q))
    for _ in range(3):
        fa2_b200.forward_backward(q, k, v, g, out=out)
    torch.cuda.synchronize()
    lib.fa2_profile_enable(1)
    ms = (ctypes.c_float * 4)(); n = (ctypes.c_int * 4)()
    lib.fa2_profile_read(ms, n)
    ms = (ctypes.c_float * 4)(); n = (ctypes.c_int * 4)()
    for _ in range(10):
        fa2_b200.forward_backward(q, k, v, g, out=out)
    torch.cuda.synchronize()
    lib.fa2_profile_read(ms, n)
    lib.fa2_profile_enable(0)
    fl = 10.0 * B * H * S * S * D
    qd, kd, vd, gd = (t_[0, 0].double() for t_ in (q, k, v, g))
    s_ = qd @ kd.T / D ** 0.5
    p_ = torch.softmax(s_, -1)
    o_ = p_ @ vd
    ds = p_ * (gd @ vd.T - (gd * o_).sum(-1, keepdim=True)) / D ** 0.5
    errs = [float((out[2][0, 0].double() - ds @ kd).abs().max()), float((out[3][0, 0].double() - ds.T @ qd).abs().max()),
            float((out[4][0, 0].double() - p_.T @ gd).abs().max())]
    print(f"B{B} H{H} S{S} D{D}: bwd kernel {ms[3] / 10:.3f} ms = {fl / (ms[3] / 10) / 1e9:.0f} TFLOP/s, fwd {ms[1] / 10:.3f} ms; "
          f"slab 0 errors dQ {errs[0]:.2e} dK {errs[1]:.2e} dV {errs[2]:.2e}", flush=True)
