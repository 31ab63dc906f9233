#!/usr/bin/env python
"""Quick GPU-side diagnostic: per-shape max-abs errors (vs fp64 truth) and device timings of the
forward / backward entry points.  Writes one line per case; not a test and not the bench."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cuda-flash-attention_b200"))
import fa2_b200  # noqa: E402
from oracle import fa2_oracle as orc  # noqa: E402


def timed(fn, iters=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bwd", action="store_true")
    ap.add_argument("--big", action="store_true")
    a = ap.parse_args()
    shapes = [(1, 1, 128, 64), (1, 1, 128, 128), (1, 2, 256, 64), (1, 2, 100, 64), (1, 1, 384, 128),
              (2, 2, 512, 32), (1, 2, 1000, 128)]
    for shp in shapes:
        rng = np.random.default_rng(0)
        Q, K, V, dO = (rng.standard_normal(shp).astype(np.float32) for _ in range(4))
        t = orc.attention_fp64(Q, K, V, dO)
        q, k, v, g = (torch.from_numpy(x).cuda() for x in (Q, K, V, dO))
        O, L = fa2_b200.forward(q, k, v)
        torch.cuda.synchronize()
        line = f"shape {shp}: O err {np.abs(O.cpu().numpy() - t[0]).max():.3e}  LSE err {np.abs(L.cpu().numpy() - t[1]).max():.3e}"
        if a.bwd:
            dQ, dK, dV = fa2_b200.backward(q, k, v, O, g, L)
            torch.cuda.synchronize()
            line += "  dQ %.3e dK %.3e dV %.3e" % tuple(
                np.abs(x.cpu().numpy() - r).max() for x, r in zip((dQ, dK, dV), t[2:]))
        print(line, flush=True)
    if a.big:
        for (B, H, S, D) in [(4, 16, 1024, 64), (8, 32, 4096, 128), (1, 16, 16384, 128)]:
            q, k, v, g = (torch.randn(B, H, S, D, device="cuda") for _ in range(4))
            out = (torch.empty_like(q), torch.empty(B, H, S, device="cuda"))
            ms = timed(lambda: fa2_b200.forward(q, k, v, out=out))
            fl = 4.0 * B * H * S * S * D
            line = f"B{B} H{H} S{S} D{D}: fwd {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s (incl. cast pre-pass)"
            if a.bwd:
                O, L = out
                outb = tuple(torch.empty_like(q) for _ in range(3))
                msb = timed(lambda: fa2_b200.backward(q, k, v, O, g, L, out=outb), iters=5)
                line += f" | bwd {msb:.3f} ms  {2.5 * fl / msb / 1e9:.1f} TFLOP/s"
            print(line, flush=True)


if __name__ == "__main__":
    main()
