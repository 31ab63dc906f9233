#!/usr/bin/env python
"""BASELINE.json configs[4]: sequence-length sweep S = 512..32768 at D = 64 and 128 with B = max(1, 16384/S),
H = 2048/D (FA-paper convention: 16k tokens, hidden 2048).  Kernel-only and with-pre-pass TFLOP/s per point,
plus achieved HBM GB/s against the compulsory fp32 bytes for the short, bandwidth-leaning shapes.  CSV to stdout."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-flash-attention_b200"))
import fa2_b200  # noqa: E402

lib = fa2_b200.load()
print("D,S,B,H,fwd_kernel_ms,fwd_total_ms,bwd_kernel_ms,bwd_total_ms,fwd_kernel_TF,fwd_total_TF,bwd_kernel_TF,bwd_total_TF,"
      "fwd_total_GBps_compulsory,bwd_total_GBps_compulsory,max_abs_err_O_rows")
for D in (64, 128):
    for S in (512, 1024, 2048, 4096, 8192, 16384, 32768):
        B, H = max(1, 16384 // S), 2048 // D
        q, k, v, g = (torch.randn(B, H, S, D, device="cuda") for _ in range(4))
        out = (torch.empty_like(q), torch.empty(B, H, S, device="cuda"), torch.empty_like(q), torch.empty_like(q), torch.empty_like(q))
        for _ in range(3):
            fa2_b200.forward_backward(q, k, v, g, out=out)
        torch.cuda.synchronize()
        lib.fa2_profile_enable(1)
        ms = (ctypes.c_float * 4)(); n = (ctypes.c_int * 4)()
        lib.fa2_profile_read(ms, n)
        ms = (ctypes.c_float * 4)(); n = (ctypes.c_int * 4)()
        iters = 10
        for _ in range(iters):
            fa2_b200.forward_backward(q, k, v, g, out=out)
        torch.cuda.synchronize()
        lib.fa2_profile_read(ms, n)
        lib.fa2_profile_enable(0)
        cast, fwd, pre, bwd = (ms[i] / iters for i in range(4))
        f = 4.0 * B * H * S * S * D
        N = B * H * S
        idx = torch.tensor([0, S // 2, S - 1], device="cuda")
        s_ = (q[0, 0, idx].double() @ k[0, 0].double().T) / D ** 0.5
        err = float((out[0][0, 0, idx].double() - torch.softmax(s_, -1) @ v[0, 0].double()).abs().max())
        tf = lambda fl, t: fl / (t * 1e-3) / 1e12
        print(f"{D},{S},{B},{H},{fwd:.4f},{fwd + cast:.4f},{bwd:.4f},{bwd + pre + cast:.4f},{tf(f, fwd):.1f},{tf(f, fwd + cast):.1f},"
              f"{tf(2.5 * f, bwd):.1f},{tf(2.5 * f, bwd + pre + cast):.1f},{(16 * N * D + 4 * N) / ((fwd + cast) * 1e-3) / 1e9:.0f},"
              f"{(32 * N * D + 4 * N) / ((bwd + pre + cast) * 1e-3) / 1e9:.0f},{err:.2e}", flush=True)
