#!/usr/bin/env python
"""Run a few forward+backward steps of one shape (target program for ncu captures)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-flash-attention_b200"))
import fa2_b200  # noqa: E402

B, H, S, D = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (2, 32, 4096, 128)))
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
q, k, v, g = (torch.randn(B, H, S, D, device="cuda") for _ in range(4))
for _ in range(steps):
    out = fa2_b200.forward_backward(q, k, v, g)
torch.cuda.synchronize()
print("ok", float(out[0].abs().mean()))
