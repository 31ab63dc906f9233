#!/usr/bin/env python
"""Run fa2_host_forward_backward on config C with FA2_HOST_TRACE=1 and print the per-chunk pipeline trace + wall time."""
import ctypes
import os
import sys
import time

os.environ["FA2_HOST_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-flash-attention_b200"))
import torch
import fa2_b200

B, H, S, D = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (8, 32, 4096, 128)))
lib = fa2_b200._lib.load()
n = B * H * S * D
bufs = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(8)]
hl = torch.empty(B * H * S, dtype=torch.float32).pin_memory()
for t in bufs[:4]:
    t.normal_()
P = lambda t: ctypes.c_void_p(t.data_ptr())
hq, hk, hv, hdo, ho, hdq, hdk, hdv = bufs
kms = ctypes.c_float(0)
for i in range(3):
    t0 = time.perf_counter()
    fa2_b200._lib.check(lib.fa2_host_forward_backward(P(hq), P(hk), P(hv), P(hdo), P(ho), P(hl), P(hdq), P(hdk), P(hdv),
                                                      B, H, S, D, 1, 1, ctypes.byref(kms)))
    print("call %d: wall %.2f ms, kernels %.2f ms" % (i, (time.perf_counter() - t0) * 1e3, kms.value), file=sys.stderr)
