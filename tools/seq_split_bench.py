#!/usr/bin/env python
"""Sequence-split vs slab split through fa2_host_forward_backward on N GPUs: kernel time (max over
devices, incl. the D_i / LSE all-gather and the dQ reduce-scatter over NVLink) and wall time per call, for shapes with
few (b,h) slabs.  usage: seq_split_bench.py N [B H S D ...]"""
import ctypes
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-flash-attention_b200"))
import fa2_b200  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
shapes = [(1, 16, 16384, 128), (1, 4, 16384, 128), (1, 1, 16384, 128), (1, 12, 8192, 128)]
if len(sys.argv) >= 6:
    shapes = [tuple(int(x) for x in sys.argv[2:6])]
lib = fa2_b200.load()
P = lambda t: ctypes.c_void_p(t.data_ptr())
print("B,H,S,D,n_gpus,split(g_bh x g_s),kernel_ms,wall_ms,kernel_TFLOPs,efficiency_vs_1gpu_kernel,max_abs_diff_vs_1gpu")
for (B, H, S, D) in shapes:
    g = torch.Generator().manual_seed(3)
    hq, hk, hv, hdo = (torch.randn(B, H, S, D, generator=g).pin_memory() for _ in range(4))
    outs = [torch.empty(B, H, S, D).pin_memory() for _ in range(4)]
    hl = torch.empty(B, H, S).pin_memory()
    flop = 14.0 * B * H * S * S * D

    def run(n_gpus, reps=3):
        ms = ctypes.c_float(0)
        best_k, best_w = 1e30, 1e30
        for i in range(reps + 1):
            t0 = time.perf_counter()
            fa2_b200._lib.check(lib.fa2_host_forward_backward(P(hq), P(hk), P(hv), P(hdo), P(outs[0]), P(hl), P(outs[1]), P(outs[2]),
                                                              P(outs[3]), B, H, S, D, 1, n_gpus, ctypes.byref(ms)))
            if i:
                best_w = min(best_w, (time.perf_counter() - t0) * 1e3)
                best_k = min(best_k, ms.value)
        return best_k, best_w, [t.clone() for t in outs] + [hl.clone()]

    os.environ.pop("FA2_SEQ_SPLIT", None)
    k1, w1, ref = run(1)
    print(f"{B},{H},{S},{D},1,1x1,{k1:.3f},{w1:.2f},{flop / k1 / 1e9:.0f},1.000,0", flush=True)
    tried = set()
    for force in [None] + [s_ for s_ in (1, 2, 4, 8) if N % s_ == 0 and s_ <= N]:
        if force is None:
            os.environ.pop("FA2_SEQ_SPLIT", None)
        else:
            os.environ["FA2_SEQ_SPLIT"] = str(force)
        gb, gs = fa2_b200.plan_split(B * H, S, N)
        if (gb, gs) in tried or (force is not None and gs != force):
            continue
        tried.add((gb, gs))
        k, w, got = run(N)
        diff = max(float((a - b).abs().max()) for a, b in zip(ref, got))
        tag = "auto" if force is None else "forced"
        print(f"{B},{H},{S},{D},{gb * gs},{gb}x{gs} ({tag}),{k:.3f},{w:.2f},{flop / k / 1e9:.0f},{k1 / (gb * gs * k):.3f},{diff:.2e}", flush=True)
