#!/usr/bin/env python
"""Measure pinned-memory PCIe bandwidth: H2D alone, D2H alone, both directions at once (the e2e path's ceiling)."""
import torch

n = 512 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=4, pieces=1):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_event(a); s2.wait_event(a)
    step = n // pieces
    for _ in range(reps):
        for p in range(pieces):
            sl = slice(p * step, (p + 1) * step)
            if h2d:
                with torch.cuda.stream(s1):
                    d_in[sl].copy_(h_in[sl], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out[sl].copy_(d_out[sl], non_blocking=True)
    e1, e2 = torch.cuda.Event(), torch.cuda.Event()
    e1.record(s1); e2.record(s2)
    torch.cuda.current_stream().wait_event(e1); torch.cuda.current_stream().wait_event(e2)
    b.record()
    torch.cuda.synchronize()
    return reps * n / (a.elapsed_time(b) * 1e-3) / 1e9


for name, h2d, d2h in (("H2D only", 1, 0), ("D2H only", 0, 1), ("both directions", 1, 1)):
    for pieces in (1, 4, 8, 16, 32, 64):
        run(h2d, d2h, 1, pieces)
        print(f"{name:16s} pieces of {n // pieces >> 20:4d} MiB: {run(h2d, d2h, 4, pieces):6.1f} GB/s per direction")
