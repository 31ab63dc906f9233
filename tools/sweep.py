#!/usr/bin/env python
"""BASELINE.json configs[4]: sequence-length sweep S = 512..32768 at D = 64 and 128 with B = max(1, 16384/S),
H = 2048/D (FA-paper convention: 16k tokens, hidden 2048), batch*head sharded over the visible ranks.

    python tools/sweep.py                                             # one GPU
    python -m torch.distributed.run --nproc-per-node N tools/sweep.py   # N GPUs: rank r takes its fa2_partition slab range

Every rank times its share of each point on device-resident fp32 tensors (CUDA events around 10 fused
forward+backward calls, per-kernel spans from the library); the point's time is the max over ranks (gloo carries it:
no GPU collective, the slabs are independent).  Rank 0 prints CSV: kernel-only and with-pre-pass TFLOP/s (aggregate
over the N GPUs), achieved GB/s against the compulsory fp32 bytes for the short, bandwidth-leaning shapes, and a
float64 spot check of rows of O."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-flash-attention_b200"))
import fa2_b200  # noqa: E402

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("gloo")


def all_max(x):
    if dist is None:
        return x
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


lib = fa2_b200.load()
if rank == 0:
    print("n_gpus,D,S,B,H,slabs_per_gpu,fwd_kernel_ms,fwd_total_ms,bwd_kernel_ms,bwd_total_ms,step_ms,fwd_kernel_TF,fwd_total_TF,"
          "bwd_kernel_TF,bwd_total_TF,step_TF,step_TF_per_gpu,fwd_total_GBps_compulsory,bwd_total_GBps_compulsory,max_abs_err_O_rows",
          flush=True)
for D in (64, 128):
    for S in (512, 1024, 2048, 4096, 8192, 16384, 32768):
        B, H = max(1, 16384 // S), 2048 // D
        _, cnt = fa2_b200.partition(B * H, world, rank)
        ms = [0.0] * 4
        step = 0.0
        err = 0.0
        if cnt > 0:
            q, k, v, g = (torch.randn(1, cnt, S, D, device="cuda") for _ in range(4))
            out = (torch.empty_like(q), torch.empty(1, cnt, S, device="cuda"), torch.empty_like(q), torch.empty_like(q), torch.empty_like(q))
            for _ in range(3):
                fa2_b200.forward_backward(q, k, v, g, out=out)
            torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        if cnt > 0:
            lib.fa2_profile_enable(1)
            m = (ctypes.c_float * 4)(); n = (ctypes.c_int * 4)()
            lib.fa2_profile_read(m, n)
            m = (ctypes.c_float * 4)(); n = (ctypes.c_int * 4)()
            iters = 10
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fa2_b200.forward_backward(q, k, v, g, out=out)
            e1.record()
            torch.cuda.synchronize()
            lib.fa2_profile_read(m, n)
            lib.fa2_profile_enable(0)
            ms = [m[i] / iters for i in range(4)]
            step = e0.elapsed_time(e1) / iters
            idx = torch.tensor([0, S // 2, S - 1], device="cuda")
            s_ = (q[0, 0, idx].double() @ k[0, 0].double().T) / D ** 0.5
            err = float((out[0][0, 0, idx].double() - torch.softmax(s_, -1) @ v[0, 0].double()).abs().max())
            del q, k, v, g, out
        cast, fwd, pre, bwd, step, err = (all_max(x) for x in (ms[0], ms[1], ms[2], ms[3], step, err))
        if rank == 0:
            f = 4.0 * B * H * S * S * D
            N = B * H * S
            tf = lambda fl, t: fl / (t * 1e-3) / 1e12 if t > 0 else 0.0
            print(f"{world},{D},{S},{B},{H},{-(-B * H // world)},{fwd:.4f},{fwd + cast:.4f},{bwd:.4f},{bwd + pre + cast:.4f},{step:.4f},"
                  f"{tf(f, fwd):.1f},{tf(f, fwd + cast):.1f},{tf(2.5 * f, bwd):.1f},{tf(2.5 * f, bwd + pre + cast):.1f},{tf(3.5 * f, step):.1f},"
                  f"{tf(3.5 * f, step) / world:.1f},{(16 * N * D + 4 * N) / ((fwd + cast) * 1e-3) / 1e9:.0f},"
                  f"{(32 * N * D + 4 * N) / ((bwd + pre + cast) * 1e-3) / 1e9:.0f},{err:.2e}", flush=True)
if dist is not None:
    dist.destroy_process_group()
