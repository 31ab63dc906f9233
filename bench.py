#!/usr/bin/env python
"""bench.py -- headline benchmark of the FA2 forward+backward path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload C|A|B|D]
    (N > 1: python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...)

A step = one forward + backward pass (cast, forward [+ fused backward pre-pass], backward) over one
synthetic batch of the workload, on fp32 [B,H,S,D] device tensors.  Default workload = BASELINE.json
configs[2], B8 H32 S4096 D128 (the shape the metric is quoted on); with N ranks every rank runs that
shape on its own slab range of a global batch 8*N (weak scaling, no data-path collective -- (batch,
head) slabs are independent).  After the timed loop the TIMED outputs are checked against a float64
reference on sampled (b,h) slabs; a wrong answer makes the run exit non-zero.  At N > 1 rank 0 also
runs ONE workload-sized problem through fa2_host_forward_backward(n_gpus=N) -- the north_star
partitioner -- and reports its strong scaling against n_gpus=1.
Rank 0 prints ONE JSON line.  FLOP convention: fwd 4*BHS^2*D, bwd 10*BHS^2*D.
"""
import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cuda-flash-attention_b200"))

WORKLOADS = {   # BASELINE.json configs
    "A": (2, 8, 512, 64), "B": (4, 16, 1024, 64), "C": (8, 32, 4096, 128), "D": (1, 16, 16384, 128),
}
PRECISION_OF = {"A": "fp32", "B": "fp16", "C": "fp32", "D": "fp32"}      # configs[1] names the fp16 SHM flag
METRIC = "FA2 fwd+bwd TFLOPS/GPU-aggregate, B8 H32 S4096 D128 per GPU"
UNIT = "TFLOP/s"
TOL = {"O": 1e-2, "LSE": 1e-3, "dQ": 1e-2, "dK": 1e-2, "dV": 1e-2}        # north_star, 16-bit operands


def flops(B, H, S, D):
    f = 4.0 * B * H * S * S * D
    return f, 2.5 * f


def compulsory_bytes(B, H, S, D):
    n = B * H * S
    return 16 * n * D + 4 * n, 32 * n * D + 4 * n          # fwd, bwd


def workload_name(key, B, H, S, D):
    return f"configs[{'ABCD'.index(key)}] B{B} H{H} S{S} D{D} fwd+bwd"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0=None, t1=None):
        """Summarise the samples that arrived inside [t0, t1] (host clock around the timed region).  The sampler
        is started before the warm-up steps (same workload), so a very short timed region that caught no sample
        of its own falls back to the samples taken under the warm-up load and says so."""
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        window = "timed region"
        rows = [r for t, r in self.rows if t0 is None or (t0 <= t <= t1 + 0.05)]
        if not rows and self.rows:
            rows = [r for t, r in self.rows if t >= t0 - 1.0] or [r for _, r in self.rows[-3:]]
            window = "warm-up + timed region (timed region shorter than one sampling period)"
        sm, mx, reasons, pw = [], [], set(), []
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def cpu_baseline(B, H, S, D, budget_s=12.0):
    """PyTorch CPU scaled_dot_product_attention fwd+bwd on the box's host cores (north_star's named
    CPU path; the reference harness's CPU oracle is the same maths, test_flash_attention2.py:197-232),
    on a bounded sample of (b,h) slabs of the workload."""
    import torch
    import torch.nn.functional as F
    torch.manual_seed(0)
    heads = 1

    def run(h):
        q, k, v = (torch.randn(1, h, S, D, requires_grad=True) for _ in range(3))
        t0 = time.perf_counter()
        o = F.scaled_dot_product_attention(q, k, v)
        o.backward(torch.ones_like(o))
        return time.perf_counter() - t0

    run(1)                                  # warm-up
    t1 = run(1)
    heads = int(max(1, min(B * H, budget_s / max(t1, 1e-3))))
    t = run(heads)
    f, b = flops(1, heads, S, D)
    return {"value": (f + b) / t / 1e12, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "host_cpus": os.cpu_count(),
            "sample": f"torch {torch.__version__} CPU F.scaled_dot_product_attention fwd+bwd (dO=1) on {heads} of "
                      f"{B * H} (b,h) slabs of S{S} D{D}, {t:.2f} s, 1 warm-up"}


def run_reference(args):
    """--impl reference: the reference's OWN fa2 fp32 kernels and host launchers (kernel_fa2_optimized.cu:351-423,
    f-attn2-backward.cu:384-485), compiled from where they lie in /root/reference into oracle/_ref/ref_any_d
    (oracle/build_ref.sh), run on the SAME workload as the b200 arm.  The reference's dispatcher refuses D=128
    (include/dispatcher.h:226-227), so ref_any_d instantiates its templates at the workload's head dim; nothing else
    is changed.  The reference has no CPU implementation of this path: its kernels run on the B200's CUDA cores.
    `value` = its TimerGPU kernel time (src/main.cpp:107 semantics), `e2e` = wall clock of its host launchers
    (pageable host buffers, cudaMalloc / H2D / D2H / cudaFree inside, as RunFlashAttention does)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_any_d")
    key = args.workload
    B, H, S, D = WORKLOADS[key]
    if not os.path.exists(exe):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_any_d not built (run __graft_entry__.build() where /root/reference exists)"}))
        return
    budget = float(os.environ.get("FA2_REF_BUDGET_S", "150"))
    try:
        out = subprocess.run([exe, "--bench", str(B), str(H), str(S), str(D), str(args.steps), str(args.warmup), str(budget)],
                             capture_output=True, text=True, timeout=1500)
    except subprocess.TimeoutExpired:
        print(json.dumps({"impl": "reference", "unavailable": "reference bench timed out"}))
        return
    m = re.search(r"^REF_BENCH (\{.*\})$", out.stdout, re.M)
    if out.returncode != 0 or not m:
        print(json.dumps({"impl": "reference", "unavailable": "reference bench failed: " + (out.stderr or out.stdout)[-300:].replace("\n", " ")}))
        return
    r = json.loads(m.group(1))
    full = r["slabs_per_step"] == r["slabs_full"]
    sample = (f"{r['steps']} timed steps of {r['slabs_per_step']} of {r['slabs_full']} (b,h) slabs of S{S} D{D} "
              f"({'the full workload' if full else 'bounded slab subset, same S and D'}); reference fa2 fp32 kernels on the "
              f"B200's CUDA cores via its own host launchers; total wall {r['total_wall_s']:.0f} s")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": r["kernel_tflops"], "unit": UNIT, "n_gpus": 1, "steps": r["steps"],
        "warmup": args.warmup, "ms_per_step": r["kernel_ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(key, B, H, S, D) + " per GPU",
                   "reference_sample": sample,
                   "timer": "reference TimerGPU total (fwd kernel + D kernel + bwd kernel), include/timer.h:50-64"},
        "cpu_baseline": {"value": r["kernel_tflops"], "unit": UNIT, "cores": 1, "kind": "reference", "sample": sample},
        "e2e": {"value": r["e2e_tflops"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "ms_per_step": r["wall_ms_per_step"],
                "note": "wall clock of host_flash_attention2_forward + _backward on pageable host buffers (their own "
                        "cudaMalloc/H2D/D2H/cudaFree inside)"},
    }))


def fp64_slab(q, k, v, g):
    """float64 O, LSE, dQ, dK, dV of one (b,h) slab on the GPU (checker only)."""
    import torch
    D = q.shape[-1]
    q, k, v, g = (t.double() for t in (q, k, v, g))
    s = (q @ k.T) / (D ** 0.5)
    lse = torch.logsumexp(s, -1)
    p = torch.exp(s - lse[:, None])
    del s
    o = p @ v
    dv = p.T @ g
    dp = g @ v.T
    ds = p * (dp - (g * o).sum(-1, keepdim=True)) / (D ** 0.5)
    del p, dp
    return o, lse, ds @ k, ds.T @ q, dv


def verify_outputs(q, k, v, g, outs, n_each=3):
    """max-abs error of the given outputs against float64 on (b,h) slabs at the first, middle and last work items."""
    import torch
    B, H = q.shape[:2]
    BH = B * H
    picks = sorted(set(x for x in (list(range(n_each)) + list(range(BH // 2 - 1, BH // 2 - 1 + n_each)) +
                                   list(range(BH - n_each, BH))) if 0 <= x < BH))
    names = ("O", "LSE", "dQ", "dK", "dV")
    err = {n: 0.0 for n in names}
    finite = True
    for bh in picks:
        b, h = divmod(bh, H)
        want = fp64_slab(q[b, h], k[b, h], v[b, h], g[b, h])
        for n, got, w in zip(names, outs, want):
            x = got[b, h]
            finite = finite and bool(torch.isfinite(x).all())
            err[n] = max(err[n], float((x.double() - w).abs().max()))
        del want
    return picks, err, finite


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C", choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg at N > 1")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import ctypes
    import numpy as np
    import torch
    import fa2_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the FA2 path has no CPU fallback")
    torch.cuda.set_device(local)
    # rank 0 samples its GPU's clocks; started here, long before the timed region, because nvidia-smi needs a while
    # to deliver its first sample (more so with eight ranks starting at once)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dist = None
    host_group = None
    if world > 1:
        # rank 0 must print exactly one JSON line on stdout, and NCCL prints its version banner there whenever
        # NCCL_DEBUG is set (the GPU image sets it): create the communicator with fd 1 pointed at stderr.
        import torch.distributed as dist
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
            # host-side barrier (gloo): ranks parked on it leave their GPU idle, unlike an NCCL barrier kernel
            host_group = dist.new_group(backend="gloo")
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    if args.gpus != world and rank == 0 and world > 1:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    key = args.workload
    B, H, S, D = WORKLOADS[key]
    prec = PRECISION_OF[key]
    BH_global = B * H * world
    bh0, cnt = fa2_b200.partition(BH_global, world, rank)        # this rank's slab range of the global batch
    assert cnt == B * H
    g = torch.Generator(device="cuda").manual_seed(1234 + bh0)
    q, k, v, do = (torch.randn(B, H, S, D, device="cuda", generator=g) for _ in range(4))
    outs = (torch.empty_like(q), torch.empty(B, H, S, device="cuda"), torch.empty_like(q), torch.empty_like(q), torch.empty_like(q))
    lib = fa2_b200.load()
    st = torch.cuda.current_stream()
    n = B * H * S * D
    # small workloads fit in the 126 MB L2: flush it between timed iterations (outside the per-kernel event spans)
    footprint = 9 * n * 4
    flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda") if footprint < (160 << 20) else None

    def step():
        if flush is not None:
            flush.fill_(1)
        fa2_b200.forward_backward(q, k, v, do, precision=prec, out=outs)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def all_max(x):
        if dist is None:
            return x
        t_ = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        return float(t_.item())

    for _ in range(args.warmup):
        step()
    barrier()
    lib.fa2_profile_enable(1)
    lib.fa2_profile_kernel_launches()                 # reset the kernel counter: only the timed region is counted
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_host0 = time.perf_counter()
    e0.record(st)
    for _ in range(args.steps):
        step()
    e1.record(st)
    barrier()
    t_host1 = time.perf_counter()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop(t_host0, t_host1)
    kms = (ctypes.c_float * 4)(0, 0, 0, 0)
    kn = (ctypes.c_int * 4)(0, 0, 0, 0)
    lib.fa2_profile_read(kms, kn)
    n_kernels = int(lib.fa2_profile_kernel_launches())        # every kernel of the library launched in the timed region
    lib.fa2_profile_enable(0)
    ms_total = all_max(ms_total)
    ms_step = ms_total / args.steps
    if flush is not None:
        # the L2 flush is not part of the path: the step time is the sum of the library's own kernel spans
        ms_step = all_max(sum(kms[i] for i in range(4)) / args.steps)
    f_fwd, f_bwd = flops(B, H, S, D)
    value = (f_fwd + f_bwd) * world / (ms_step * 1e-3) / 1e12

    # ---- the TIMED outputs against float64 on sampled slabs (first / middle / last work items)
    verify = None
    if not args.no_verify:
        picks, err, finite = verify_outputs(q, k, v, do, outs)
        err = {n_: all_max(e_) for n_, e_ in err.items()}
        ok = all_max(0.0 if finite else 1.0) == 0.0 and all(err[n_] <= TOL[n_] for n_ in TOL)
        verify = {"checked": "outputs of the last timed step vs float64", "slabs_per_rank": picks, "max_abs_err": err,
                  "tol": TOL, "finite": finite, "ok": ok}

    # ---- end to end through the host-pointer C ABI (pinned host buffers, H2D + D2H inside the timed region)
    kernel_ms = ctypes.c_float(0)
    P = lambda t_: ctypes.c_void_p(t_.data_ptr())
    PREC = fa2_b200._lib.PRECISION[prec]
    e2e_val = None
    t_e2e = float("inf")          # a rank that cannot run the leg reports inf; every rank still joins the collectives
    host = None
    try:
        hq, hk, hv, hdo = (torch.randn(B, H, S, D).pin_memory() for _ in range(4))
        ho, hdq, hdk, hdv = (torch.empty(B, H, S, D).pin_memory() for _ in range(4))
        hl = torch.empty(B, H, S).pin_memory()
        host = (hq, hk, hv, hdo, ho, hl, hdq, hdk, hdv)

        def e2e_step(n_gpus=1):
            fa2_b200._lib.check(lib.fa2_host_forward_backward(P(hq), P(hk), P(hv), P(hdo), P(ho), P(hl), P(hdq), P(hdk),
                                                              P(hdv), B, H, S, D, PREC, n_gpus, ctypes.byref(kernel_ms)))
            return kernel_ms.value
        e2e_step()
        ready = 0.0
    except (fa2_b200.FA2Error, RuntimeError, MemoryError) as ex:
        print("e2e leg unavailable on rank %d: %s" % (rank, ex), file=sys.stderr)
        ready = 1.0
    e2e_kernel_ms = None
    if all_max(ready) == 0.0:     # (doubles as the barrier before the timed region)
        try:
            t0 = time.perf_counter()
            kk = 0.0
            for _ in range(args.e2e_steps):
                kk += e2e_step()
            t_e2e = (time.perf_counter() - t0) / args.e2e_steps
            e2e_kernel_ms = kk / args.e2e_steps
        except fa2_b200.FA2Error as ex:
            print("e2e failed on rank %d: %s" % (rank, ex), file=sys.stderr)
        t_e2e = all_max(t_e2e)
        if t_e2e != float("inf"):
            e2e_val = (f_fwd + f_bwd) * world / t_e2e / 1e12
    h2d = 4 * n * 4
    d2h = 4 * n * 4 + B * H * S * 4

    # PCIe ceiling with every rank copying at once (pinned, both directions together): what e2e can reach at best
    pcie = None
    if host is not None:
        try:
            dbuf_in, dbuf_out = torch.empty_like(q), torch.empty_like(q)
            s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
            barrier()
            t0 = time.perf_counter()
            for _ in range(2):
                with torch.cuda.stream(s_in):
                    dbuf_in.copy_(hq, non_blocking=True)
                with torch.cuda.stream(s_out):
                    ho.copy_(dbuf_out, non_blocking=True)
            torch.cuda.synchronize()
            dt = all_max(time.perf_counter() - t0)
            gbs = 2 * n * 4 / dt / 1e9
            floor_ms = max(h2d, d2h) / (gbs * 1e9) * 1e3
            pcie = {"duplex_gbs_per_direction_per_gpu": gbs, "ranks_copying_at_once": world,
                    "e2e_floor_ms_per_step": floor_ms,
                    "e2e_frac_of_floor": floor_ms / (t_e2e * 1e3) if t_e2e not in (0.0, float("inf")) else None}
            del dbuf_in, dbuf_out
        except RuntimeError as ex:
            print("pcie probe failed on rank %d: %s" % (rank, ex), file=sys.stderr)

    # ---- strong scaling, device-resident: ONE workload-sized problem split over the N ranks by fa2_partition; every
    # rank times its share (B*H / N slabs) with the same events as above, T_N = max over ranks, T_1 = the full
    # workload on one GPU (the weak-scaling step above is exactly that).  No collective: slabs are independent.
    strong_dev = None
    if world > 1 and not args.no_strong:
        _, share = fa2_b200.partition(B * H, world, rank)
        if share > 0:
            sq, sk, sv, sg = (t_.view(1, B * H, S, D)[:, :share] for t_ in (q, k, v, do))
            souts = tuple(t_.view(1, B * H, *t_.shape[2:])[:, :share] for t_ in outs)
            sq, sk, sv, sg = (t_.contiguous() for t_ in (sq, sk, sv, sg))
            souts = tuple(t_.contiguous() for t_ in souts)
            for _ in range(args.warmup):
                fa2_b200.forward_backward(sq, sk, sv, sg, precision=prec, out=souts)
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(st)
            for _ in range(args.steps):
                fa2_b200.forward_backward(sq, sk, sv, sg, precision=prec, out=souts)
            s1.record(st)
            barrier()
            tN = all_max(s0.elapsed_time(s1) / args.steps)
        else:
            barrier(); barrier()
            tN = all_max(0.0)
        items_f, items_b = share * ((S + 255) // 256), share * ((S + 127) // 128)
        strong_dev = {"problem": f"ONE B{B} H{H} S{S} D{D} problem, {B * H // world} slabs per GPU (fa2_partition), device-resident",
                      "T1_ms": ms_step, "TN_ms": tN, "efficiency": ms_step / (world * tN) if tN > 0 else None,
                      "tflops_aggregate": (f_fwd + f_bwd) / (tN * 1e-3) / 1e12 if tN > 0 else None,
                      "limiter": f"wave quantisation on 148 persistent CTAs: forward {items_f} items = {items_f / 148:.2f} waves "
                                 f"-> {-(-items_f // 148)} rounds, backward {items_b} items = {items_b / 148:.2f} waves -> "
                                 f"{-(-items_b // 148)} rounds; predicted efficiency "
                                 f"{(2 * items_f / 148 + 5 * items_b / 148) / (2 * -(-items_f // 148) + 5 * -(-items_b // 148)):.3f}"}

    # ---- strong scaling of the product's partitioner: ONE workload-sized problem over N GPUs (rank 0 drives all of
    # them through fa2_host_forward_backward; the other ranks wait on a host-side barrier with their GPUs idle)
    strong = None
    if world > 1 and not args.no_strong and host is not None:
        torch.cuda.synchronize()
        dist.barrier(group=host_group)
        if rank == 0:
            try:
                def timed(n_gpus, reps=3):
                    e2e_step(n_gpus)                                   # warm-up: arenas / streams of every device
                    best_k, best_w = float("inf"), float("inf")
                    for _ in range(reps):
                        t0 = time.perf_counter()
                        km = e2e_step(n_gpus)
                        best_w = min(best_w, (time.perf_counter() - t0) * 1e3)
                        best_k = min(best_k, km)
                    return best_k, best_w
                k1, w1 = timed(1)
                ref = [t_.clone() for t_ in (ho, hl, hdq, hdk, hdv)]
                kN, wN = timed(world)
                diffs = {n_: float((a - b).abs().max()) for n_, a, b in zip(("O", "LSE", "dQ", "dK", "dV"), ref, (ho, hl, hdq, hdk, hdv))}
                same = diffs["O"] == 0 and diffs["LSE"] == 0 and diffs["dK"] == 0 and diffs["dV"] == 0 and diffs["dQ"] <= 1e-5
                n_items_fwd = (B * H // world) * ((S + 255) // 256)
                strong = {"problem": f"ONE B{B} H{H} S{S} D{D} problem split over {world} GPUs by fa2_partition (bh slabs), host API",
                          "kernel_ms": {"T1": k1, "TN": kN, "efficiency": k1 / (world * kN)},
                          "wall_ms": {"T1": w1, "TN": wN, "efficiency": w1 / (world * wN)},
                          "tflops_aggregate_kernel": (f_fwd + f_bwd) / (kN * 1e-3) / 1e12,
                          "max_abs_diff_vs_1gpu": diffs, "equal_to_1gpu": same,
                          "limiter": f"wave quantisation: {n_items_fwd} forward work items per GPU on 148 persistent CTAs "
                                     f"= {n_items_fwd / 148:.2f} waves (per chunk of the host pipeline), plus per-chunk launch tails"}
            except fa2_b200.FA2Error as ex:
                strong = {"error": str(ex)}
        dist.barrier(group=host_group)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak_burst, peak_sus, hbm, peak_src = measured_peaks()
    cast_ms, fwd_ms, pre_ms, bwd_ms = (kms[i] / max(kn[i], 1) for i in range(4))
    traffic = traffic_fwd = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath) and key == "C":    # the ncu DRAM figures were captured on workload C only
        with open(tpath) as fh:
            tj = json.load(fh)
            traffic, traffic_fwd = tj.get("bwd_kernel_dram_bytes_per_launch"), tj.get("fwd_kernel_dram_bytes_per_launch")
    tf_bwd = f_bwd / (bwd_ms * 1e-3) / 1e12 if bwd_ms > 0 else None
    tf_fwd = f_fwd / (fwd_ms * 1e-3) / 1e12 if fwd_ms > 0 else None
    by_fwd, by_bwd = compulsory_bytes(B, H, S, D)
    # MEASURED_PEAKS: the burst figure is for a kernel timed in a short run, the sustained one for a seconds-long loop
    loop_s = ms_total * 1e-3
    peak = peak_sus if loop_s >= 2.0 else peak_burst
    peak_name = "bf16_tflops_sustained" if loop_s >= 2.0 else "bf16_tflops (burst)"
    intensity = f_bwd / by_bwd                               # FLOP per compulsory byte of the dominant kernel
    ridge = peak_burst * 1e12 / (hbm * 1e9)
    if intensity >= ridge:
        roofline = {"bound": "tensor", "kernel": "fa2_bwd_kernel (dominant: %.0f %% of the step)" % (100.0 * bwd_ms / ms_step),
                    "achieved": tf_bwd, "peak": peak, "unit": "TFLOP/s", "frac": tf_bwd / peak if tf_bwd else None,
                    "traffic": traffic,
                    "peak_source": f"MEASURED_PEAKS.json {peak_name} ({peak_src}); timed loop lasted {loop_s:.2f} s",
                    "frac_of_burst_peak": tf_bwd / peak_burst if tf_bwd else None,
                    "frac_of_sustained_peak": tf_bwd / peak_sus if tf_bwd else None,
                    "frac_of_nominal_2250": tf_bwd / 2250 if tf_bwd else None,
                    "algorithmic_flop_per_launch": f_bwd,
                    "fwd_kernel": {"achieved": tf_fwd, "frac": tf_fwd / peak if tf_fwd else None,
                                   "frac_of_burst_peak": tf_fwd / peak_burst if tf_fwd else None,
                                   "frac_of_nominal_2250": tf_fwd / 2250 if tf_fwd else None, "traffic": traffic_fwd}}
    else:
        gbs = by_bwd / (bwd_ms * 1e-3) / 1e9 if bwd_ms > 0 else None
        roofline = {"bound": "hbm", "kernel": "fa2_bwd_kernel (dominant: %.0f %% of the step)" % (100.0 * bwd_ms / ms_step),
                    "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm if gbs else None, "traffic": None,
                    "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src})",
                    "algorithmic_bytes_per_launch": by_bwd,
                    "note": "intensity %.0f FLOP/B is below the ridge (%.0f): bandwidth/latency side of the roofline; "
                            "ideal time at the HBM peak is %.2f us, so the launch is latency-bound" %
                            (intensity, ridge, by_bwd / (hbm * 1e9) * 1e6),
                    "tensor_side": {"achieved_tflops": tf_bwd, "frac_of_burst_peak": tf_bwd / peak_burst if tf_bwd else None},
                    "fwd_kernel": {"achieved_gbs": by_fwd / (fwd_ms * 1e-3) / 1e9 if fwd_ms else None, "achieved_tflops": tf_fwd},
                    "step_gbs": (by_fwd + by_bwd) / (ms_step * 1e-3) / 1e9}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 operands / f32 accumulate (fp32 API tensors)", "data": "synthetic",
        "config": {"workload": workload_name(key, B, H, S, D) + f" per GPU (global batch {B * world})",
                   "precision_flag": prec,
                   "parallelism": f"bh-shard x{world} (fa2_partition), no collective",
                   "l2": ("inputs (4 x %d MiB fp32) exceed the 126 MB L2" % (n * 4 >> 20)) if flush is None else
                         "L2 flushed (192 MiB write) before every step; step time = sum of the kernel spans",
                   "timed_region": "fp32 device tensors in -> fp32 device tensors out: cast + fwd (+ fused bwd pre-pass) + bwd"},
        "clocks": clocks,
        "gpu_launches": n_kernels,
        "kernel_ms": {"cast_qkv": cast_ms, "fwd": fwd_ms, "bwd_prepass": pre_ms, "bwd": bwd_ms},
        "tflops": {"fwd_kernel": tf_fwd, "bwd_kernel": tf_bwd, "step_of_burst_peak": value / world / peak_burst,
                   "step_of_nominal_2250": value / world / 2250},
        "roofline": roofline,
        "verify": verify,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "fa2_host_forward_backward (C ABI, pinned host buffers)", "steps": args.e2e_steps,
                "ms_per_step": t_e2e * 1e3 if t_e2e != float("inf") else None,
                "per_gpu": e2e_val / world if e2e_val else None, "kernel_ms_inside": e2e_kernel_ms, "pcie": pcie},
    }
    if strong is not None or strong_dev is not None:
        line["strong_scaling"] = {"device": strong_dev, "host_api": strong}
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(B, H, S, D)
    print(json.dumps(line))
    sys.stdout.flush()
    if dist is not None:
        dist.destroy_process_group()
    bad = (verify is not None and not verify["ok"]) or (strong is not None and not strong.get("equal_to_1gpu", False))
    strong = {"device": strong_dev, "host_api": strong}
    if bad:
        print("bench.py: OUTPUT CHECK FAILED: " + json.dumps({"verify": verify, "strong": strong}), file=sys.stderr)
        sys.exit(1)


if __name__ == "__main__":
    main()
