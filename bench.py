#!/usr/bin/env python
"""bench.py -- headline benchmark of the FA2 forward+backward path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload C|A|B|D]
    (N > 1: python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...)

A step = one forward + backward pass (4 kernels: fp32->fp16 cast, forward, backward pre-pass,
backward) over one synthetic batch of the workload, on fp32 [B,H,S,D] device tensors.
Default workload = BASELINE.json configs[2], B8 H32 S4096 D128 (the shape the metric is quoted
on); with N ranks every rank runs that shape on its own slab range of a global batch 8*N
(weak scaling, no data-path collective -- (batch, head) slabs are independent).
Rank 0 prints ONE JSON line.  FLOP convention: fwd 4*BHS^2*D, bwd 10*BHS^2*D.
"""
import argparse
import json
import os
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
METRIC = "FA2 fwd+bwd TFLOPS/GPU-aggregate, B8 H32 S4096 D128 per GPU"
UNIT = "TFLOP/s"


def flops(B, H, S, D):
    f = 4.0 * B * H * S * S * D
    return f, 2.5 * f


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
    """--impl reference: the UNMODIFIED reference CLI (oracle/_ref/FlashAttention_ref, its own sources
    compiled for sm_100) run through its own argv surface; the number is its own TimerGPU total
    ('Kernel execution completed', src/main.cpp:107).  The reference rejects D=128
    (include/dispatcher.h:226-227), so each step is a bounded equal-S sample at D=64."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    exe = os.path.join(ROOT, "oracle", "_ref", "FlashAttention_ref")
    if not os.path.exists(exe):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/FlashAttention_ref not built (run __graft_entry__.build() where /root/reference exists)"}))
        return
    import numpy as np
    B, H, S, D = 2, 16, 4096, 64            # 32 slabs x S4096: fills the GPU, 1/8 of config C's fwd FLOPs... at D=64
    f, b = flops(B, H, S, D)
    times = []
    with tempfile.TemporaryDirectory() as tmp:
        d = os.path.join(tmp, f"B{B}_H{H}_S{S}_D{D}")
        os.makedirs(d)
        rng = np.random.default_rng(42)
        for n in "QKV":
            rng.standard_normal((B, H, S, D), dtype=np.float32).tofile(os.path.join(d, f"{n}.bin"))
        t_wall0 = time.perf_counter()
        for i in range(args.warmup + args.steps):
            out = subprocess.run([exe, "fa2", "forward_backward", "fp32", d], capture_output=True, text=True, timeout=600)
            if out.returncode != 0:
                print(json.dumps({"impl": "reference", "unavailable": "reference CLI failed: " + (out.stderr or out.stdout)[-200:].replace("\n", " ")}))
                return
            secs = [float(l.split(":")[1].split()[0]) for l in out.stdout.splitlines() if l.startswith("Kernel execution completed")]
            if i >= args.warmup:
                times.append(secs[0])
        wall = time.perf_counter() - t_wall0
    t = statistics.mean(times)
    val = (f + b) / t / 1e12
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"B{B}_H{H}_S{S}_D{D} fa2 forward_backward fp32 through the reference CLI: bounded sample "
                               "of the S4096 workload at D=64 because the reference rejects D=128 "
                               "(include/dispatcher.h:226-227); device = B200 CUDA cores, kernels as shipped",
                   "timer": "reference TimerGPU total (fwd kernel + D kernel + bwd kernel), src/main.cpp:107"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "reference",
                         "sample": f"{args.steps} CLI runs of B{B}_H{H}_S{S}_D{D}; the reference has no CPU implementation, its fa2 kernels run on the GPU's CUDA cores; wall {wall:.1f} s incl. file I/O"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C", choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    if args.gpus != world and rank == 0 and world > 1:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    B, H, S, D = WORKLOADS[args.workload]
    BH_global = B * H * world
    bh0, cnt = fa2_b200.partition(BH_global, world, rank)        # this rank's slab range of the global batch
    assert cnt == B * H
    g = torch.Generator(device="cuda").manual_seed(1234 + bh0)
    q, k, v, do = (torch.randn(B, H, S, D, device="cuda", generator=g) for _ in range(4))
    outs = (torch.empty_like(q), torch.empty(B, H, S, device="cuda"), torch.empty_like(q), torch.empty_like(q), torch.empty_like(q))
    lib = fa2_b200.load()
    st = torch.cuda.current_stream()

    def step():
        fa2_b200.forward_backward(q, k, v, do, out=outs)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    lib.fa2_profile_enable(1)
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
    lib.fa2_profile_enable(0)
    if dist is not None:
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    f_fwd, f_bwd = flops(B, H, S, D)
    value = (f_fwd + f_bwd) * world / (ms_step * 1e-3) / 1e12

    # ---- end to end through the host-pointer C ABI (pinned host buffers, H2D + D2H inside the timed region)
    n = B * H * S * D
    kernel_ms = ctypes.c_float(0)
    P = lambda t_: ctypes.c_void_p(t_.data_ptr())
    e2e_val = None
    t_e2e = float("inf")          # a rank that cannot run the leg reports inf; every rank still joins the collectives

    def all_max(x):
        if dist is None:
            return x
        t_ = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        return float(t_.item())

    try:
        hq, hk, hv, hdo = (torch.randn(B, H, S, D).pin_memory() for _ in range(4))
        ho, hdq, hdk, hdv = (torch.empty(B, H, S, D).pin_memory() for _ in range(4))
        hl = torch.empty(B, H, S).pin_memory()

        def e2e_step():
            fa2_b200._lib.check(lib.fa2_host_forward_backward(P(hq), P(hk), P(hv), P(hdo), P(ho), P(hl), P(hdq), P(hdk),
                                                              P(hdv), B, H, S, D, 1, 1, ctypes.byref(kernel_ms)))
        e2e_step()
        ready = 0.0
    except (fa2_b200.FA2Error, RuntimeError, MemoryError) as ex:
        print("e2e leg unavailable on rank %d: %s" % (rank, ex), file=sys.stderr)
        ready = 1.0
    if all_max(ready) == 0.0:     # (doubles as the barrier before the timed region)
        try:
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                e2e_step()
            t_e2e = (time.perf_counter() - t0) / args.e2e_steps
        except fa2_b200.FA2Error as ex:
            print("e2e failed on rank %d: %s" % (rank, ex), file=sys.stderr)
        t_e2e = all_max(t_e2e)
        if t_e2e != float("inf"):
            e2e_val = (f_fwd + f_bwd) * world / t_e2e / 1e12
    h2d = 4 * n * 4
    d2h = 4 * n * 4 + B * H * S * 4

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak_burst, peak_sus, hbm, peak_src = measured_peaks()
    bwd_ms = kms[3] / max(kn[3], 1)
    fwd_ms = kms[1] / max(kn[1], 1)
    traffic = traffic_fwd = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            tj = json.load(fh)
            traffic, traffic_fwd = tj.get("bwd_kernel_dram_bytes_per_launch"), tj.get("fwd_kernel_dram_bytes_per_launch")
    if args.workload != "C":
        traffic = traffic_fwd = None            # the ncu DRAM figures were captured on workload C only
    achieved = f_bwd / (bwd_ms * 1e-3) / 1e12 if bwd_ms > 0 else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 operands / f32 accumulate (fp32 API tensors)", "data": "synthetic",
        "config": {"workload": f"configs[2] B{B} H{H} S{S} D{D} fwd+bwd per GPU (global batch {B * world})",
                   "parallelism": f"bh-shard x{world} (fa2_partition), no collective",
                   "l2": "inputs (4 x %d MiB fp32) exceed the 126 MB L2" % (n * 4 >> 20),
                   "timed_region": "fp32 device tensors in -> fp32 device tensors out: cast + fwd + bwd pre-pass + bwd"},
        "clocks": clocks,
        "gpu_launches": int(sum(kn)),
        "kernel_ms": {"cast_qkv": kms[0] / max(kn[0], 1), "fwd": fwd_ms, "bwd_prepass": kms[2] / max(kn[2], 1), "bwd": bwd_ms},
        "tflops": {"fwd_kernel": f_fwd / (fwd_ms * 1e-3) / 1e12 if fwd_ms else None,
                   "bwd_kernel": achieved,
                   "frac_of_nominal_2250": {"fwd": f_fwd / (fwd_ms * 1e-3) / 1e12 / 2250 if fwd_ms else None,
                                            "bwd": achieved / 2250 if achieved else None}},
        # the kernels are timed inside a long back-to-back loop (clocks under sw_power_cap, see "clocks"), so the
        # denominator is the SUSTAINED measured cuBLAS bf16 peak; the burst figure is given beside it
        "roofline": {"bound": "tensor", "kernel": "fa2_bwd_kernel<128,false> (dominant: %.0f %% of the step)" % (100.0 * bwd_ms / ms_step),
                     "achieved": achieved, "peak": peak_sus, "unit": "TFLOP/s", "frac": achieved / peak_sus if achieved else None,
                     "traffic": traffic,
                     "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peak_src}); kernel timed inside the {args.steps}-step loop",
                     "frac_of_burst_peak": achieved / peak_burst if achieved else None, "burst_peak": peak_burst,
                     "algorithmic_flop_per_launch": f_bwd,
                     "secondary_limit": "fp32 reduce-add of dQ into L2: 64 KB per 128x128 tile pair at ~24 B/clk/SM (tools/reduce_probe.cu)",
                     "fwd_kernel": {"achieved": f_fwd / (fwd_ms * 1e-3) / 1e12 if fwd_ms else None,
                                    "frac": (f_fwd / (fwd_ms * 1e-3) / 1e12) / peak_sus if fwd_ms else None,
                                    "frac_of_burst_peak": (f_fwd / (fwd_ms * 1e-3) / 1e12) / peak_burst if fwd_ms else None,
                                    "traffic": traffic_fwd}},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "fa2_host_forward_backward (C ABI, pinned host buffers)", "steps": args.e2e_steps},
    }
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(B, H, S, D)
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
