"""Port of the reference's test/benchmark harness (test_flash_attention2.py) onto libfa2_b200.so.

Same command line (`--mode/--kernel/--experiment/--seqlen-experiment/--tolerance/--no-stop-on-failure/
--save-results/--output-dir/--no-gpu-reference`, reference :1462-1491), same named configurations
(:1370-1408) and sequence-length sweep (:1436-1457), same data (torch.manual_seed(42) + torch.rand,
:177-195), same PyTorch CPU oracle (matmul -> /sqrt(D) -> softmax -> matmul, autograd with dO = ones,
:197-232), same pass criterion (max-abs error < tolerance and no NaN/Inf, :718-750), same timing
(1 warm-up + 10 timed launches between CUDA events, :283-308) and the same CSV schema (:1108-1123).
Differences: kernels are called through the C ABI instead of CuPy NVRTC (CuPy is not in this image); plots are
skipped when matplotlib is unavailable; `--extra-configs` adds the BASELINE.json shapes (D=128, S up to 16384).
Comparison rows, as in the reference's CSVs (plots/experiment_results.csv: NAIVE, PYTORCH CPU, PYTORCH GPU,
NAIVE-ATTN, FA2 per configuration): "PyTorch CPU" and "PyTorch GPU" rows are always written; with `--experiment`
(or `--kernel fa1|vanilla-attn`) and the reference CLI compiled as it lies (oracle/build_ref.sh ->
oracle/_ref/FlashAttention_ref, or $FA2_BASELINE_CLI) its own CUDA-core kernels are run on the same data through its
file interface and reported as FA2-REFERENCE / FA1 / NAIVE-ATTN rows (D <= 64: its dispatcher refuses more,
include/dispatcher.h:226-227).  Those are reported baselines only: nothing of them is linked into libfa2_b200, and
the fa2-naive kernel (reachable only through the reference's CuPy path, test_flash_attention2.py:315-475) is not run.
"""
from __future__ import annotations

import argparse
import csv
import dataclasses
import os
import re
import shutil
import subprocess
import tempfile
import time
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

CSV_COLUMNS = ["Test", "Kernel", "Type", "Batch", "Heads", "SeqLen", "HeadDim", "Status", "MaxError", "MeanError",
               "MSE", "MaxRelError", "KernelTime_ms", "TorchTime_ms", "Speedup", "TFLOPS", "Bandwidth_GBps",
               "ErrorMessage"]


@dataclass
class TestConfig:
    name: str
    batch_size: int
    num_heads: int
    seq_len: int
    head_dim: int
    test_backward: bool = False
    test_both: bool = False
    kernel_type: str = "fa2"
    seed: int = 42


@dataclass
class TestResult:
    config: TestConfig
    passed: bool
    max_abs_error: float
    mean_abs_error: float
    mse: float
    max_rel_error: float
    kernel_time_ms: float
    torch_time_ms: float
    speedup: float
    tflops: float
    bandwidth_gbps: float
    test_type: str = "forward"
    error_message: str = ""


def create_test_configs(test_mode="forward", kernel_type="fa2") -> List[TestConfig]:
    tb, both = test_mode == "backward", test_mode == "both"
    shapes = [("Small-1", 1, 1, 128), ("Small-2", 2, 4, 256), ("Small-3", 2, 8, 256), ("Medium-1", 2, 8, 512),
              ("Medium-2", 4, 8, 512), ("Large-1", 2, 8, 1024), ("Large-2", 4, 12, 1024),
              ("Edge-NonPowerOf2", 8, 16, 100), ("Edge-SmallSeq", 8, 16, 32), ("Stress-1", 8, 16, 2048)]
    return [TestConfig(n, b, h, s, 64, tb, both, kernel_type) for n, b, h, s in shapes]


def create_sequence_length_experiment_configs(mode) -> List[TestConfig]:
    return [TestConfig(f"SeqLen-S{s}-FA2", 4, 8, s, 64, mode == "backward", mode == "both", "fa2")
            for s in (128, 256, 512, 1024, 2048, 4096)]


def create_extra_configs(mode) -> List[TestConfig]:
    """BASELINE.json shapes the reference itself cannot run (D=128) or never tested."""
    tb, both = mode == "backward", mode == "both"
    return [TestConfig("Baseline-B", 4, 16, 1024, 64, tb, both), TestConfig("Baseline-C-slab", 1, 8, 4096, 128, tb, both),
            TestConfig("Baseline-D-slab", 1, 2, 16384, 128, tb, both), TestConfig("D32", 2, 4, 512, 32, tb, both)]


def baseline_cli() -> Optional[str]:
    """Path of the reference's own CLI built for this GPU (reported baselines only), or None."""
    cand = [os.environ.get("FA2_BASELINE_CLI", ""),
            os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "oracle", "_ref",
                         "FlashAttention_ref")]
    for c in cand:
        if c and os.path.isfile(c) and os.access(c, os.X_OK):
            return c
    return None


BASELINE_METHODS = {"fa2": "FA2-REFERENCE", "fa1": "FA1", "naive": "NAIVE-ATTN"}     # CLI method token -> CSV label


class FlashAttention2Tester:
    def __init__(self, stop_on_failure=True, tolerance=1e-3, test_mode="forward", save_results=False,
                 output_dir="./experiment_results", use_gpu_reference=True, precision="fp32", baseline_methods=()):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("a CUDA device is required: the FA2 path has no CPU fallback")
        self.stop_on_failure, self.tolerance, self.test_mode = stop_on_failure, tolerance, test_mode
        self.save_results, self.output_dir, self.use_gpu_reference = save_results, output_dir, use_gpu_reference
        self.precision = precision
        self.baseline_methods = tuple(baseline_methods) if baseline_cli() else ()
        self.results: List[TestResult] = []
        if save_results:
            os.makedirs(output_dir, exist_ok=True)

    # ---- data and oracles (reference :177-232) --------------------------------------------------
    def generate_test_data(self, config: TestConfig):
        import torch
        torch.manual_seed(config.seed)
        np.random.seed(config.seed)
        shape = (config.batch_size, config.num_heads, config.seq_len, config.head_dim)
        return tuple(torch.rand(*shape, dtype=torch.float32) for _ in range(3))

    @staticmethod
    def compute_reference(Q, K, V):
        import torch
        import torch.nn.functional as F
        scores = torch.matmul(Q, K.transpose(-2, -1)) / (Q.shape[-1] ** 0.5)
        return torch.matmul(F.softmax(scores, dim=-1), V)

    @staticmethod
    def compute_lse(Q, K):
        import torch
        scores = torch.matmul(Q, K.transpose(-2, -1)) / (Q.shape[-1] ** 0.5)
        mx = scores.max(dim=-1, keepdim=True).values
        return (mx + torch.log(torch.exp(scores - mx).sum(dim=-1, keepdim=True))).squeeze(-1)

    def compute_reference_grads(self, Q, K, V):
        import torch
        q, k, v = (t.detach().clone().requires_grad_(True) for t in (Q, K, V))
        out = self.compute_reference(q, k, v)
        t0 = time.time()
        out.backward(torch.ones_like(out))
        return out.detach(), (q.grad, k.grad, v.grad), (time.time() - t0) * 1e3

    # ---- kernel launches through the C ABI (reference :252-313, :477-567) ---------------------------
    @staticmethod
    def _timed(fn, num_runs=10):
        import torch
        fn()                                      # warm-up
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(num_runs):
            fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) / num_runs

    def run_fa2_forward_kernel(self, Q, K, V):
        import torch
        from . import api
        q, k, v = (t.detach().cuda() for t in (Q, K, V))
        out = (torch.empty_like(q), torch.empty(q.shape[:3], device="cuda"))
        ms = self._timed(lambda: api.forward(q, k, v, precision=self.precision, out=out))
        return out[0].cpu().numpy(), out[1].cpu().numpy(), ms

    def run_cuda_fa2_backward_kernel(self, Q, K, V, O, logsumexp):
        import torch
        from . import api
        q, k, v = (t.detach().cuda() for t in (Q, K, V))
        o, l = torch.as_tensor(O).cuda(), torch.as_tensor(logsumexp).cuda()
        g = torch.ones_like(q)
        out = tuple(torch.empty_like(q) for _ in range(3))
        ms = self._timed(lambda: api.backward(q, k, v, o, g, l, precision=self.precision, out=out))
        return {n: t.cpu().numpy() for n, t in zip(("dQ", "dK", "dV"), out)}, ms

    # ---- reported baselines: the reference's own kernels through its CLI (file in, file out) ----------------
    def run_reference_cli(self, method: str, mode: str, Q, K, V):
        """Runs `<reference CLI> <method> <mode> fp32 <dir>` on these tensors (dO = ones: no dO.bin, src/main.cpp:83-93)
        and returns ({name: array}, kernel_ms) with its own TimerGPU figure ("Kernel execution completed", :107)."""
        exe = baseline_cli()
        B, H, S, D = Q.shape
        tmp = tempfile.mkdtemp(prefix="fa2_baseline_")
        try:
            d = os.path.join(tmp, f"B{B}_H{H}_S{S}_D{D}")
            os.makedirs(d)
            for n, t in zip("QKV", (Q, K, V)):
                t.detach().cpu().numpy().astype(np.float32).tofile(os.path.join(d, f"{n}.bin"))
            r = subprocess.run([exe, method, mode, "fp32", d], capture_output=True, text=True, timeout=600)
            if r.returncode != 0:
                raise RuntimeError((r.stderr or r.stdout).strip().splitlines()[-1] if (r.stderr or r.stdout).strip() else "failed")
            m = re.search(r"Kernel execution completed:\s*([0-9.eE+-]+)\s*seconds", r.stdout)
            ms = float(m.group(1)) * 1e3 if m else float("nan")
            out = {}
            for n, shp in (("O", (B, H, S, D)), ("dQ", (B, H, S, D)), ("dK", (B, H, S, D)), ("dV", (B, H, S, D))):
                f = os.path.join(d, n + ".bin")
                if os.path.exists(f):
                    out[n] = np.fromfile(f, np.float32).reshape(shp)
            return out, ms
        finally:
            shutil.rmtree(tmp, ignore_errors=True)

    def baseline_rows(self, config, Q, K, V, expected, exp_g, torch_fwd_ms, torch_bwd_ms):
        rows = []
        if config.head_dim > 64:
            return rows
        cat = lambda d: np.concatenate([np.asarray(d[n]).ravel() for n in ("dQ", "dK", "dV")])
        for method in self.baseline_methods:
            label = BASELINE_METHODS[method]
            cfg = dataclasses.replace(config, kernel_type=label)
            try:
                if self.test_mode == "forward":
                    out, ms = self.run_reference_cli(method, "forward", Q, K, V)
                    rows.append(self._result(cfg, "forward", out["O"], expected.numpy(), ms, torch_fwd_ms, 1.0))
                elif method == "fa2":               # the reference has a backward for fa2 only (dispatcher.h:74-83)
                    out, ms = self.run_reference_cli(method, "forward_backward", Q, K, V)
                    rows.append(self._result(cfg, "both", cat(out), exp_g, ms, torch_fwd_ms + torch_bwd_ms, 3.5))
            except Exception as ex:               # a baseline that cannot run is reported, never fatal
                rows.append(TestResult(cfg, False, float("nan"), float("nan"), float("nan"), float("nan"), 0.0, 0.0, 0.0,
                                       0.0, 0.0, self.test_mode, f"baseline unavailable: {ex}"))
        return rows

    def pytorch_gpu_backward_row(self, config, Q, K, V, exp_g, torch_ms):
        """PyTorch GPU (SDPA, math backend) forward + autograd backward with dO = ones (reference :651-706)."""
        import torch
        import torch.nn.functional as F
        from torch.nn.attention import SDPBackend, sdpa_kernel
        q, k, v = (t.detach().cuda().requires_grad_(True) for t in (Q, K, V))

        def run():
            for t in (q, k, v):
                t.grad = None
            with sdpa_kernel(SDPBackend.MATH):
                o = F.scaled_dot_product_attention(q, k, v)
            o.backward(torch.ones_like(o))
        ms = self._timed(run, num_runs=5)
        got = np.concatenate([t.grad.cpu().numpy().ravel() for t in (q, k, v)])
        cfg = dataclasses.replace(config, kernel_type="PyTorch GPU")
        return self._result(cfg, "backward", got, exp_g, ms, torch_ms, 3.5)

    @staticmethod
    def pytorch_cpu_row(config, test_type, torch_ms, flop_mult):
        """The reference writes its CPU oracle as a row of its own (errors 0, speedup 1; :635-648)."""
        flops = 4.0 * config.batch_size * config.num_heads * config.seq_len ** 2 * config.head_dim * flop_mult
        nbytes = config.batch_size * config.num_heads * config.seq_len * config.head_dim * 4 * 4
        cfg = dataclasses.replace(config, kernel_type="PyTorch CPU")
        t = max(torch_ms, 1e-9) * 1e-3
        return TestResult(cfg, True, 0.0, 0.0, 0.0, 0.0, torch_ms, torch_ms, 1.0, flops / t / 1e12, nbytes / t / 1e9, test_type)

    # ---- metrics (reference :569-606) ---------------------------------------------------------------
    @staticmethod
    def compute_metrics(actual, expected, kernel_time, torch_time, config, flop_mult=1.0):
        err = np.abs(actual - expected)
        rel = np.where(np.abs(expected) > 1e-8, err / np.maximum(np.abs(expected), 1e-30), 0.0)
        flops = 4.0 * config.batch_size * config.num_heads * config.seq_len ** 2 * config.head_dim * flop_mult
        nbytes = config.batch_size * config.num_heads * config.seq_len * config.head_dim * 4 * 4
        return dict(max_abs_error=float(err.max()), mean_abs_error=float(err.mean()),
                    mse=float(np.mean((actual - expected) ** 2)), max_rel_error=float(rel.max()),
                    tflops=flops / (kernel_time * 1e-3) / 1e12, bandwidth_gbps=nbytes / (kernel_time * 1e-3) / 1e9,
                    speedup=torch_time / kernel_time if kernel_time > 0 else 0.0)

    def _result(self, config, test_type, actual, expected, ms, torch_ms, flop_mult):
        m = self.compute_metrics(actual, expected, ms, torch_ms, config, flop_mult)
        finite = bool(np.isfinite(actual).all())
        passed = finite and m["max_abs_error"] < self.tolerance
        msg = "" if passed else ("NaN/Inf in output" if not finite else
                                 f"Max error {m['max_abs_error']:.2e} exceeds tolerance {self.tolerance:.2e}")
        return TestResult(config, passed, m["max_abs_error"], m["mean_abs_error"], m["mse"], m["max_rel_error"], ms,
                          torch_ms, m["speedup"], m["tflops"], m["bandwidth_gbps"], test_type, msg)

    def run_test(self, config: TestConfig) -> List[TestResult]:
        import torch
        print(f"\nRunning test: {config.name}  (B={config.batch_size}, H={config.num_heads}, S={config.seq_len}, "
              f"D={config.head_dim}, mode={self.test_mode})")
        Q, K, V = self.generate_test_data(config)
        t0 = time.time()
        expected = self.compute_reference(Q, K, V)
        torch_fwd_ms = (time.time() - t0) * 1e3
        out: List[TestResult] = []
        cat = lambda d: np.concatenate([np.asarray(d[n]).ravel() for n in ("dQ", "dK", "dV")])
        if self.test_mode == "forward":
            out.append(self.pytorch_cpu_row(config, "forward", torch_fwd_ms, 1.0))
        if self.use_gpu_reference and self.test_mode == "forward":
            import torch.nn.functional as F
            qg, kg, vg = (t.cuda() for t in (Q, K, V))
            F.scaled_dot_product_attention(qg, kg, vg)
            torch.cuda.synchronize()
            ms = self._timed(lambda: F.scaled_dot_product_attention(qg, kg, vg))
            cfg = dataclasses.replace(config, kernel_type="PyTorch GPU")
            out.append(self._result(cfg, "forward", F.scaled_dot_product_attention(qg, kg, vg).cpu().numpy(),
                                    expected.numpy(), ms, torch_fwd_ms, 1.0))
        if self.test_mode == "forward":
            out.extend(self.baseline_rows(config, Q, K, V, expected, None, torch_fwd_ms, 0.0))
            O, _, ms = self.run_fa2_forward_kernel(Q, K, V)
            out.append(self._result(config, "forward", O, expected.numpy(), ms, torch_fwd_ms, 1.0))
            return out
        _, grads, torch_bwd_ms = self.compute_reference_grads(Q, K, V)
        exp_g = cat(dict(zip(("dQ", "dK", "dV"), (g.numpy() for g in grads))))
        out.append(self.pytorch_cpu_row(config, self.test_mode, torch_fwd_ms + torch_bwd_ms, 3.5))
        if self.use_gpu_reference:
            out.append(self.pytorch_gpu_backward_row(config, Q, K, V, exp_g, torch_fwd_ms + torch_bwd_ms))
        out.extend(self.baseline_rows(config, Q, K, V, expected, exp_g, torch_fwd_ms, torch_bwd_ms))
        if self.test_mode == "backward":           # PyTorch forward feeds the CUDA backward (:917-928)
            g, ms = self.run_cuda_fa2_backward_kernel(Q, K, V, expected.numpy(), self.compute_lse(Q, K).numpy())
            out.append(self._result(config, "backward", cat(g), exp_g, ms, torch_fwd_ms + torch_bwd_ms, 2.5))
            return out
        O, L, ms_f = self.run_fa2_forward_kernel(Q, K, V)          # both: our own O / LSE feed the backward (:722-728)
        out.append(self._result(config, "forward", O, expected.numpy(), ms_f, torch_fwd_ms, 1.0))
        g, ms_b = self.run_cuda_fa2_backward_kernel(Q, K, V, O, L)
        out.append(self._result(config, "backward", cat(g), exp_g, ms_b, torch_bwd_ms, 2.5))
        return out

    def run_all_tests(self, configs: List[TestConfig]):
        for cfg in configs:
            try:
                results = self.run_test(cfg)
            except Exception as ex:                          # same behaviour as the reference: record and go on / stop
                results = [TestResult(cfg, False, float("nan"), float("nan"), float("nan"), float("nan"), 0.0, 0.0, 0.0,
                                      0.0, 0.0, self.test_mode, str(ex))]
            self.results.extend(results)
            for r in results:
                print(f"  {r.config.kernel_type:12s} {r.test_type:8s} {'PASS' if r.passed else 'FAIL'}  max_err={r.max_abs_error:.2e} "
                      f"time={r.kernel_time_ms:.4f} ms  {r.tflops:.1f} TFLOPS  speedup vs CPU {r.speedup:.1f}x")
            if self.stop_on_failure and not all(r.passed for r in results if r.config.kernel_type == "fa2"):
                print("Stopping on first failure")
                break
        self.results.sort(key=lambda x: x.config.batch_size * x.config.num_heads * x.config.seq_len * x.config.head_dim)
        self.print_summary()
        if self.save_results:
            self.save_results_to_files()

    def print_summary(self):
        from tabulate import tabulate
        rows = [[r.config.name, r.config.kernel_type.upper(), r.test_type.upper(),
                 f"B{r.config.batch_size}_H{r.config.num_heads}_S{r.config.seq_len}_D{r.config.head_dim}",
                 "PASS" if r.passed else "FAIL", f"{r.max_abs_error:.2e}", f"{r.mean_abs_error:.2e}",
                 f"{r.kernel_time_ms:.6f}", f"{r.tflops:.2f}", f"{r.bandwidth_gbps:.2f}"] for r in self.results]
        print("\n" + "=" * 80 + "\nTEST SUMMARY\n" + "=" * 80)
        print(tabulate(rows, headers=["Test", "Kernel", "Type", "Config", "Status", "Max Err", "Mean Err", "Time (ms)",
                                      "TFLOPS", "BW (GB/s)"], tablefmt="grid"))
        total, passed = len(self.results), sum(r.passed for r in self.results)
        print(f"\nTotal Tests: {total}\nPassed: {passed}\nFailed: {total - passed}")

    def save_results_to_files(self):
        path = os.path.join(self.output_dir, "experiment_results.csv")
        with open(path, "w", newline="") as f:
            w = csv.DictWriter(f, fieldnames=CSV_COLUMNS)
            w.writeheader()
            for r in self.results:
                c = r.config
                w.writerow({"Test": c.name, "Kernel": c.kernel_type.upper(), "Type": r.test_type.upper()[:3], "Batch": c.batch_size,
                            "Heads": c.num_heads, "SeqLen": c.seq_len, "HeadDim": c.head_dim,
                            "Status": "PASS" if r.passed else "FAIL", "MaxError": r.max_abs_error, "MeanError": r.mean_abs_error,
                            "MSE": r.mse, "MaxRelError": r.max_rel_error, "KernelTime_ms": r.kernel_time_ms,
                            "TorchTime_ms": r.torch_time_ms, "Speedup": r.speedup, "TFLOPS": r.tflops,
                            "Bandwidth_GBps": r.bandwidth_gbps, "ErrorMessage": r.error_message})
        print(f"Saved results to: {path}  (plots skipped: matplotlib/seaborn are not part of this image)")


def main(argv: Optional[List[str]] = None) -> int:
    ap = argparse.ArgumentParser(description="Test the B200 FlashAttention-2 kernels (forward and backward passes)")
    ap.add_argument("--mode", required=True, choices=["forward", "backward", "both"])
    ap.add_argument("--kernel", default="fa2", choices=["fa2", "fa1", "vanilla-attn", "fa2-naive"])
    ap.add_argument("--experiment", action="store_true")
    ap.add_argument("--seqlen-experiment", action="store_true")
    ap.add_argument("--tolerance", type=float, default=1e-3)
    ap.add_argument("--no-stop-on-failure", action="store_true")
    ap.add_argument("--save-results", action="store_true")
    ap.add_argument("--output-dir", default="./experiment_results")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp16", "bf16"])
    ap.add_argument("--extra-configs", action="store_true", help="add the BASELINE.json shapes (D=128, long S)")
    ap.add_argument("--no-baselines", action="store_true", help="do not run the reference's own kernels as comparison rows")
    a = ap.parse_args(argv)
    baselines = ()
    if a.kernel == "fa2-naive":
        ap.error("kernel 'fa2-naive' is reachable only through the reference's CuPy path and is not provided here")
    if a.kernel != "fa2":
        # the reference's comparison kernels run as reported baselines through its own CLI; fa2 is always the kernel under test
        if a.mode != "forward":
            ap.error("Backward pass testing is only supported for the FA2 kernel")        # reference :1494-1495
        if baseline_cli() is None:
            ap.error(f"kernel '{a.kernel}' is a comparison baseline of the reference: build its CLI with oracle/build_ref.sh "
                     "(or point $FA2_BASELINE_CLI at it); only 'fa2' is implemented here")
        baselines = ({"fa1": "fa1", "vanilla-attn": "naive"}[a.kernel],)
    elif a.experiment and not a.no_baselines:
        baselines = ("fa2", "fa1", "naive") if a.mode == "forward" else ("fa2",)
    configs = (create_sequence_length_experiment_configs(a.mode) if a.seqlen_experiment
               else create_test_configs(a.mode, "fa2"))
    if a.extra_configs:
        configs += create_extra_configs(a.mode)
    t = FlashAttention2Tester(not a.no_stop_on_failure, a.tolerance, a.mode, a.save_results, a.output_dir,
                              not a.no_gpu_reference, a.precision, baselines)
    t.run_all_tests(configs)
    own = [r for r in t.results if r.config.kernel_type == "fa2"]       # baseline rows are reported, not judged
    return 0 if own and all(r.passed for r in own) else 1


if __name__ == "__main__":
    raise SystemExit(main())
