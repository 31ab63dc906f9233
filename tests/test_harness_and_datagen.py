"""Harness port / data generator: CPU-side checks (schema, configs, generator bytes) and a GPU run."""
import csv
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cuda-flash-attention_b200")


def test_configs_match_reference_lists():
    from fa2_b200 import harness
    cfgs = harness.create_test_configs("both")
    assert [(c.name, c.batch_size, c.num_heads, c.seq_len, c.head_dim) for c in cfgs] == [
        ("Small-1", 1, 1, 128, 64), ("Small-2", 2, 4, 256, 64), ("Small-3", 2, 8, 256, 64), ("Medium-1", 2, 8, 512, 64),
        ("Medium-2", 4, 8, 512, 64), ("Large-1", 2, 8, 1024, 64), ("Large-2", 4, 12, 1024, 64),
        ("Edge-NonPowerOf2", 8, 16, 100, 64), ("Edge-SmallSeq", 8, 16, 32, 64), ("Stress-1", 8, 16, 2048, 64)]
    assert all(c.test_both for c in cfgs)
    sweep = harness.create_sequence_length_experiment_configs("forward")
    assert [c.seq_len for c in sweep] == [128, 256, 512, 1024, 2048, 4096] and all((c.batch_size, c.num_heads) == (4, 8) for c in sweep)
    assert harness.CSV_COLUMNS == ["Test", "Kernel", "Type", "Batch", "Heads", "SeqLen", "HeadDim", "Status", "MaxError",
                                   "MeanError", "MSE", "MaxRelError", "KernelTime_ms", "TorchTime_ms", "Speedup", "TFLOPS",
                                   "Bandwidth_GBps", "ErrorMessage"]


def test_datagen_bytes_equal_reference_golden(tmp_path, golden_dir):
    sys.path.insert(0, PKG)
    import generate_test_data as gen
    d = gen.generate_test_data(1, 2, 64, 64, output_dir=str(tmp_path), seed=42)
    z = np.load(os.path.join(golden_dir, "cli_B1_H2_S64_D64.npz"))          # bytes written by the reference generator
    for n in "QKV":
        assert np.array_equal(np.fromfile(os.path.join(d, f"{n}.bin"), np.float32).reshape(1, 2, 64, 64), z[n])
    assert os.path.basename(d) == "B1_H2_S64_D64"


def test_harness_rejects_backward_for_baseline_kernels():
    r = subprocess.run([sys.executable, os.path.join(PKG, "test_flash_attention2.py"), "--mode", "backward", "--kernel", "fa1"],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "fa1" in r.stderr


@pytest.mark.gpu
def test_harness_both_mode_end_to_end(tmp_path):
    from fa2_b200 import harness
    rc = harness.main(["--mode", "both", "--no-stop-on-failure", "--save-results", "--output-dir", str(tmp_path),
                       "--no-gpu-reference"])
    rows = list(csv.DictReader(open(tmp_path / "experiment_results.csv")))
    assert rc == 0, [r for r in rows if r["Status"] != "PASS"]
    ours = [r for r in rows if r["Kernel"] == "FA2"]
    assert len(ours) == 20 and set(rows[0]) == set(harness.CSV_COLUMNS)       # 10 configs x (forward, backward)
    assert all(float(r["MaxError"]) < 1e-3 for r in ours)                    # the reference's own tolerance
    assert sum(r["Kernel"] == "PYTORCH CPU" for r in rows) == 10             # the CPU oracle as a row of its own (:635-648)


@pytest.mark.gpu
def test_harness_forward_with_gpu_reference_and_extra_configs(tmp_path):
    from fa2_b200 import harness
    t = harness.FlashAttention2Tester(stop_on_failure=False, tolerance=1e-3, test_mode="forward")
    t.run_all_tests(harness.create_extra_configs("forward")[:2] + harness.create_test_configs("forward")[:2])
    ours = [r for r in t.results if r.config.kernel_type == "fa2"]
    assert len(ours) == 4 and all(r.passed for r in ours)
    assert any(r.config.kernel_type == "PyTorch GPU" for r in t.results)


@pytest.mark.gpu
def test_harness_experiment_reports_the_reference_kernels_as_baseline_rows(tmp_path):
    """--experiment lays our FA2 beside the reference's own CUDA-core kernels (fa2, fa1, vanilla) run through its CLI
    on the same data, like plots/experiment_results.csv; those rows are reported, only the FA2 rows are judged."""
    from fa2_b200 import harness
    if harness.baseline_cli() is None:
        pytest.skip("oracle/_ref/FlashAttention_ref not built")
    t = harness.FlashAttention2Tester(stop_on_failure=False, tolerance=1e-3, test_mode="forward", save_results=True,
                                      output_dir=str(tmp_path), baseline_methods=("fa2", "fa1", "naive"))
    t.run_all_tests(harness.create_test_configs("forward")[1:4])
    kinds = {}
    for r in t.results:
        kinds.setdefault(r.config.kernel_type, []).append(r)
    for k in ("fa2", "FA2-REFERENCE", "FA1", "NAIVE-ATTN", "PyTorch CPU", "PyTorch GPU"):
        assert len(kinds.get(k, [])) == 3, (k, list(kinds))
    assert all(r.passed for r in kinds["fa2"])
    assert all(r.passed and r.max_abs_error < 1e-5 for r in kinds["FA2-REFERENCE"])   # fp32 CUDA-core kernels
    assert all(np.isfinite(r.kernel_time_ms) and r.kernel_time_ms > 0 for r in kinds["FA1"] + kinds["NAIVE-ATTN"])
    t2 = harness.FlashAttention2Tester(stop_on_failure=False, tolerance=1e-3, test_mode="both", baseline_methods=("fa2",))
    t2.run_all_tests(harness.create_test_configs("both")[1:2])
    ref = [r for r in t2.results if r.config.kernel_type == "FA2-REFERENCE"]
    assert len(ref) == 1 and ref[0].passed and ref[0].test_type == "both"
    rows = list(csv.DictReader(open(tmp_path / "experiment_results.csv")))
    assert {"FA2", "FA2-REFERENCE", "FA1", "NAIVE-ATTN", "PYTORCH CPU", "PYTORCH GPU"} <= {r["Kernel"] for r in rows}
