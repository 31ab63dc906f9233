"""CPU-side checks: the C-ABI library loads and exports every symbol include/fa2_b200.h declares,
the partitioner arithmetic, error behaviour without a GPU, and the CLI's argument grammar."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import fa2_b200
from fa2_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fa2_b200.h")
CLI = os.path.join(ROOT, "cuda-flash-attention_b200", "FlashAttention")


def header_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fa2_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/fa2_b200.h but not exported"
    assert set(syms) == set(_lib.SIGNATURES), "ctypes SIGNATURES out of sync with the header"


def test_version_and_workspace_bytes():
    lib = fa2_b200.load()
    assert lib.fa2_version() >= 100
    rows, DP = 8 * 32 * 4096, 128
    assert lib.fa2_workspace_bytes(8, 32, 4096, 128, 0) == 1024 + 3 * rows * DP * 2          # 1 KB range block first
    assert lib.fa2_workspace_bytes(8, 32, 4096, 128, 2) == 1024 + 4 * rows * DP * 2 + 2 * rows * 4
    assert lib.fa2_workspace_bytes(1, 1, 16, 48, 0) == 0          # unsupported head dim


@pytest.mark.parametrize("BH,n", [(256, 8), (16, 8), (2, 8), (7, 3), (1, 1), (128, 4), (5, 8)])
def test_partition_covers_exactly_once(BH, n):
    seen = []
    for part in range(n):
        bh0, cnt = fa2_b200.partition(BH, n, part)
        assert 0 <= cnt <= -(-BH // n)
        seen.extend(range(bh0, bh0 + cnt))
    assert seen == list(range(BH))                                 # contiguous, ordered, disjoint, complete
    sizes = [fa2_b200.partition(BH, n, p)[1] for p in range(n)]
    assert max(sizes) - min(sizes) <= 1


def test_partition_rejects_bad_requests():
    with pytest.raises(fa2_b200.FA2Error):
        fa2_b200.partition(8, 0, 0)
    with pytest.raises(fa2_b200.FA2Error):
        fa2_b200.partition(8, 2, 2)


def test_no_cpu_fallback_host_entry_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    x = np.zeros((1, 1, 16, 64), np.float32)
    with pytest.raises(fa2_b200.FA2Error) as ei:
        fa2_b200.run_flash_attention(x, x, x)
    assert ei.value.code == 2 and "no CPU fallback" in str(ei.value)
    assert fa2_b200.load().fa2_device_count() == 0


def test_dispatcher_surface_errors_match_reference():
    x = np.zeros((1, 1, 16, 48), np.float32)
    with pytest.raises(fa2_b200.FA2Error, match="Unsupported head dimension 48"):     # dispatcher.h:137
        fa2_b200.run_flash_attention(x, x, x)
    y = np.zeros((1, 1, 16, 64), np.float32)
    with pytest.raises(ValueError):
        fa2_b200.run_flash_attention(y, y, y, method="fa3")
    with pytest.raises(fa2_b200.FA2Error):
        fa2_b200.run_flash_attention(y, y, y, method="fa1")
    with pytest.raises(ValueError):
        fa2_b200.run_flash_attention(y, y, y, mode="backward")                       # needs O and logsumexp
    with pytest.raises(ValueError):
        fa2_b200.run_flash_attention(y.astype(np.float64), y, y)


def test_pinned_alloc_falls_back_to_malloc_without_gpu():
    lib = fa2_b200.load()
    p = lib.fa2_host_alloc(1 << 20)
    assert p
    ctypes.memset(p, 0, 1 << 20)
    lib.fa2_host_free(p)


def run_cli(*argv):
    return subprocess.run([CLI, *argv], capture_output=True, text=True, timeout=60)


def test_cli_usage_on_too_few_args():
    r = run_cli("fa2", "forward", "fp32")
    assert r.returncode == 1 and r.stderr.startswith("USAGE:")                      # error_utils.h:15-19


@pytest.mark.parametrize("argv", [("fa3", "forward", "fp32", "x/B1_H1_S8_D64"),
                                  ("fa2", "sideways", "fp32", "x/B1_H1_S8_D64"),
                                  ("fa2", "forward", "fp64", "x/B1_H1_S8_D64")])
def test_cli_usage_on_bad_tokens(argv):
    r = run_cli(*argv)
    assert r.returncode == 1 and "USAGE:" in r.stderr


def test_cli_bad_folder_name_dies_like_reference():
    r = run_cli("fa2", "forward", "fp32", "/tmp/not_a_config")
    assert r.returncode == 1 and "sscanf" in r.stderr                                # utils.cpp:48


@pytest.mark.parametrize("mode", ["forward", "backward", "forward_backward", "both", "forward-backward"])
def test_cli_missing_files_and_mode_aliases(tmp_path, mode):
    d = tmp_path / "B1_H2_S16_D64"
    d.mkdir()
    r = run_cli("fa2", mode, "fp32", str(d) + "/")                                   # trailing slash allowed
    assert r.returncode == 1 and "Data files not found" in r.stderr                  # main.cpp:97
    assert "Num heads:     2" in r.stdout and "Sequence len:  16" in r.stdout


def test_cli_unsupported_head_dim_and_methods(tmp_path):
    d = tmp_path / "B1_H1_S16_D48"
    d.mkdir()
    r = run_cli("fa2", "forward", "fp32", str(d))
    assert r.returncode == 1 and "Unsupported head dimension 48" in r.stderr
    d2 = tmp_path / "B1_H1_S16_D64"
    d2.mkdir()
    r = run_cli("fa1", "backward", "fp32", str(d2))
    assert r.returncode == 1 and "Flash Attention 1 backward pass not implemented" in r.stderr   # dispatcher.h:74-78
    r = run_cli("naive", "forward_backward", "fp32", str(d2))
    assert r.returncode == 1 and "Vanilla Attention backward pass not implemented" in r.stderr


def test_cli_backward_needs_forward_outputs(tmp_path):
    d = tmp_path / "B1_H1_S16_D64"
    d.mkdir()
    for n in "QKV":
        np.zeros(16 * 64, np.float32).tofile(d / f"{n}.bin")
    r = run_cli("fa2", "backward", "fp32", str(d))
    assert r.returncode == 1 and "Data files not found" in r.stderr                  # O.bin / logsumexp.bin missing


def test_cli_short_file_is_an_error(tmp_path):
    import torch
    d = tmp_path / "B1_H1_S16_D64"
    d.mkdir()
    for n in "QKV":
        np.zeros(10, np.float32).tofile(d / f"{n}.bin")
    r = run_cli("fa2", "forward", "fp32", str(d))
    assert r.returncode == 1 and "fread" in r.stderr                                 # utils.cpp:17


def test_host_surface_rejects_mismatched_shapes_before_touching_the_library():
    q = np.zeros((1, 2, 16, 64), np.float32)
    with pytest.raises(ValueError, match="V has shape"):
        fa2_b200.run_flash_attention(q, q, np.zeros((1, 2, 8, 64), np.float32))
    with pytest.raises(ValueError, match="dO has shape"):
        fa2_b200.run_flash_attention(q, q, q, dO=np.zeros((1, 2, 16, 32), np.float32), mode="forward_backward")
    with pytest.raises(ValueError, match="logsumexp has shape"):
        fa2_b200.run_flash_attention(q, q, q, O=q, logsumexp=np.zeros((1, 2, 15), np.float32), mode="backward")


def test_device_surface_validates_every_tensor_on_cpu_tensors():
    import torch
    from fa2_b200 import api
    x = torch.zeros(1, 1, 16, 64)
    with pytest.raises(ValueError, match="CUDA tensors"):
        api.forward(x, x, x)
    with pytest.raises(TypeError):
        api._check_dev((1, 1, 16, 64), Q=np.zeros((1, 1, 16, 64), np.float32))


@pytest.mark.parametrize("count,S,D,mode", [(256, 4096, 128, 2), (32, 4096, 128, 2), (16, 16384, 128, 2), (80, 1024, 64, 2),
                                            (48, 4096, 64, 0), (1, 100, 64, 1), (16, 512, 64, 2), (7, 33, 32, 2)])
def test_plan_chunks_covers_the_share(count, S, D, mode):
    """Host pipeline chunking (fa2_plan_chunks): chunk sizes add up to the device's share, in order; large jobs ramp
    up geometrically from a small first chunk to ~20 MiB copies and back down (PCIe-bound path: short head and
    tail, full duplex rate in between), tiny jobs stay in one chunk."""
    lib = fa2_b200.load()
    buf = (ctypes.c_int * 1024)()
    n = lib.fa2_plan_chunks(count, S, D, mode, buf, 1024)
    sizes = list(buf[:n])
    assert n >= 1 and sum(sizes) == count and all(s > 0 for s in sizes)
    peak = sizes.index(max(sizes))
    assert all(a <= b for a, b in zip(sizes[:peak], sizes[1:peak + 1]))          # up ...
    last_big = n - 1 - sizes[::-1].index(max(sizes))
    assert all(a >= b for a, b in zip(sizes[last_big:], sizes[last_big + 1:]))   # ... and down again
    if (count, S) == (256, 4096):
        assert sizes[0] <= 3 and sizes[-1] <= 3 and 8 <= max(sizes) <= 12        # config C: short head / tail, ~20 MiB copies
    if (count, S) == (16, 512):
        assert n == 1                                   # config A: 8 MB in total, chunking would only add latency
    assert lib.fa2_plan_chunks(-1, S, D, mode, buf, 1024) == -1
