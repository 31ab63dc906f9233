"""-m gpu: the FlashAttention CLI end to end (files in, files out) and, when the unmodified reference
CLI was built into oracle/_ref/ (oracle/build_ref.sh), file-level parity against it on the same inputs."""
import os
import shutil
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "cuda-flash-attention_b200", "FlashAttention")
REF = os.path.join(ROOT, "oracle", "_ref", "FlashAttention_ref")


def make_dir(base, B, H, S, D, seed=42, with_dO=False):
    """generate_test_data.py semantics: np.random.seed(seed); randn Q, K, V in that order (:10,27-33)."""
    d = os.path.join(base, f"B{B}_H{H}_S{S}_D{D}")
    os.makedirs(d)
    np.random.seed(seed)
    for n in "QKV":
        np.random.randn(B, H, S, D).astype(np.float32).tofile(os.path.join(d, f"{n}.bin"))
    if with_dO:
        np.random.seed(seed + 1)
        np.random.randn(B, H, S, D).astype(np.float32).tofile(os.path.join(d, "dO.bin"))
    return d


def load(d, name, shape):
    return np.fromfile(os.path.join(d, name + ".bin"), np.float32).reshape(shape)


def run(exe, *argv):
    r = subprocess.run([exe, *argv], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr + r.stdout
    return r.stdout


@pytest.mark.parametrize("B,H,S,D,prec", [(2, 8, 512, 64, "fp32"), (1, 4, 100, 64, "fp16"), (1, 2, 300, 128, "fp32")])
def test_cli_forward_then_backward_files(tmp_path, B, H, S, D, prec):
    from oracle import fa2_oracle as orc
    d = make_dir(str(tmp_path), B, H, S, D)
    out = run(CLI, "fa2", "forward", prec, d)
    assert "Kernel execution completed:" in out and "Output saved successfully." in out
    shp = (B, H, S, D)
    Q, K, V = (load(d, n, shp) for n in "QKV")
    O, L = load(d, "O", shp), load(d, "logsumexp", (B, H, S))
    t = orc.attention_fp64(Q, K, V, np.ones(shp))
    assert np.abs(O - t[0]).max() < 1e-2 and np.abs(L - t[1]).max() < 1e-3
    run(CLI, "fa2", "backward", prec, d + "/")             # reads O.bin / logsumexp.bin written above; dO = 1
    for n, want in zip(("dQ", "dK", "dV"), t[2:]):
        assert np.abs(load(d, n, shp) - want).max() < 1e-2, n


def test_cli_forward_backward_with_dO_file_and_alias(tmp_path):
    from oracle import fa2_oracle as orc
    B, H, S, D = 1, 3, 257, 64
    d = make_dir(str(tmp_path), B, H, S, D, with_dO=True)
    run(CLI, "fa2", "both", "fp32", d)
    shp = (B, H, S, D)
    Q, K, V, dO = (load(d, n, shp) for n in ("Q", "K", "V", "dO"))
    t = orc.attention_fp64(Q, K, V, dO)
    for n, want in zip(("O", "dQ", "dK", "dV"), (t[0],) + t[2:]):
        assert np.abs(load(d, n, shp) - want).max() < 1e-2, n


def test_cli_streamed_chunks_equal_serial_order(tmp_path):
    """The CLI streams read -> compute -> write in chunks of slabs through three buffer sets; FA2_CLI_STREAM=0 is the
    reference's load-all / compute / save-all order (src/main.cpp:74-118).  Same files either way (dQ up to the
    reduce-add order), for all three modes, with and without dO.bin."""
    B, H, S, D = 2, 12, 512, 64                                  # 24 slabs of 128 KiB: 1 MiB chunks -> 3 chunks
    a = make_dir(str(tmp_path / "a"), B, H, S, D, with_dO=True)
    b = os.path.join(str(tmp_path / "b"), os.path.basename(a))
    shutil.copytree(a, b)
    env_s = dict(os.environ, FA2_CLI_CHUNK_MB="1")
    env_n = dict(os.environ, FA2_CLI_STREAM="0")
    shp = (B, H, S, D)
    for mode, names in (("forward", ("O", "logsumexp")), ("backward", ("dQ", "dK", "dV")),
                        ("forward_backward", ("O", "logsumexp", "dQ", "dK", "dV"))):
        r1 = subprocess.run([CLI, "fa2", mode, "fp32", a], capture_output=True, text=True, timeout=300, env=env_s)
        r2 = subprocess.run([CLI, "fa2", mode, "fp32", b], capture_output=True, text=True, timeout=300, env=env_n)
        assert r1.returncode == 0 and r2.returncode == 0, r1.stderr + r2.stderr
        assert "streamed in 3 chunk(s)" in r1.stdout and "serial load" in r2.stdout
        for n in names:
            x = load(a, n, (B, H, S) if n == "logsumexp" else shp)
            y = load(b, n, (B, H, S) if n == "logsumexp" else shp)
            if n == "dQ":
                assert np.abs(x - y).max() < 1e-5
            else:
                assert np.array_equal(x, y), (mode, n)


def test_cli_hands_baseline_methods_to_the_reference_cli(tmp_path):
    """fa1 / naive are comparison baselines of the reference (include/dispatcher.h:30-51): with the reference's own CLI
    available the call is handed over to it, otherwise it is refused with a message."""
    d = make_dir(str(tmp_path), 1, 2, 128, 64)
    if os.path.exists(REF):
        r = subprocess.run([CLI, "fa1", "forward", "fp32", d], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "comparison baseline" in r.stderr and "Kernel execution completed" in r.stdout
        Q, K, V = (load(d, n, (1, 2, 128, 64)) for n in "QKV")
        from oracle import fa2_oracle as orc
        assert np.abs(load(d, "O", (1, 2, 128, 64)) - orc.attention_fp64(Q, K, V)[0]).max() < 1e-4
    r = subprocess.run([CLI, "naive", "forward", "fp32", d], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, FA2_BASELINE_CLI="/nonexistent"))
    assert r.returncode == 1 and "comparison baseline" in r.stderr


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/FlashAttention_ref not built")
@pytest.mark.parametrize("B,H,S,D,prec", [(2, 8, 512, 64, "fp32"), (4, 16, 1024, 64, "fp16")])
def test_baseline_configs_at_full_size_against_the_reference_cli(tmp_path, B, H, S, D, prec):
    """BASELINE.json configs[0] and configs[1] at their FULL size: ours (with the config's precision flag) against the
    reference's fp32 fa2 kernels on the same files (its fp16-SHM mode keeps accumulators in half and is only a loose
    oracle)."""
    ours = make_dir(str(tmp_path / "ours"), B, H, S, D, with_dO=True)
    theirs = os.path.join(str(tmp_path / "ref"), os.path.basename(ours))
    shutil.copytree(ours, theirs)
    run(REF, "fa2", "forward_backward", "fp32", theirs)
    run(CLI, "fa2", "forward_backward", prec, ours)
    shp = (B, H, S, D)
    for n, tol, s in (("O", 1e-2, shp), ("logsumexp", 1e-3, (B, H, S)), ("dQ", 1e-2, shp), ("dK", 1e-2, shp), ("dV", 1e-2, shp)):
        a, b = load(ours, n, s), load(theirs, n, s)
        assert np.isfinite(a).all()
        assert np.abs(a - b).max() < tol, (n, float(np.abs(a - b).max()))


def test_cli_multi_gpu_flag_single_device_ok(tmp_path):
    d = make_dir(str(tmp_path), 1, 4, 128, 64)
    run(CLI, "fa2", "forward", "fp32", d, "--gpus", "1")


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/FlashAttention_ref not built")
@pytest.mark.parametrize("B,H,S,D", [(2, 8, 512, 64), (1, 4, 100, 64), (2, 2, 256, 32)])
def test_file_parity_with_unmodified_reference_cli(tmp_path, B, H, S, D):
    """Same Q.bin/K.bin/V.bin through the reference's own fa2 fp32 kernels (compiled for sm_100 as they
    lie) and through ours: O/dQ/dK/dV within 1e-2, logsumexp within 1e-3 (north_star)."""
    ours = make_dir(str(tmp_path / "ours"), B, H, S, D)
    theirs = os.path.join(str(tmp_path / "ref"), os.path.basename(ours))
    shutil.copytree(ours, theirs)
    run(REF, "fa2", "forward_backward", "fp32", theirs)
    run(CLI, "fa2", "forward_backward", "fp32", ours)
    shp = (B, H, S, D)
    for n, tol, s in (("O", 1e-2, shp), ("logsumexp", 1e-3, (B, H, S)), ("dQ", 1e-2, shp), ("dK", 1e-2, shp), ("dV", 1e-2, shp)):
        a, b = load(ours, n, s), load(theirs, n, s)
        assert np.isfinite(a).all()
        assert np.abs(a - b).max() < tol, n


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/FlashAttention_ref not built")
def test_c_oracle_matches_unmodified_reference_kernels(tmp_path):
    """Pins the CPU restatement to the reference's real kernels run on this GPU."""
    from oracle import fa2_oracle as orc
    B, H, S, D = 1, 2, 100, 64
    d = make_dir(str(tmp_path), B, H, S, D)
    run(REF, "fa2", "forward_backward", "fp32", d)
    shp = (B, H, S, D)
    Q, K, V = (load(d, n, shp) for n in "QKV")
    O, L = orc.forward(Q, K, V)
    dQ, dK, dV = orc.backward(Q, K, V, O, np.ones(shp, np.float32), L)
    assert np.abs(O - load(d, "O", shp)).max() < 1e-5
    assert np.abs(L - load(d, "logsumexp", (B, H, S))).max() < 1e-5
    for n, a in (("dQ", dQ), ("dK", dK), ("dV", dV)):
        assert np.abs(a - load(d, n, shp)).max() < 5e-5, n


REF_ANY_D = os.path.join(ROOT, "oracle", "_ref", "ref_any_d")


@pytest.mark.skipif(not os.path.exists(REF_ANY_D), reason="oracle/_ref/ref_any_d not built")
@pytest.mark.parametrize("B,H,S,D,with_dO", [(1, 4, 512, 128, False), (1, 2, 300, 128, True), (1, 2, 1024, 128, True),
                                              (2, 2, 200, 64, True)])
def test_parity_with_reference_kernels_instantiated_at_d128(tmp_path, B, H, S, D, with_dO):
    """The reference dispatcher refuses D=128 (include/dispatcher.h:226-227), but its kernel TEMPLATES compile for
    it: oracle/ref_any_d instantiates them as they lie (no copy, > 48 KB smem opt-in) -- the headline head dim
    checked against the reference's own kernel code on the same inputs."""
    ours = make_dir(str(tmp_path / "ours"), B, H, S, D, with_dO=with_dO)
    theirs = os.path.join(str(tmp_path / "ref"), os.path.basename(ours))
    shutil.copytree(ours, theirs)
    run(REF_ANY_D, theirs)
    run(CLI, "fa2", "forward_backward", "fp32", ours)
    shp = (B, H, S, D)
    for n, tol, s in (("O", 1e-2, shp), ("logsumexp", 1e-3, (B, H, S)), ("dQ", 1e-2, shp), ("dK", 1e-2, shp), ("dV", 1e-2, shp)):
        a, b = load(ours, n, s), load(theirs, n, s)
        assert np.isfinite(b).all(), f"reference {n} not finite"
        assert np.abs(a - b).max() < tol, (n, float(np.abs(a - b).max()))
