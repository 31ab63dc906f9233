"""-m gpu parity: the sm_100a forward through the C ABI vs golden fixtures, the C oracle, fp64 truth."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


@pytest.fixture(scope="module")
def U():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device; the FA2 path has no CPU fallback")
    from tests import gpu_util
    return gpu_util


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_forward_vs_golden(U, path, precision):
    z = np.load(path)
    O, L = U.gpu_forward(z["Q"], z["K"], z["V"], precision)
    assert np.isfinite(O).all() and np.isfinite(L).all()
    assert U.maxerr(O, z["O"]) < U.TOL_O
    assert U.maxerr(L, z["LSE"]) < U.TOL_LSE


# reference harness shapes (test_flash_attention2.py:1370-1408) at reduced B*H plus D=32/128 and ragged tails
SHAPES = [
    (1, 1, 128, 64), (2, 4, 256, 64), (2, 2, 512, 64), (1, 2, 1024, 64),
    (2, 3, 100, 64), (2, 3, 32, 64), (1, 2, 129, 64), (1, 2, 257, 64), (1, 1, 1000, 64),
    (1, 2, 1, 64), (1, 1, 383, 128), (1, 2, 512, 128), (2, 2, 256, 32), (1, 1, 77, 32), (1, 1, 2048, 128),
]


@pytest.mark.parametrize("shape", SHAPES, ids=["B%d_H%d_S%d_D%d" % s for s in SHAPES])
def test_forward_vs_oracle_randn(U, shape):
    Q, K, V, _ = U.randn_case(shape, seed=11)
    O, L = U.gpu_forward(Q, K, V)
    tO, tL = U.orc.attention_fp64(Q, K, V)
    assert U.maxerr(O, tO) < U.TOL_O
    assert U.maxerr(L, tL) < U.TOL_LSE
    if shape[2] <= 512:                       # C restatement of the reference kernel (slow, small cases)
        cO, cL = U.orc.forward(Q, K, V)
        assert U.maxerr(O, cO) < U.TOL_O
        assert U.maxerr(L, cL) < U.TOL_LSE


def test_forward_bf16_operands(U):
    Q, K, V, _ = U.randn_case((1, 2, 300, 64), seed=5)
    O, L = U.gpu_forward(Q, K, V, "bf16")
    tO, tL = U.orc.attention_fp64(Q, K, V)
    assert U.maxerr(O, tO) < U.TOL_O
    assert U.maxerr(L, tL) < 5e-3             # bf16 Q/K rounding: SURVEY F5


def test_forward_large_scores_rescale_path(U):
    # scores with a growing running max force the lazy-rescale branch
    Q, K, V, _ = U.randn_case((1, 1, 640, 64), seed=9)
    K = (K * np.linspace(0.2, 4.0, 640, dtype=np.float32)[None, None, :, None]).astype(np.float32)
    O, L = U.gpu_forward(Q, K, V)
    tO, tL = U.orc.attention_fp64(Q, K, V)
    assert U.maxerr(O, tO) < U.TOL_O
    assert U.maxerr(L, tL) < 2e-3 * max(1.0, float(np.abs(tL).max()) / 10)


def test_forward_properties_full_size(U):
    """BASELINE config sizes: size-independent properties (V = const -> O = const; linearity in V)."""
    import torch
    import fa2_b200
    for (B, H, S, D) in [(2, 8, 512, 64), (1, 4, 4096, 128)]:
        g = torch.Generator(device="cuda").manual_seed(1)
        Q = torch.randn(B, H, S, D, device="cuda", generator=g)
        K = torch.randn(B, H, S, D, device="cuda", generator=g)
        V1 = torch.randn(B, H, S, D, device="cuda", generator=g)
        V2 = torch.randn(B, H, S, D, device="cuda", generator=g)
        ones = torch.full_like(V1, 0.75)
        Oc, Lc = fa2_b200.forward(Q, K, ones)
        assert float((Oc - 0.75).abs().max()) < 2e-3
        O1, L1 = fa2_b200.forward(Q, K, V1)
        O2, L2 = fa2_b200.forward(Q, K, V2)
        O12, _ = fa2_b200.forward(Q, K, (V1 + V2).contiguous())
        assert float((O12 - (O1 + O2)).abs().max()) < 1e-2
        assert float((L1 - L2).abs().max()) == 0.0           # LSE does not depend on V; deterministic
        # spot-check a few rows against an fp64 torch computation
        idx = torch.tensor([0, S // 3, S - 1], device="cuda")
        q = Q[0, 0, idx].double()
        s = (q @ K[0, 0].double().T) / (D ** 0.5)
        ref = torch.softmax(s, -1) @ V1[0, 0].double()
        assert float((O1[0, 0, idx].double() - ref).abs().max()) < 1e-2
        assert float((L1[0, 0, idx].double() - torch.logsumexp(s, -1)).abs().max()) < 1e-3


def test_host_api_forward_matches_device_api(U):
    import fa2_b200
    Q, K, V, _ = U.randn_case((2, 3, 200, 64), seed=3)
    (O, L), secs = fa2_b200.run_flash_attention(Q, K, V, mode="forward")
    dO, dL = U.gpu_forward(Q, K, V)
    assert np.array_equal(O, dO) and np.array_equal(L, dL)
    assert secs > 0


def test_unsupported_head_dim_fails_loudly(U):
    import fa2_b200
    x = np.zeros((1, 1, 16, 48), np.float32)
    with pytest.raises(fa2_b200.FA2Error):
        fa2_b200.run_flash_attention(x, x, x)


def test_misaligned_tensor_is_rejected_not_miscomputed(U):
    """float4 / TMA / reduce-add paths need 16-byte aligned tensors: an odd view must fail loudly."""
    import ctypes
    import torch
    import fa2_b200
    buf = torch.zeros(4 * 1 * 64 * 64 + 4, device="cuda")
    q = buf[1:1 + 64 * 64]                                   # 4-byte offset into the allocation
    ok = torch.zeros(1, 1, 64, 64, device="cuda")
    lse = torch.zeros(1, 1, 64, device="cuda")
    lib = fa2_b200.load()
    rc = lib.fa2_forward(ctypes.c_void_p(q.data_ptr()), ctypes.c_void_p(ok.data_ptr()), ctypes.c_void_p(ok.data_ptr()),
                         ctypes.c_void_p(ok.data_ptr()), ctypes.c_void_p(lse.data_ptr()), 1, 1, 64, 64, 1, None)
    assert rc == 1 and b"16-byte" in lib.fa2_last_error()
