"""-m gpu parity: the sm_100a backward (and fused forward_backward) through the C ABI."""
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
def test_backward_vs_golden_harness_mode(U, path):
    """Harness '--mode backward' (test_flash_attention2.py:917-928): reference O and LSE in, dO = ones."""
    z = np.load(path)
    dQ, dK, dV = U.gpu_backward(z["Q"], z["K"], z["V"], z["O"], np.ones_like(z["O"]), z["LSE"])
    got = np.concatenate([dQ.ravel(), dK.ravel(), dV.ravel()])
    want = np.concatenate([z["dQ"].ravel(), z["dK"].ravel(), z["dV"].ravel()])
    assert np.isfinite(got).all()
    assert U.maxerr(got, want) < U.TOL_GRAD


@pytest.mark.parametrize("path", GOLDEN[:3], ids=[os.path.basename(p)[:-4] for p in GOLDEN[:3]])
def test_forward_backward_vs_golden_both_mode(U, path):
    """Harness '--mode both' (:608-728): our own O/LSE feed the backward."""
    import torch
    import fa2_b200
    z = np.load(path)
    outs = fa2_b200.forward_backward(U.dev(z["Q"]), U.dev(z["K"]), U.dev(z["V"]), U.dev(np.ones_like(z["Q"])))
    torch.cuda.synchronize()
    O, L, dQ, dK, dV = (U.host(t) for t in outs)
    assert U.maxerr(O, z["O"]) < U.TOL_O and U.maxerr(L, z["LSE"]) < U.TOL_LSE
    for a, n in ((dQ, "dQ"), (dK, "dK"), (dV, "dV")):
        assert U.maxerr(a, z[n]) < U.TOL_GRAD, n


SHAPES = [
    (1, 1, 128, 64), (2, 4, 256, 64), (2, 2, 512, 64), (1, 2, 1024, 64),
    (2, 3, 100, 64), (2, 3, 32, 64), (1, 2, 129, 64), (1, 2, 257, 64), (1, 1, 1000, 64),
    (1, 2, 1, 64), (1, 1, 383, 128), (1, 2, 512, 128), (2, 2, 256, 32), (1, 1, 77, 32), (1, 1, 2048, 128),
]


@pytest.mark.parametrize("shape", SHAPES, ids=["B%d_H%d_S%d_D%d" % s for s in SHAPES])
def test_backward_vs_fp64_truth_random_dO(U, shape):
    Q, K, V, dO = U.randn_case(shape, seed=21)
    tO, tL, tdQ, tdK, tdV = U.orc.attention_fp64(Q, K, V, dO)
    dQ, dK, dV = U.gpu_backward(Q, K, V, tO.astype(np.float32), dO, tL.astype(np.float32))
    assert U.maxerr(dQ, tdQ) < U.TOL_GRAD
    assert U.maxerr(dK, tdK) < U.TOL_GRAD
    assert U.maxerr(dV, tdV) < U.TOL_GRAD
    if shape[2] <= 257:                       # C restatement of the reference kernels on the same inputs
        cO, cL = U.orc.forward(Q, K, V)
        cdQ, cdK, cdV = U.orc.backward(Q, K, V, cO, dO, cL)
        assert max(U.maxerr(dQ, cdQ), U.maxerr(dK, cdK), U.maxerr(dV, cdV)) < U.TOL_GRAD


def test_backward_is_rerunnable_dq_zeroed_internally(U):
    """The reference needs dQ memset before every launch (test :546-548); ours zero-fills itself."""
    import torch
    import fa2_b200
    Q, K, V, dO = U.randn_case((1, 2, 300, 64), seed=4)
    tO, tL, tdQ, _, _ = U.orc.attention_fp64(Q, K, V, dO)
    q, k, v, o, g, l = (U.dev(x) for x in (Q, K, V, tO, dO, tL))
    out = tuple(torch.full_like(q, 123.0) for _ in range(3))
    for _ in range(3):
        fa2_b200.backward(q, k, v, o, g, l, out=out)
    torch.cuda.synchronize()
    assert U.maxerr(U.host(out[0]), tdQ) < U.TOL_GRAD


def test_backward_properties_full_size(U):
    """Larger sizes: linearity in dO, and an fp64 spot check of one head."""
    import torch
    import fa2_b200
    for (B, H, S, D) in [(2, 8, 512, 64), (1, 2, 4096, 128)]:
        g = torch.Generator(device="cuda").manual_seed(2)
        Q, K, V, G1, G2 = (torch.randn(B, H, S, D, device="cuda", generator=g) for _ in range(5))
        O, L = fa2_b200.forward(Q, K, V)
        a = fa2_b200.backward(Q, K, V, O, G1, L)
        b = fa2_b200.backward(Q, K, V, O, G2, L)
        c = fa2_b200.backward(Q, K, V, O, (G1 + G2).contiguous(), L)
        for x, y, z in zip(a, b, c):
            assert float((z - (x + y)).abs().max()) < 2e-2
        q, k, v, go = (t[0, 0].double() for t in (Q, K, V, G1))
        s = (q @ k.T) / (D ** 0.5)
        P = torch.softmax(s, -1)
        o = P @ v
        dV = P.T @ go
        dP = go @ v.T
        dS = P * (dP - (go * o).sum(-1, keepdim=True)) / (D ** 0.5)
        assert float((a[2][0, 0].double() - dV).abs().max()) < 1e-2
        assert float((a[0][0, 0].double() - dS @ k).abs().max()) < 1e-2
        assert float((a[1][0, 0].double() - dS.T @ q).abs().max()) < 1e-2


def test_host_api_backward_default_dO_is_ones(U):
    """CLI rule: no dO.bin -> dO = 1 (src/main.cpp:83-93)."""
    import fa2_b200
    Q, K, V, _ = U.randn_case((1, 2, 160, 64), seed=8)
    (O, L, dQ, dK, dV), secs = fa2_b200.run_flash_attention(Q, K, V, mode="forward_backward")
    t = U.orc.attention_fp64(Q, K, V, np.ones_like(Q))
    for got, want in zip((O, L, dQ, dK, dV), t):
        assert U.maxerr(got, want) < 1e-2
    (dQ2, dK2, dV2), _ = fa2_b200.run_flash_attention(Q, K, V, O, L, mode="backward")
    # the fused call forms D_i = rowsum(dO o O) from the unnormalised accumulator inside the forward epilogue, the
    # stand-alone backward from the stored fp32 O: an ulp of difference in D_i can flip a 16-bit rounding of dS
    assert U.maxerr(dK2, dK) < 2e-4 and U.maxerr(dV2, dV) < 1e-5 and U.maxerr(dQ2, dQ) < 2e-4


def test_host_api_chunked_pipeline_matches_device_api(U):
    """fa2_host_* streams the slabs through three buffer sets in chunks (H2D / kernels / D2H overlapped);
    80 slabs of S=1024 make several chunks (fa2_plan_chunks).  Results must equal the one-shot device path."""
    import torch
    import fa2_b200
    Q, K, V, dO = U.randn_case((1, 80, 1024, 64), seed=17)
    (O, L, dQ, dK, dV), secs = fa2_b200.run_flash_attention(Q, K, V, dO=dO, mode="forward_backward")
    q, k, v, g = (U.dev(x) for x in (Q, K, V, dO))
    o, l, dq, dk, dv = fa2_b200.forward_backward(q, k, v, g)
    torch.cuda.synchronize()
    assert np.array_equal(O, U.host(o)) and np.array_equal(L, U.host(l))          # forward is deterministic
    assert np.array_equal(dK, U.host(dk)) and np.array_equal(dV, U.host(dv))
    assert U.maxerr(dQ, U.host(dq)) < 1e-5                                          # reduce-add order differs
    assert secs > 0


def test_host_api_many_chunks_match_device_api(U):
    """48 slabs of S=4096 are streamed in about nine chunks (fa2_plan_chunks: small first and last chunk, larger copies in
    between); every slab must land where the one-shot device path puts it."""
    import torch
    import fa2_b200
    Q, K, V, dO = U.randn_case((1, 48, 4096, 64), seed=23)
    (O, L, dQ, dK, dV), _ = fa2_b200.run_flash_attention(Q, K, V, dO=dO, mode="forward_backward")
    o, l, dq, dk, dv = fa2_b200.forward_backward(*(U.dev(x) for x in (Q, K, V, dO)))
    torch.cuda.synchronize()
    assert np.array_equal(O, U.host(o)) and np.array_equal(L, U.host(l))
    assert np.array_equal(dK, U.host(dk)) and np.array_equal(dV, U.host(dv))
    assert U.maxerr(dQ, U.host(dq)) < 1e-5


def test_backward_bf16_operands_and_fp16_flag(U):
    """The explicit bf16 flag (wider range, coarser mantissa) and the reference's `fp16` SHM-precision flag."""
    Q, K, V, dO = U.randn_case((1, 2, 300, 128), seed=6)
    tO, tL, tdQ, tdK, tdV = U.orc.attention_fp64(Q, K, V, dO)
    for prec, tol in (("bf16", 3e-2), ("fp16", U.TOL_GRAD)):
        dQ, dK, dV = U.gpu_backward(Q, K, V, tO.astype(np.float32), dO, tL.astype(np.float32), precision=prec)
        assert max(U.maxerr(dQ, tdQ), U.maxerr(dK, tdK), U.maxerr(dV, tdV)) < tol, prec


def _fp64_truth_gpu(q, k, v, g):
    """All heads at once in float64 on the GPU: O, LSE, dQ, dK, dV."""
    import torch
    D = q.shape[-1]
    q, k, v, g = (t.double() for t in (q, k, v, g))
    s = torch.einsum("bhqd,bhkd->bhqk", q, k) / (D ** 0.5)
    lse = torch.logsumexp(s, -1)
    p = torch.exp(s - lse[..., None])
    o = p @ v
    dv = p.transpose(-1, -2) @ g
    dp = g @ v.transpose(-1, -2)
    ds = p * (dp - (g * o).sum(-1, keepdim=True)) / (D ** 0.5)
    return o, lse, ds @ k, ds.transpose(-1, -2) @ q, dv


@pytest.mark.parametrize("shape", [(2, 40, 640, 64), (1, 100, 384, 128), (1, 160, 200, 32)],
                         ids=lambda s: "B%d_H%d_S%d_D%d" % s)
def test_persistent_schedule_several_items_per_cta(U, shape):
    """More work items than SMs, so every persistent CTA walks several of them (barrier parities come from running
    counters), and S chosen so that the last forward item of each head has only ONE query tile (the second
    softmax warpgroup skips it) and the last KV / Q tile is ragged.  Checked against float64 for every head, for
    the stand-alone calls and the fused call."""
    import torch
    import fa2_b200
    gen = torch.Generator(device="cuda").manual_seed(31)
    q, k, v, g = (torch.randn(*shape, device="cuda", generator=gen) for _ in range(4))
    t = _fp64_truth_gpu(q, k, v, g)
    o, l = fa2_b200.forward(q, k, v)
    grads = fa2_b200.backward(q, k, v, o, g, l)
    fused = fa2_b200.forward_backward(q, k, v, g)
    torch.cuda.synchronize()
    for got in ((o, l) + tuple(grads), fused):
        for x, want, tol in zip(got, t, (U.TOL_O, U.TOL_LSE, U.TOL_GRAD, U.TOL_GRAD, U.TOL_GRAD)):
            assert float((x.double() - want).abs().max()) < tol


def test_long_sequence_config_D_slice(U):
    """Two heads of config D (S = 16384, D = 128): 128 KV steps per forward item, 128 Q steps per backward item."""
    import torch
    import fa2_b200
    gen = torch.Generator(device="cuda").manual_seed(5)
    q, k, v, g = (torch.randn(1, 2, 16384, 128, device="cuda", generator=gen) for _ in range(4))
    got = fa2_b200.forward_backward(q, k, v, g)
    torch.cuda.synchronize()
    want = _fp64_truth_gpu(q[:, :1], k[:, :1], v[:, :1], g[:, :1])
    for x, w, tol in zip(got, want, (U.TOL_O, U.TOL_LSE, U.TOL_GRAD, U.TOL_GRAD, U.TOL_GRAD)):
        assert float((x[:, :1].double() - w).abs().max()) < tol


def test_fused_forward_backward_bf16_operands(U):
    """The fused call with the explicit bf16 flag (forward epilogue writes D_i / LSE*log2e, donor warps cast dO as bf16)."""
    import torch
    import fa2_b200
    gen = torch.Generator(device="cuda").manual_seed(9)
    q, k, v, g = (torch.randn(1, 3, 700, 128, device="cuda", generator=gen) for _ in range(4))
    want = _fp64_truth_gpu(q, k, v, g)
    got = fa2_b200.forward_backward(q, k, v, g, precision="bf16")
    torch.cuda.synchronize()
    for x, w, tol in zip(got, want, (1e-2, 5e-3, 3e-2, 3e-2, 3e-2)):     # bf16: 8-bit mantissa (SURVEY finding F5)
        assert float((x.double() - w).abs().max()) < tol


def test_release_workspaces_then_reuse(U):
    """fa2_release_workspaces() frees the per-device arenas and the cached pipeline streams / events; the next calls
    (device API and host API) must rebuild them and give the same results."""
    import fa2_b200
    Q, K, V, dO = U.randn_case((1, 4, 300, 64), seed=12)
    (O1, L1, dQ1, dK1, dV1), _ = fa2_b200.run_flash_attention(Q, K, V, dO=dO, mode="forward_backward")
    fa2_b200._lib.check(fa2_b200._lib.load().fa2_release_workspaces())
    (O2, L2, dQ2, dK2, dV2), _ = fa2_b200.run_flash_attention(Q, K, V, dO=dO, mode="forward_backward")
    assert np.array_equal(O1, O2) and np.array_equal(L1, L2) and np.array_equal(dK1, dK2) and np.array_equal(dV1, dV2)
    assert U.maxerr(dQ1, dQ2) < 1e-5
    fa2_b200._lib.check(fa2_b200._lib.load().fa2_release_workspaces())
    O3, L3 = U.gpu_forward(Q, K, V)
    assert np.array_equal(O1, O3) and np.array_equal(L1, L3)


def test_host_api_two_gpus_match_one_gpu(U):
    """fa2_host_forward_backward with n_gpus = 2: each device takes a contiguous (b,h) slab range (fa2_partition), no
    collective; results must equal the single-GPU run (dQ up to reduce order).  Needs a 2-GPU box."""
    import torch
    import fa2_b200
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    Q, K, V, dO = U.randn_case((3, 5, 640, 128), seed=41)          # 15 slabs: 7 + 8
    assert fa2_b200.plan_split(15, 640, 2) == (2, 1)
    one, _ = fa2_b200.run_flash_attention(Q, K, V, dO=dO, mode="forward_backward", n_gpus=1)
    two, _ = fa2_b200.run_flash_attention(Q, K, V, dO=dO, mode="forward_backward", n_gpus=2)
    for a, b, exact in zip(one, two, (True, True, False, True, True)):
        if exact:
            assert np.array_equal(a, b)
        else:
            assert U.maxerr(a, b) < 5e-5                 # dQ: fp32 reduce-add order depends on the chunking


@pytest.mark.parametrize("shape", [(1, 2, 100, 64), (1, 2, 200, 128), (2, 1, 129, 32)], ids=lambda s: "B%d_H%d_S%d_D%d" % s)
def test_backward_ragged_kv_tile_with_strongly_negative_lse(U, shape):
    """Padded KV lanes of a ragged last tile see S^T = 0, so P = 2^(0 - lse2): with every score near -32 that is
    ~e^32 / S, an fp16 inf, and inf * 0 (zero-filled K row) used to poison whole dQ rows with NaN.  The compute
    warps now force P = dS = 0 in padded lanes."""
    B, H, S, D = shape
    rng = np.random.default_rng(77)
    u = rng.standard_normal(D).astype(np.float32)
    u *= np.sqrt(D) / np.linalg.norm(u)                                  # |u|^2 = D
    Q = (u + 0.05 * rng.standard_normal(shape)).astype(np.float32)
    K = (-4.0 * u + 0.05 * rng.standard_normal(shape)).astype(np.float32)   # scores ~ -4 sqrt(D): LSE << -10
    V, dO = (rng.standard_normal(shape).astype(np.float32) for _ in range(2))
    tO, tL, tdQ, tdK, tdV = U.orc.attention_fp64(Q, K, V, dO)
    assert tL.max() < -10
    dQ, dK, dV = U.gpu_backward(Q, K, V, tO.astype(np.float32), dO, tL.astype(np.float32))
    for got, want, n in ((dQ, tdQ, "dQ"), (dK, tdK, "dK"), (dV, tdV, "dV")):
        assert np.isfinite(got).all(), n
        assert U.maxerr(got, want) < U.TOL_GRAD, n
    import torch
    import fa2_b200
    outs = fa2_b200.forward_backward(*(U.dev(x) for x in (Q, K, V, dO)))
    torch.cuda.synchronize()
    for got, want, tol in zip(outs, (tO, tL, tdQ, tdK, tdV), (U.TOL_O, 5e-3, U.TOL_GRAD, U.TOL_GRAD, U.TOL_GRAD)):
        assert torch.isfinite(got).all()
        assert U.maxerr(U.host(got), want) < tol         # (|LSE| ~ 30 here: fp16 operand rounding scales with |score|)


def test_fused_call_accepts_16_byte_aligned_dO_view(U):
    """The fused forward reads dO rows with 256-bit loads; a dO view that is only 16-byte aligned must not fault:
    the library routes it through the stand-alone pre-pass."""
    import torch
    import fa2_b200
    shape = (1, 2, 300, 64)
    Q, K, V, dO = U.randn_case(shape, seed=14)
    buf = torch.zeros(dO.size + 4, device="cuda")
    g = buf[4:].view(shape)                                # 16 bytes into a 256-byte aligned allocation
    g.copy_(torch.from_numpy(dO))
    assert g.data_ptr() % 32 == 16 and g.is_contiguous()
    got = fa2_b200.forward_backward(U.dev(Q), U.dev(K), U.dev(V), g)
    torch.cuda.synchronize()
    want = U.orc.attention_fp64(Q, K, V, dO)
    for x, w, tol in zip(got, want, (U.TOL_O, U.TOL_LSE, U.TOL_GRAD, U.TOL_GRAD, U.TOL_GRAD)):
        assert U.maxerr(U.host(x), w) < tol


def test_concurrent_host_calls_on_one_device_are_serialised(U):
    """Two threads calling fa2_host_* on the same device share that device's arenas and streams: the library runs
    them one after the other (per-device lock), so both get the right answer even with different sizes."""
    import threading
    import fa2_b200
    cases = [U.randn_case((1, 6, 700, 64), seed=51), U.randn_case((2, 9, 1500, 128), seed=52),
             U.randn_case((1, 3, 260, 32), seed=53), U.randn_case((1, 12, 1100, 64), seed=54)]
    results = [None] * len(cases)

    def work(i):
        Q, K, V, dO = cases[i]
        results[i] = fa2_b200.run_flash_attention(Q, K, V, dO=dO, mode="forward_backward")[0]

    for _ in range(2):
        th = [threading.Thread(target=work, args=(i,)) for i in range(len(cases))]
        [t.start() for t in th]
        [t.join() for t in th]
        for (Q, K, V, dO), got in zip(cases, results):
            want = U.orc.attention_fp64(Q, K, V, dO)
            for x, w, tol in zip(got, want, (U.TOL_O, U.TOL_LSE, U.TOL_GRAD, U.TOL_GRAD, U.TOL_GRAD)):
                assert U.maxerr(x, w) < tol


class _FakeCuPy:
    """Duck-types the parts of a cupy.ndarray the binding touches (.data.ptr, .dtype, .flags, .shape, .device.id)
    on top of a torch CUDA tensor -- CuPy itself is not in this image."""

    class _Mem:
        def __init__(self, ptr):
            self.ptr = ptr

    class _Flags:
        c_contiguous = True

    class _Dev:
        def __init__(self, i):
            self.id = i

    def __init__(self, t):
        self.t = t
        self.data = self._Mem(t.data_ptr())
        self.dtype = "float32"
        self.flags = self._Flags()
        self.shape = tuple(t.shape)
        self.device = self._Dev(t.device.index or 0)


def test_cupy_style_arrays_go_through_the_data_ptr_branch(U, monkeypatch):
    """The harness of the reference passes CuPy arrays (test_flash_attention2.py:278-289): `.data.ptr` and an explicit
    stream handle are all the binding needs."""
    import torch
    import fa2_b200
    from fa2_b200 import api
    Q, K, V, dO = U.randn_case((1, 2, 200, 64), seed=61)
    tens = [U.dev(x) for x in (Q, K, V, dO)]
    outs_t = [torch.empty_like(tens[0]), torch.empty(1, 2, 200, device="cuda"), torch.empty_like(tens[0]),
              torch.empty_like(tens[0]), torch.empty_like(tens[0])]
    monkeypatch.setattr(api, "_on_device", lambda like, dev: torch.cuda.device(dev))      # (no cupy.cuda.Device here)
    q, k, v, g = (_FakeCuPy(t) for t in tens)
    out = tuple(_FakeCuPy(t) for t in outs_t)
    api.forward_backward(q, k, v, g, stream=torch.cuda.current_stream().cuda_stream, out=out)
    torch.cuda.synchronize()
    want = U.orc.attention_fp64(Q, K, V, dO)
    for x, w, tol in zip(outs_t, want, (U.TOL_O, U.TOL_LSE, U.TOL_GRAD, U.TOL_GRAD, U.TOL_GRAD)):
        assert U.maxerr(U.host(x), w) < tol
    with pytest.raises(ValueError):
        api.forward(q, k, _FakeCuPy(torch.empty(1, 2, 100, 64, device="cuda")), stream=0)


@pytest.mark.parametrize("shape,force", [((1, 1, 1024, 64), "2"), ((1, 3, 1536, 128), "2"), ((2, 1, 700, 32), "2")],
                         ids=["B1_H1_S1024_D64", "B1_H3_S1536_D128", "B2_H1_S700_D32"])
def test_host_api_sequence_split_two_gpus(U, shape, force, monkeypatch):
    """Sequence split (chosen for very long sequences with fewer slabs than devices; forced here with FA2_SEQ_SPLIT): the
    devices of a group split the ROWS -- forward on a range of
    query rows per device, backward on the same range of key/value rows, partial dQ summed by P2P loads over NVLink
    (the one collective of the design).  All three modes must reproduce the one-GPU results."""
    import torch
    import fa2_b200
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    if force:
        monkeypatch.setenv("FA2_SEQ_SPLIT", force)
    B, H, S, D = shape
    assert fa2_b200.plan_split(B * H, S, 2)[1] == 2
    Q, K, V, dO = U.randn_case(shape, seed=43)
    monkeypatch.delenv("FA2_SEQ_SPLIT", raising=False)
    one, _ = fa2_b200.run_flash_attention(Q, K, V, dO=dO, mode="forward_backward", n_gpus=1)
    if force:
        monkeypatch.setenv("FA2_SEQ_SPLIT", force)
    two, ms = fa2_b200.run_flash_attention(Q, K, V, dO=dO, mode="forward_backward", n_gpus=2)
    assert ms > 0
    truth = U.orc.attention_fp64(Q, K, V, dO)
    for a, b, t, n in zip(one, two, truth, ("O", "LSE", "dQ", "dK", "dV")):
        assert np.isfinite(b).all(), n
        assert U.maxerr(b, t) < (U.TOL_LSE if n == "LSE" else U.TOL_GRAD), n
        assert U.maxerr(a, b) < 2e-4, n                     # same kernels; D_i comes from the pre-pass instead of the fused epilogue
    (O2, L2), _ = fa2_b200.run_flash_attention(Q, K, V, mode="forward", n_gpus=2)
    assert np.array_equal(O2, one[0]) and np.array_equal(L2, one[1])          # forward row ranges: same tiles, same bits
    (dQ3, dK3, dV3), _ = fa2_b200.run_flash_attention(Q, K, V, one[0], one[1], dO=dO, mode="backward", n_gpus=2)
    for got, t, n in zip((dQ3, dK3, dV3), truth[2:], ("dQ", "dK", "dV")):
        assert U.maxerr(got, t) < U.TOL_GRAD, n
