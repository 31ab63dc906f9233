"""-m gpu: the `fp32` precision flag keeps fp32 RANGE although the tensor cores see fp16 operands.

The reference's fp32 kernels (kernel_fa2_optimized.cu:19-347, f-attn2-backward.cu:243-266) take any fp32 input.  fp16
operands overflow at 65504 and lose precision below 6e-5, so the library measures max|x| of Q, K, V, dO while it
casts them and, when a tensor does not fit, re-casts it with a power-of-two scale whose inverse is folded into the
softmax scale and the epilogues (csrc/fa2_prepass.cu range_fix_*).  Checked here by RELATIVE error against float64:
||err||_inf / ||ref||_inf <= 1e-2 (north_star's 16-bit tolerance, made scale-free)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL = 1e-2


@pytest.fixture(scope="module")
def U():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device; the FA2 path has no CPU fallback")
    from tests import gpu_util
    return gpu_util


def rel(got, want):
    want = np.asarray(want, np.float64)
    return float(np.abs(np.asarray(got, np.float64) - want).max() / max(np.abs(want).max(), 1e-300))


CASES = {
    # mean-reduced loss: dO ~ 1 / (B S D); below the fp16 normal range, partly below its subnormals
    "dO_1e-7": dict(dO=1e-7),
    # loss scaling: dO * 2^16 overflows fp16 as it is, and so does the 16-bit (dP - D_i)
    "dO_2^16": dict(dO=65536.0),
    # |Q| beyond 65504 with |K| tiny: the scores are ordinary, the operands are not
    "Q_1e6_K_1e-6": dict(Q=1e6, K=1e-6),
    "V_1e6_dO_1e-3": dict(V=1e6, dO=1e-3),
    "V_1e-8_dO_1e-8": dict(V=1e-8, dO=1e-8),
    "Q_1e-5_K_1e5_V_300": dict(Q=1e-5, K=1e5, V=300.0),
}


@pytest.mark.parametrize("shape", [(1, 2, 300, 128), (2, 2, 200, 64)], ids=lambda s: "B%d_H%d_S%d_D%d" % s)
@pytest.mark.parametrize("case", sorted(CASES))
def test_fp32_flag_keeps_fp32_range(U, case, shape):
    import torch
    import fa2_b200
    mul = CASES[case]
    Q, K, V, dO = U.randn_case(shape, seed=71)
    Q, K, V, dO = (np.float32(mul.get(n, 1.0)) * x for n, x in zip(("Q", "K", "V", "dO"), (Q, K, V, dO)))
    tO, tL, tdQ, tdK, tdV = U.orc.attention_fp64(Q, K, V, dO)
    # fused call
    outs = fa2_b200.forward_backward(*(U.dev(x) for x in (Q, K, V, dO)))
    torch.cuda.synchronize()
    O, L, dQ, dK, dV = (U.host(t) for t in outs)
    for got, want, n in ((O, tO, "O"), (dQ, tdQ, "dQ"), (dK, tdK, "dK"), (dV, tdV, "dV")):
        assert np.isfinite(got).all(), (case, n)
        assert rel(got, want) <= REL, (case, n, rel(got, want))
    assert U.maxerr(L, tL) < U.TOL_LSE
    # stand-alone forward, then backward on the float64 O / LSE
    O2, L2 = U.gpu_forward(Q, K, V)
    assert rel(O2, tO) <= REL and U.maxerr(L2, tL) < U.TOL_LSE
    g = U.gpu_backward(Q, K, V, tO.astype(np.float32), dO, tL.astype(np.float32))
    for got, want, n in zip(g, (tdQ, tdK, tdV), ("dQ", "dK", "dV")):
        assert np.isfinite(got).all(), (case, n)
        assert rel(got, want) <= REL, (case, n, rel(got, want))
    # the reference's own `fp16` SHM-precision flag takes the same path
    g16 = U.gpu_backward(Q, K, V, tO.astype(np.float32), dO, tL.astype(np.float32), precision="fp16")
    assert max(rel(a, b) for a, b in zip(g16, (tdQ, tdK, tdV))) <= REL


@pytest.mark.parametrize("case", ["dO_1e-7", "Q_1e6_K_1e-6", "V_1e6_dO_1e-3"])
def test_fp32_range_on_the_large_problem_path(U, case):
    """Problems above ~1.2M elements per tensor take the cast + last-block decision + re-cast kernels (and the fused
    forward's donor warps for dO) instead of the single cooperative launch of the small path."""
    import torch
    import fa2_b200
    mul = CASES[case]
    shape = (1, 24, 1024, 128)
    gen = torch.Generator(device="cuda").manual_seed(75)
    q, k, v, g = (torch.randn(*shape, device="cuda", generator=gen) * mul.get(n, 1.0) for n in ("Q", "K", "V", "dO"))
    q, k, v, g = (t.contiguous() for t in (q, k, v, g))
    D = shape[-1]
    qd, kd, vd, gd = (t.double() for t in (q, k, v, g))
    s_ = torch.einsum("bhqd,bhkd->bhqk", qd, kd) / D ** 0.5
    lse = torch.logsumexp(s_, -1)
    p_ = torch.exp(s_ - lse[..., None])
    o = p_ @ vd
    dv = p_.transpose(-1, -2) @ gd
    ds = p_ * (gd @ vd.transpose(-1, -2) - (gd * o).sum(-1, keepdim=True)) / D ** 0.5
    want = (o, lse, ds @ kd, ds.transpose(-1, -2) @ qd, dv)
    fused = fa2_b200.forward_backward(q, k, v, g)
    o2, l2 = fa2_b200.forward(q, k, v)
    grads = fa2_b200.backward(q, k, v, o2, g, l2)
    torch.cuda.synchronize()
    for got in (fused, (o2, l2) + tuple(grads)):
        for x, w, n in zip(got, want, ("O", "LSE", "dQ", "dK", "dV")):
            assert torch.isfinite(x).all(), (case, n)
            if n == "LSE":
                assert float((x.double() - w).abs().max()) < U.TOL_LSE
            else:
                assert float((x.double() - w).abs().max() / w.abs().max()) <= REL, (case, n)


def test_range_scaling_leaves_ordinary_inputs_bit_identical(U):
    """Inside the windows (randn / rand / ones data) no scale is applied: a call after an out-of-range call gives
    exactly what it gave before (the amax slots alternate and are cleared between calls)."""
    import torch
    import fa2_b200
    Q, K, V, dO = U.randn_case((1, 3, 260, 64), seed=72)
    dev = [U.dev(x) for x in (Q, K, V, dO)]
    a = [t.clone() for t in fa2_b200.forward_backward(*dev)]
    big = [U.dev(x) for x in (Q * 1e6, K * 1e-6, V * 1e5, dO * 1e-9)]
    for _ in range(3):
        fa2_b200.forward_backward(*big)
    b = fa2_b200.forward_backward(*dev)
    torch.cuda.synchronize()
    for x, y, n in zip(a, b, ("O", "LSE", "dQ", "dK", "dV")):
        if n == "dQ":
            assert float((x - y).abs().max()) < 1e-5          # reduce-add order
        else:
            assert torch.equal(x, y), n


def test_host_api_chunks_scale_independently(U):
    """fa2_host_* processes the slabs in chunks (fa2_plan_chunks); every chunk measures and scales its own tensors,
    so slab groups of very different magnitude are fine as long as they fall into different chunks.  (WITHIN one
    launch the scale is per tensor: slabs that differ by more than ~2^20 in one tensor lose the small ones, as any
    single-scale 16-bit copy must.)"""
    import ctypes
    import fa2_b200
    B, H, S, D = 1, 40, 1024, 64
    buf = (ctypes.c_int * 64)()
    n = fa2_b200.load().fa2_plan_chunks(B * H, S, D, 2, buf, 64)
    assert n >= 2
    c0 = buf[0]
    Q, K, V, dO = U.randn_case((B, H, S, D), seed=73)
    dO = (dO * np.float32(3e-8)).astype(np.float32)
    V[:, c0:] *= np.float32(1e5)                                  # the later chunks see a very different V
    (O, L, dQ, dK, dV), _ = fa2_b200.run_flash_attention(Q, K, V, dO=dO, mode="forward_backward")
    for lo, hi in ((0, c0), (c0, H)):
        t = U.orc.attention_fp64(Q[:, lo:hi], K[:, lo:hi], V[:, lo:hi], dO[:, lo:hi])
        for got, want, nm in zip((O, dQ, dK, dV), (t[0], t[2], t[3], t[4]), ("O", "dQ", "dK", "dV")):
            assert rel(got[:, lo:hi], want) <= REL, (nm, lo)
        assert U.maxerr(L[:, lo:hi], t[1]) < U.TOL_LSE


def test_large_scores_are_exact_on_the_rounded_operands(U):
    """Q, K = 30 randn: |scores| ~ 1e4 and the softmax is nearly one-hot.  No 16-bit (10-bit mantissa: fp16 or TF32)
    operand can resolve near-ties between such scores, so against float64 on the ORIGINAL inputs single rows flip
    (a property of the operand precision north_star prescribes, not of the range handling).  What the kernels do
    guarantee: exact fp32 accumulation of the ROUNDED operands -- checked against float64 on fp16-rounded Q, K, V, dO."""
    shape = (1, 2, 256, 128)
    Q, K, V, dO = U.randn_case(shape, seed=74)
    Q, K = Q * np.float32(30), K * np.float32(30)
    r16 = lambda x: x.astype(np.float16).astype(np.float32)
    tO, tL, tdQ, tdK, tdV = U.orc.attention_fp64(r16(Q), r16(K), r16(V), r16(dO))
    O, L = U.gpu_forward(Q, K, V)
    assert rel(O, tO) <= REL
    assert float(np.abs(L - tL).max()) <= 1e-3 + 2e-6 * float(np.abs(tL).max())
    dQ, dK, dV = U.gpu_backward(Q, K, V, tO.astype(np.float32), dO, tL.astype(np.float32))
    for got, want, n in zip((dQ, dK, dV), (tdQ, tdK, tdV), ("dQ", "dK", "dV")):
        assert np.isfinite(got).all()
        assert rel(got, want) <= 2e-2, (n, rel(got, want))    # dS of near-ties is itself rounded to 16 bit
