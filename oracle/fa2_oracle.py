"""Python face of the CPU oracle for the FA2 forward/backward path.

TEST INFRASTRUCTURE ONLY -- see the header of oracle/fa2_oracle.c.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

Three checkers live here:
  * ``forward`` / ``backward`` / ``rowdot``: ctypes calls into liboracle_fa2.so, the fp32
    restatement of the reference kernels (kernel_fa2_optimized.cu:19-347,
    f-attn2-backward.cu:33-380).
  * ``attention_fp64``: the same maths in float64 numpy, the "truth" that operand-rounding
    error of the sm_100a kernels is measured against
    (formulas of test_flash_attention2.py:197-208, :917-921 and the autograd grads of :220-232).
  * ``torch_reference``: the reference harness's own PyTorch CPU oracle restated
    (compute_reference + autograd with grad_output = dO).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_fa2.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile liboracle_fa2.so (and oracle/_ref when /root/reference exists)."""
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "fa2_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, os.path.join(_HERE, "liboracle_fa2.so")])
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB_PATH)
        fp = ctypes.POINTER(ctypes.c_float)
        lib.fa2_oracle_forward.argtypes = [fp] * 5 + [ctypes.c_int] * 4
        lib.fa2_oracle_forward.restype = ctypes.c_int
        lib.fa2_oracle_rowdot.argtypes = [fp] * 3 + [ctypes.c_long, ctypes.c_int]
        lib.fa2_oracle_rowdot.restype = ctypes.c_int
        lib.fa2_oracle_backward.argtypes = [fp] * 9 + [ctypes.c_int] * 4
        lib.fa2_oracle_backward.restype = ctypes.c_int
        _lib = lib
    return _lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def forward(Q, K, V):
    """fp32 restatement of the reference forward. Returns (O [B,H,S,D], LSE [B,H,S])."""
    B, H, S, D = Q.shape
    Q, q = _f32(Q); K, k = _f32(K); V, v = _f32(V)
    O = np.empty_like(Q); LSE = np.empty((B, H, S), np.float32)
    rc = _load().fa2_oracle_forward(q, k, v, _f32(O)[1], _f32(LSE)[1], B, H, S, D)
    if rc:
        raise RuntimeError(f"fa2_oracle_forward failed rc={rc}")
    return O, LSE


def rowdot(dO, O):
    B, H, S, D = O.shape
    dO, a = _f32(dO); O, b = _f32(O)
    out = np.empty((B, H, S), np.float32)
    rc = _load().fa2_oracle_rowdot(a, b, _f32(out)[1], B * H * S, D)
    if rc:
        raise RuntimeError("fa2_oracle_rowdot failed")
    return out


def backward(Q, K, V, O, dO, LSE):
    """fp32 restatement of the reference backward. Returns (dQ, dK, dV)."""
    B, H, S, D = Q.shape
    Q, q = _f32(Q); K, k = _f32(K); V, v = _f32(V); O, o = _f32(O); dO, g = _f32(dO); LSE, l = _f32(LSE)
    dQ = np.empty_like(Q); dK = np.empty_like(Q); dV = np.empty_like(Q)
    rc = _load().fa2_oracle_backward(q, k, v, o, g, l, _f32(dQ)[1], _f32(dK)[1], _f32(dV)[1], B, H, S, D)
    if rc:
        raise RuntimeError(f"fa2_oracle_backward failed rc={rc}")
    return dQ, dK, dV


def attention_fp64(Q, K, V, dO=None):
    """float64 truth: O, LSE (natural log) and, if dO is given, dQ, dK, dV."""
    Q = np.asarray(Q, np.float64); K = np.asarray(K, np.float64); V = np.asarray(V, np.float64)
    D = Q.shape[-1]
    S_ = np.einsum("bhqd,bhkd->bhqk", Q, K) / np.sqrt(D)
    m = S_.max(axis=-1, keepdims=True)
    E = np.exp(S_ - m)
    l = E.sum(axis=-1, keepdims=True)
    P = E / l
    O = np.einsum("bhqk,bhkd->bhqd", P, V)
    LSE = (np.log(l) + m)[..., 0]
    if dO is None:
        return O, LSE
    dO = np.asarray(dO, np.float64)
    dV = np.einsum("bhqk,bhqd->bhkd", P, dO)
    dP = np.einsum("bhqd,bhkd->bhqk", dO, V)
    Di = (dO * O).sum(axis=-1, keepdims=True)
    dS = P * (dP - Di) / np.sqrt(D)
    dQ = np.einsum("bhqk,bhkd->bhqd", dS, K)
    dK = np.einsum("bhqk,bhqd->bhkd", dS, Q)
    return O, LSE, dQ, dK, dV


def torch_reference(Q, K, V, dO=None):
    """The reference harness's PyTorch CPU oracle (test_flash_attention2.py:197-232):
    matmul -> /sqrt(D) -> softmax -> matmul, gradients by autograd (dO = ones when None)."""
    import torch
    import torch.nn.functional as F

    q = torch.from_numpy(np.ascontiguousarray(Q, np.float32)).requires_grad_(True)
    k = torch.from_numpy(np.ascontiguousarray(K, np.float32)).requires_grad_(True)
    v = torch.from_numpy(np.ascontiguousarray(V, np.float32)).requires_grad_(True)
    scores = torch.matmul(q, k.transpose(-2, -1)) / (q.shape[-1] ** 0.5)
    out = torch.matmul(F.softmax(scores, dim=-1), v)
    # LSE as the harness computes it for backward-only mode (:917-921)
    mx = scores.max(dim=-1, keepdim=True).values
    lse = (mx + torch.log(torch.exp(scores - mx).sum(dim=-1, keepdim=True))).squeeze(-1)
    g = torch.ones_like(out) if dO is None else torch.from_numpy(np.ascontiguousarray(dO, np.float32))
    out.backward(g)
    return (out.detach().numpy(), lse.detach().numpy(),
            q.grad.numpy(), k.grad.numpy(), v.grad.numpy())
