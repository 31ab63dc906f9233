"""Host-side mirror of the reference's operator surface for the fa2 path.

Reference interface mirrored (argument meaning and error behaviour):
  * RunFlashAttention(Q,K,V,O,lse,dO,dQ,dK,dV,B,H,S,D,dtype,method,mode,tm)  include/dispatcher.h:220-246
    -> run_flash_attention(...): host (numpy) tensors in, host tensors out, kernel seconds returned.
  * the harness's run_fa2_forward_kernel / run_cuda_fa2_backward_kernel      test_flash_attention2.py:252-313,:477-567
    -> forward(...) / backward(...): device tensors (torch CUDA, or anything with a CuPy-style
       .data.ptr) in and out, asynchronous on the current torch stream.
Unsupported combinations fail like the reference's dispatcher does (method fa1/naive have no
backward, dispatcher.h:74-83; other head dims: "Unsupported head dimension", :137), but as a
Python exception instead of exit(EXIT_FAILURE).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import numpy as np

from . import _lib
from ._lib import FA2Error, MODE, PRECISION, check

SUPPORTED_HEAD_DIMS = (32, 64, 128)


def _ptr(x) -> int:
    """Raw device pointer of a torch CUDA tensor or a CuPy array."""
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "data") and hasattr(x.data, "ptr"):
        return x.data.ptr
    raise TypeError(f"cannot take a device pointer from {type(x)!r}")


def _stream_ptr(stream=None) -> int:
    if stream is not None:
        return int(getattr(stream, "cuda_stream", getattr(stream, "ptr", stream)))
    import torch
    return torch.cuda.current_stream().cuda_stream


def _check_dev(*tensors):
    import torch
    for t in tensors:
        if isinstance(t, torch.Tensor):
            if not t.is_cuda:
                raise ValueError("device entry points take CUDA tensors (use host_* for numpy arrays)")
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("tensors must be contiguous float32 [B,H,S,D]")


def partition(BH: int, n_parts: int, part: int) -> Tuple[int, int]:
    """Slab range (bh0, count) of `part`; same arithmetic the library uses (fa2_partition)."""
    a, b = ctypes.c_int(), ctypes.c_int()
    check(_lib.load().fa2_partition(BH, n_parts, part, ctypes.byref(a), ctypes.byref(b)))
    return a.value, b.value


def forward(Q, K, V, precision: str = "fp32", stream=None, out=None):
    """O, LSE = FA2 forward on device tensors [B,H,S,D] fp32. LSE is natural-log [B,H,S]."""
    import torch
    _check_dev(Q, K, V)
    B, H, S, D = Q.shape
    O, LSE = out if out is not None else (torch.empty_like(Q), torch.empty((B, H, S), device=Q.device, dtype=torch.float32))
    with torch.cuda.device(Q.device):
        check(_lib.load().fa2_forward(_ptr(Q), _ptr(K), _ptr(V), _ptr(O), _ptr(LSE), B, H, S, D,
                                      PRECISION[precision], _stream_ptr(stream)))
    return O, LSE


def backward(Q, K, V, O, dO, LSE, precision: str = "fp32", stream=None, out=None):
    """dQ, dK, dV = FA2 backward. D_i = rowsum(dO*O) and the dQ zero-fill happen inside."""
    import torch
    _check_dev(Q, K, V, O, dO, LSE)
    B, H, S, D = Q.shape
    dQ, dK, dV = out if out is not None else (torch.empty_like(Q), torch.empty_like(Q), torch.empty_like(Q))
    with torch.cuda.device(Q.device):
        check(_lib.load().fa2_backward(_ptr(Q), _ptr(K), _ptr(V), _ptr(O), _ptr(dO), _ptr(LSE), _ptr(dQ), _ptr(dK),
                                       _ptr(dV), B, H, S, D, PRECISION[precision], _stream_ptr(stream)))
    return dQ, dK, dV


def forward_backward(Q, K, V, dO, precision: str = "fp32", stream=None, out=None):
    """O, LSE, dQ, dK, dV in one call; O/LSE never leave the device between the passes."""
    import torch
    _check_dev(Q, K, V, dO)
    B, H, S, D = Q.shape
    if out is None:
        out = (torch.empty_like(Q), torch.empty((B, H, S), device=Q.device, dtype=torch.float32),
               torch.empty_like(Q), torch.empty_like(Q), torch.empty_like(Q))
    O, LSE, dQ, dK, dV = out
    with torch.cuda.device(Q.device):
        check(_lib.load().fa2_forward_backward(_ptr(Q), _ptr(K), _ptr(V), _ptr(dO), _ptr(O), _ptr(LSE), _ptr(dQ),
                                               _ptr(dK), _ptr(dV), B, H, S, D, PRECISION[precision],
                                               _stream_ptr(stream)))
    return O, LSE, dQ, dK, dV


# --------------------------------------------------------------------------------------
# host-pointer surface == RunFlashAttention
# --------------------------------------------------------------------------------------
def _hp(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _host(a, name):
    a = np.asarray(a)
    if a.dtype != np.float32 or not a.flags["C_CONTIGUOUS"]:
        raise ValueError(f"{name} must be C-contiguous float32")
    return a


def run_flash_attention(Q, K, V, O=None, logsumexp=None, dO=None, *, method: str = "fa2",
                        mode: str = "forward", precision: str = "fp32", n_gpus: int = 1):
    """Host tensors in, host tensors out: the RunFlashAttention of include/dispatcher.h:220-246.

    mode 'forward'          -> (O, logsumexp), kernel_seconds
    mode 'backward'         -> (dQ, dK, dV), kernel_seconds         (needs O, logsumexp; dO defaults to ones,
                                                                      src/main.cpp:83-93)
    mode 'forward_backward' -> (O, logsumexp, dQ, dK, dV), kernel_seconds
    """
    if mode in ("both", "forward-backward"):
        mode = "forward_backward"
    if method not in ("fa2", "fa1", "naive"):
        raise ValueError("Error: Unknown compute method")
    if method != "fa2":
        # comparison baselines of the reference; deliberately not part of this library
        raise FA2Error(3, f"method '{method}' is a reference comparison baseline and is not provided "
                          "(only fa2 is implemented)")
    if mode not in MODE:
        raise ValueError(f"unknown mode {mode!r}")
    Q = _host(Q, "Q"); K = _host(K, "K"); V = _host(V, "V")
    B, H, S, D = Q.shape
    if D not in SUPPORTED_HEAD_DIMS:
        raise FA2Error(1, f"Error: Unsupported head dimension {D}")
    lib = _lib.load()
    ms = ctypes.c_float(0.0)
    prec = PRECISION[precision]
    if mode == "forward":
        O = np.empty_like(Q); L = np.empty((B, H, S), np.float32)
        check(lib.fa2_host_forward(_hp(Q), _hp(K), _hp(V), _hp(O), _hp(L), B, H, S, D, prec, n_gpus, ctypes.byref(ms)))
        return (O, L), ms.value * 1e-3
    if dO is None:
        dO = np.ones_like(Q)
    dO = _host(dO, "dO")
    dQ = np.empty_like(Q); dK = np.empty_like(Q); dV = np.empty_like(Q)
    if mode == "backward":
        if O is None or logsumexp is None:
            raise ValueError("backward mode needs O and logsumexp (O.bin / logsumexp.bin of a forward run)")
        O = _host(O, "O"); L = _host(logsumexp, "logsumexp")
        check(lib.fa2_host_backward(_hp(Q), _hp(K), _hp(V), _hp(O), _hp(dO), _hp(L), _hp(dQ), _hp(dK), _hp(dV),
                                    B, H, S, D, prec, n_gpus, ctypes.byref(ms)))
        return (dQ, dK, dV), ms.value * 1e-3
    O = np.empty_like(Q); L = np.empty((B, H, S), np.float32)
    check(lib.fa2_host_forward_backward(_hp(Q), _hp(K), _hp(V), _hp(dO), _hp(O), _hp(L), _hp(dQ), _hp(dK), _hp(dV),
                                        B, H, S, D, prec, n_gpus, ctypes.byref(ms)))
    return (O, L, dQ, dK, dV), ms.value * 1e-3
