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


def _is_cupy(x) -> bool:
    return hasattr(x, "data") and hasattr(x.data, "ptr") and not hasattr(x, "data_ptr")


def _ptr(x) -> int:
    """Raw device pointer of a torch CUDA tensor or a CuPy array."""
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if _is_cupy(x):
        return x.data.ptr
    raise TypeError(f"cannot take a device pointer from {type(x)!r}")


def _stream_ptr(stream=None, like=None) -> int:
    """cudaStream_t of `stream` (torch / CuPy stream object or a raw handle); None = the current stream of the
    framework that owns `like` (torch is imported only for torch tensors)."""
    if stream is not None:
        return int(getattr(stream, "cuda_stream", getattr(stream, "ptr", stream)))
    if like is not None and _is_cupy(like):
        import cupy
        return int(cupy.cuda.get_current_stream().ptr)
    import torch
    return torch.cuda.current_stream().cuda_stream


def _device_of(x) -> int:
    if hasattr(x, "data_ptr"):
        return x.device.index if x.device.index is not None else 0
    return int(x.device.id)                               # CuPy


def _check_dev(shape4, **tensors):
    """Every tensor must be a contiguous float32 device array on ONE device: [B,H,S,D] like Q, or [B,H,S]
    for the logsumexp.  Anything else would be an out-of-bounds device access inside the library."""
    dev = None
    for name, t in tensors.items():
        want = tuple(shape4[:3]) if name in ("LSE", "logsumexp") else tuple(shape4)
        if hasattr(t, "data_ptr"):                        # torch
            import torch
            if not t.is_cuda:
                raise ValueError(f"{name}: device entry points take CUDA tensors (use run_flash_attention for numpy arrays)")
            ok = t.dtype == torch.float32 and t.is_contiguous()
        elif _is_cupy(t):
            ok = str(t.dtype) == "float32" and bool(t.flags.c_contiguous)
        else:
            raise TypeError(f"{name}: expected a torch CUDA tensor or a CuPy array, got {type(t)!r}")
        if not ok:
            raise ValueError(f"{name} must be contiguous float32")
        if tuple(t.shape) != want:
            raise ValueError(f"{name} has shape {tuple(t.shape)}, expected {want}")
        d = _device_of(t)
        if dev is None:
            dev = d
        elif d != dev:
            raise ValueError(f"{name} lives on device {d}, Q on device {dev}")
    return dev


class _on_device:
    """Make `dev` current for the duration of the call (cudaSetDevice via the owning framework)."""

    def __init__(self, like, dev):
        self.like, self.dev, self.ctx = like, dev, None

    def __enter__(self):
        if _is_cupy(self.like):
            import cupy
            self.ctx = cupy.cuda.Device(self.dev)
        else:
            import torch
            self.ctx = torch.cuda.device(self.dev)
        return self.ctx.__enter__()

    def __exit__(self, *a):
        return self.ctx.__exit__(*a)


def _empty_like(x, shape=None):
    if _is_cupy(x):
        import cupy
        return cupy.empty(tuple(x.shape) if shape is None else shape, dtype=cupy.float32)
    import torch
    return torch.empty(tuple(x.shape) if shape is None else shape, device=x.device, dtype=torch.float32)


def partition(BH: int, n_parts: int, part: int) -> Tuple[int, int]:
    """Slab range (bh0, count) of `part`; same arithmetic the library uses (fa2_partition)."""
    a, b = ctypes.c_int(), ctypes.c_int()
    check(_lib.load().fa2_partition(BH, n_parts, part, ctypes.byref(a), ctypes.byref(b)))
    return a.value, b.value


def plan_split(BH: int, S: int, n_gpus: int) -> Tuple[int, int]:
    """(g_bh, g_s): groups of devices over the slabs x devices per group over the rows (fa2_plan_split)."""
    a, b = ctypes.c_int(), ctypes.c_int()
    check(_lib.load().fa2_plan_split(BH, S, n_gpus, ctypes.byref(a), ctypes.byref(b)))
    return a.value, b.value


def seq_range(S: int, parts: int, part: int) -> Tuple[int, int]:
    """Row range [r0, r1) of `part` in a sequence split over `parts` devices (fa2_seq_range)."""
    a, b = ctypes.c_int(), ctypes.c_int()
    check(_lib.load().fa2_seq_range(S, parts, part, ctypes.byref(a), ctypes.byref(b)))
    return a.value, b.value


def forward(Q, K, V, precision: str = "fp32", stream=None, out=None):
    """O, LSE = FA2 forward on device tensors [B,H,S,D] fp32. LSE is natural-log [B,H,S]."""
    if len(Q.shape) != 4:
        raise ValueError("Q must be [B,H,S,D]")
    B, H, S, D = Q.shape
    _check_dev(Q.shape, Q=Q, K=K, V=V)
    O, LSE = out if out is not None else (_empty_like(Q), _empty_like(Q, (B, H, S)))
    dev = _check_dev(Q.shape, Q=Q, K=K, V=V, O=O, LSE=LSE)
    with _on_device(Q, dev):
        check(_lib.load().fa2_forward(_ptr(Q), _ptr(K), _ptr(V), _ptr(O), _ptr(LSE), B, H, S, D,
                                      PRECISION[precision], _stream_ptr(stream, Q)))
    return O, LSE


def backward(Q, K, V, O, dO, LSE, precision: str = "fp32", stream=None, out=None):
    """dQ, dK, dV = FA2 backward. D_i = rowsum(dO*O) and the dQ zero-fill happen inside."""
    if len(Q.shape) != 4:
        raise ValueError("Q must be [B,H,S,D]")
    B, H, S, D = Q.shape
    _check_dev(Q.shape, Q=Q, K=K, V=V, O=O, dO=dO, LSE=LSE)
    dQ, dK, dV = out if out is not None else (_empty_like(Q), _empty_like(Q), _empty_like(Q))
    dev = _check_dev(Q.shape, Q=Q, K=K, V=V, O=O, dO=dO, LSE=LSE, dQ=dQ, dK=dK, dV=dV)
    with _on_device(Q, dev):
        check(_lib.load().fa2_backward(_ptr(Q), _ptr(K), _ptr(V), _ptr(O), _ptr(dO), _ptr(LSE), _ptr(dQ), _ptr(dK),
                                       _ptr(dV), B, H, S, D, PRECISION[precision], _stream_ptr(stream, Q)))
    return dQ, dK, dV


def forward_backward(Q, K, V, dO, precision: str = "fp32", stream=None, out=None):
    """O, LSE, dQ, dK, dV in one call; O/LSE never leave the device between the passes."""
    if len(Q.shape) != 4:
        raise ValueError("Q must be [B,H,S,D]")
    B, H, S, D = Q.shape
    _check_dev(Q.shape, Q=Q, K=K, V=V, dO=dO)
    if out is None:
        out = (_empty_like(Q), _empty_like(Q, (B, H, S)), _empty_like(Q), _empty_like(Q), _empty_like(Q))
    O, LSE, dQ, dK, dV = out
    dev = _check_dev(Q.shape, Q=Q, K=K, V=V, dO=dO, O=O, LSE=LSE, dQ=dQ, dK=dK, dV=dV)
    with _on_device(Q, dev):
        check(_lib.load().fa2_forward_backward(_ptr(Q), _ptr(K), _ptr(V), _ptr(dO), _ptr(O), _ptr(LSE), _ptr(dQ),
                                               _ptr(dK), _ptr(dV), B, H, S, D, PRECISION[precision],
                                               _stream_ptr(stream, Q)))
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
    if Q.ndim != 4:
        raise ValueError("Q must be [B,H,S,D]")
    B, H, S, D = Q.shape
    for name, a, want in (("K", K, Q.shape), ("V", V, Q.shape), ("O", O, Q.shape), ("dO", dO, Q.shape),
                          ("logsumexp", logsumexp, (B, H, S))):
        if a is not None and tuple(np.shape(a)) != tuple(want):
            raise ValueError(f"{name} has shape {tuple(np.shape(a))}, expected {tuple(want)}")
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
