"""ctypes loader for libfa2_b200.so (the C ABI declared in include/fa2_b200.h).

The library is built in-tree by `make -C cuda-flash-attention_b200` (or
`__graft_entry__.build()`); there is NO fallback: if the shared object is missing, or no
B200 is visible when a compute entry point is called, the call fails loudly.
"""
from __future__ import annotations

import ctypes
import os

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG_ROOT, "libfa2_b200.so")

FA2_OK = 0
PRECISION = {"fp16": 0, "fp32": 1, "bf16": 2}
MODE = {"forward": 0, "backward": 1, "forward_backward": 2}

_fp = ctypes.c_void_p
_i = ctypes.c_int

# name -> (restype, argtypes); mirrors include/fa2_b200.h one to one
SIGNATURES = {
    "fa2_version": (_i, []),
    "fa2_last_error": (ctypes.c_char_p, []),
    "fa2_workspace_bytes": (ctypes.c_size_t, [_i] * 5),
    "fa2_forward": (_i, [_fp] * 5 + [_i] * 5 + [_fp]),
    "fa2_backward": (_i, [_fp] * 9 + [_i] * 5 + [_fp]),
    "fa2_forward_backward": (_i, [_fp] * 9 + [_i] * 5 + [_fp]),
    "fa2_host_forward": (_i, [_fp] * 5 + [_i] * 6 + [ctypes.POINTER(ctypes.c_float)]),
    "fa2_host_backward": (_i, [_fp] * 9 + [_i] * 6 + [ctypes.POINTER(ctypes.c_float)]),
    "fa2_host_forward_backward": (_i, [_fp] * 9 + [_i] * 6 + [ctypes.POINTER(ctypes.c_float)]),
    "fa2_partition": (_i, [_i, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "fa2_plan_chunks": (_i, [_i, _i, _i, _i, ctypes.POINTER(_i), _i]),
    "fa2_plan_split": (_i, [_i, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "fa2_seq_range": (_i, [_i, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "fa2_device_count": (_i, []),
    "fa2_release_workspaces": (_i, []),
    "fa2_profile_enable": (_i, [_i]),
    "fa2_profile_read": (_i, [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(_i)]),
    "fa2_profile_kernel_launches": (ctypes.c_longlong, []),
    "fa2_host_alloc": (ctypes.c_void_p, [ctypes.c_size_t]),
    "fa2_host_free": (None, [ctypes.c_void_p]),
}

_lib = None


class FA2Error(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libfa2_b200 error {code}: {message}")
        self.code = code


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} not found: build it with `make -C {_PKG_ROOT}` "
                "(there is no Python/CPU fallback for the FA2 path)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != FA2_OK:
        raise FA2Error(rc, load().fa2_last_error().decode("utf-8", "replace"))
