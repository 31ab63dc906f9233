"""Shared helpers for the -m gpu parity tests (call the CUDA path through the C ABI)."""
import numpy as np
import torch

import fa2_b200
from oracle import fa2_oracle as orc

# north_star tolerances: max-abs <= 1e-2 on O/dQ/dK/dV for 16-bit operands, LSE <= 1e-3
TOL_O = 1e-2
TOL_GRAD = 1e-2
TOL_LSE = 1e-3


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def randn_case(shape, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    return tuple((rng.standard_normal(shape) * scale).astype(np.float32) for _ in range(4))


def gpu_forward(Q, K, V, precision="fp32"):
    O, L = fa2_b200.forward(dev(Q), dev(K), dev(V), precision=precision)
    torch.cuda.synchronize()
    return host(O), host(L)


def gpu_backward(Q, K, V, O, dO, L, precision="fp32"):
    g = fa2_b200.backward(dev(Q), dev(K), dev(V), dev(O), dev(dO), dev(L), precision=precision)
    torch.cuda.synchronize()
    return tuple(host(x) for x in g)


def maxerr(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max())
