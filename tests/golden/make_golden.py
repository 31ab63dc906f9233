#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ FROM THE REFERENCE'S OWN PYTHON CODE.

Runs only in the build container (needs /root/reference); the fixtures it writes are
committed so that the GPU box (no /root/reference) can check against them.

What is imported from the reference, unmodified:
  * test_flash_attention2.FlashAttention2Tester.generate_test_data   (:177-195, torch.manual_seed(42) + torch.rand)
  * ...compute_reference                                             (:197-208)
  * ...compute_reference_backward                                    (:220-232, dO = ones)
  * generate_test_data.generate_test_data                            (CLI data: np.random.seed(42) + randn)
CuPy / matplotlib / seaborn are absent in this image and are stubbed in sys.modules only so
that the module imports (it sys.exit(1)s without CuPy, :16-23); no stub is ever called.
LSE has no function in the reference harness; it is the inline formula of :917-921.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

REF = os.environ.get("FA2_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))

for name in ("cupy", "cupy.cuda", "matplotlib", "matplotlib.pyplot", "seaborn"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["cupy"].cuda = sys.modules["cupy.cuda"]
sys.modules["cupy.cuda"].compiler = types.ModuleType("compiler")
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, REF)
import test_flash_attention2 as ref_harness  # noqa: E402
import generate_test_data as ref_datagen  # noqa: E402

# name -> (B, H, S, D); reduced-batch versions of the harness's named shapes (:1370-1408)
HARNESS_CASES = {
    "small1_B1_H1_S128_D64": (1, 1, 128, 64),       # Small-1
    "edge_nonpow2_B1_H2_S100_D64": (1, 2, 100, 64),  # Edge-NonPowerOf2 at B1 H2
    "edge_smallseq_B1_H2_S32_D64": (1, 2, 32, 64),   # Edge-SmallSeq at B1 H2
    "d32_B1_H1_S64_D32": (1, 1, 64, 32),
    "d128_B1_H1_S256_D128": (1, 1, 256, 128),
}


def harness_case(tester, B, H, S, D):
    cfg = ref_harness.TestConfig(name="g", batch_size=B, num_heads=H, seq_len=S, head_dim=D,
                                 test_both=True)
    Q, K, V = tester.generate_test_data(cfg)
    O = tester.compute_reference(Q, K, V)
    # LSE: the inline formula the harness uses in backward-only mode (:917-921)
    scores = torch.matmul(Q, K.transpose(-2, -1)) / (D ** 0.5)
    mx = scores.max(dim=-1, keepdim=True).values
    lse = (mx + torch.log(torch.exp(scores - mx).sum(dim=-1, keepdim=True))).squeeze(-1)
    grads = tester.compute_reference_backward(O, Q, K, V)
    f = lambda t: t.detach().numpy().astype(np.float32)
    return dict(Q=f(Q), K=f(K), V=f(V), O=f(O), LSE=f(lse),
                dQ=f(grads["dQ"]), dK=f(grads["dK"]), dV=f(grads["dV"]))


def main():
    tester = ref_harness.FlashAttention2Tester(test_mode="both", use_gpu_reference=False)
    for name, (B, H, S, D) in HARNESS_CASES.items():
        np.savez(os.path.join(OUT, f"harness_{name}.npz"), **harness_case(tester, B, H, S, D))
        print("wrote", name)

    # CLI data path: the generator's own bytes for a small folder, plus torch-reference outputs.
    with tempfile.TemporaryDirectory() as tmp:
        folder = ref_datagen.generate_test_data(1, 2, 64, 64, output_dir=tmp, seed=42)
        Q, K, V = (np.fromfile(os.path.join(folder, f"{n}.bin"), np.float32).reshape(1, 2, 64, 64)
                   for n in "QKV")
    Qt, Kt, Vt = (torch.from_numpy(a.copy()).requires_grad_(True) for a in (Q, K, V))
    O = tester.compute_reference(Qt, Kt, Vt)
    scores = torch.matmul(Qt, Kt.transpose(-2, -1)) / 8.0
    mx = scores.max(dim=-1, keepdim=True).values
    lse = (mx + torch.log(torch.exp(scores - mx).sum(dim=-1, keepdim=True))).squeeze(-1)
    grads = tester.compute_reference_backward(O, Qt, Kt, Vt)
    f = lambda t: t.detach().numpy().astype(np.float32)
    np.savez(os.path.join(OUT, "cli_B1_H2_S64_D64.npz"), Q=Q, K=K, V=V, O=f(O), LSE=f(lse),
             dQ=f(grads["dQ"]), dK=f(grads["dK"]), dV=f(grads["dV"]))
    print("wrote cli_B1_H2_S64_D64")


if __name__ == "__main__":
    main()
