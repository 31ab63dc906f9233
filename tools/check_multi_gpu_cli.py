#!/usr/bin/env python
"""Run the FlashAttention CLI on the same folder with --gpus 1 and --gpus N and compare every output file."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "cuda-flash-attention_b200", "FlashAttention")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
with tempfile.TemporaryDirectory() as tmp:
    d = os.path.join(tmp, "B4_H16_S1024_D64")
    os.makedirs(d)
    np.random.seed(42)
    for name in "QKV":
        np.random.randn(4, 16, 1024, 64).astype(np.float32).tofile(f"{d}/{name}.bin")
    outs = {}
    for g in (1, n):
        r = subprocess.run([CLI, "fa2", "forward_backward", "fp32", d, "--gpus", str(g)], capture_output=True, text=True)
        print(g, "rc", r.returncode, [l for l in r.stdout.splitlines() if "Kernel" in l], r.stderr.strip()[-200:])
        outs[g] = {k: np.fromfile(f"{d}/{k}.bin", np.float32) for k in ("O", "logsumexp", "dQ", "dK", "dV")}
    diff = {k: float(np.abs(outs[1][k] - outs[n][k]).max()) for k in outs[1]}
    print("max |1-GPU - %d-GPU|:" % n, diff)
    assert all(v < 1e-5 for v in diff.values())
    print("OK")
