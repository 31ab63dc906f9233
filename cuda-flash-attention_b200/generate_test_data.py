#!/usr/bin/env python
"""Synthetic Q/K/V generator with the reference's interface and bytes (generate_test_data.py:6-84):
positional `batch_size num_heads seq_len head_dim [--output-dir data] [--seed 42]`; np.random.seed(seed) then
randn Q, K, V (in that order) as float32 into <output-dir>/B{B}_H{H}_S{S}_D{D}/{Q,K,V}.bin.
`--dO` additionally writes dO.bin (randn, seed+1) so that gradients are not the degenerate dO = 1 case."""
import argparse
import os

import numpy as np


def generate_test_data(batch_size, num_heads, seq_len, head_dim, output_dir="data", seed=42, with_dO=False):
    np.random.seed(seed)
    path = os.path.join(output_dir, f"B{batch_size}_H{num_heads}_S{seq_len}_D{head_dim}")
    os.makedirs(path, exist_ok=True)
    shape = (batch_size, num_heads, seq_len, head_dim)
    for name in ("Q", "K", "V"):
        np.random.randn(*shape).astype(np.float32).tofile(os.path.join(path, f"{name}.bin"))
    if with_dO:
        np.random.seed(seed + 1)
        np.random.randn(*shape).astype(np.float32).tofile(os.path.join(path, "dO.bin"))
    print(f"Done! Data saved to: {path}")
    return path


def main():
    p = argparse.ArgumentParser(description="Generate Q, K, V matrices for Flash Attention testing")
    for n in ("batch_size", "num_heads", "seq_len", "head_dim"):
        p.add_argument(n, type=int)
    p.add_argument("--output-dir", default="data")
    p.add_argument("--seed", type=int, default=42)
    p.add_argument("--dO", action="store_true")
    a = p.parse_args()
    if min(a.batch_size, a.num_heads, a.seq_len, a.head_dim) <= 0:
        p.error("All dimensions must be positive integers")
    generate_test_data(a.batch_size, a.num_heads, a.seq_len, a.head_dim, a.output_dir, a.seed, a.dO)


if __name__ == "__main__":
    main()
