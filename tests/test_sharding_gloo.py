"""World-size-2 gloo test of the N>1 host logic: every rank takes its fa2_partition slab range of a
global [B,H,S,D] batch, processes it independently (here with the CPU oracle standing in for the
device work -- this is a test of the sharding/gather logic, not of the kernels) and the gathered
result equals the unsharded one.  No data-path collective is involved; gloo only carries the check."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, shape, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "cuda-flash-attention_b200"))
    import fa2_b200
    from oracle import fa2_oracle as orc
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, H, S, D = shape
    rng = np.random.default_rng(99)                       # same global batch on every rank
    Q, K, V = (rng.standard_normal((B * H, S, D)).astype(np.float32) for _ in range(3))
    bh0, cnt = fa2_b200.partition(B * H, world, rank)
    sl = slice(bh0, bh0 + cnt)
    O, L = orc.forward(Q[sl][None], K[sl][None], V[sl][None])    # (1, cnt, S, D) == launch with B'=1, H'=cnt
    # gather: disjoint contiguous ranges of one buffer, exactly what fa2_host_* does with D2H copies
    full = torch.zeros(B * H, S, D)
    full[sl] = torch.from_numpy(O[0])
    dist.all_reduce(full)                                  # sum of disjoint ranges == concatenation
    counts = torch.zeros(B * H)
    counts[sl] = 1
    dist.all_reduce(counts)
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), full.numpy())
        np.save(os.path.join(out_dir, "counts.npy"), counts.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("shape", [(1, 5, 40, 64), (2, 2, 33, 32)])
def test_two_rank_slab_sharding_matches_unsharded(tmp_path, shape):
    from oracle import fa2_oracle as orc
    port = _free_port()
    mp.spawn(_worker, args=(2, port, shape, str(tmp_path)), nprocs=2, join=True)
    B, H, S, D = shape
    rng = np.random.default_rng(99)
    Q, K, V = (rng.standard_normal((B * H, S, D)).astype(np.float32) for _ in range(3))
    O, _ = orc.forward(Q[None], K[None], V[None])
    got = np.load(tmp_path / "gathered.npy")
    assert np.array_equal(np.load(tmp_path / "counts.npy"), np.ones(B * H))      # every slab exactly once
    assert np.abs(got - O[0]).max() < 1e-6
