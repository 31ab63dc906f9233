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


def _seq_worker(rank, world, port, shape, out_dir):
    """Sequence-split logic of fa2_host_* (csrc/fa2_api.cu host_dispatch_seqsplit) with numpy standing in for the
    kernels: rank r runs the forward on ITS query rows (K/V replicated), contributes D_i / LSE of those rows to an
    all-gather, runs the backward on ITS key/value rows against all query rows, and the partial dQ are summed
    (the one collective of the design) -- here an all-reduce over gloo, on the device a reduce-scatter over NVLink."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "cuda-flash-attention_b200"))
    import fa2_b200
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    BH, S, D = shape
    rng = np.random.default_rng(7)
    Q, K, V, dO = (rng.standard_normal((BH, S, D)) for _ in range(4))
    os.environ["FA2_SEQ_SPLIT"] = str(world)                       # (a 1024-row slab is one round of work: not split by default)
    g_bh, g_s = fa2_b200.plan_split(BH, S, world)
    assert (g_bh, g_s) == (1, world)
    r0, r1 = fa2_b200.seq_range(S, g_s, rank)
    sc = 1.0 / np.sqrt(D)
    # forward on the own query rows
    s_own = np.einsum("hqd,hkd->hqk", Q[:, r0:r1], K) * sc
    lse_own = np.log(np.exp(s_own - s_own.max(-1, keepdims=True)).sum(-1)) + s_own.max(-1)
    O_own = np.einsum("hqk,hkd->hqd", np.exp(s_own - lse_own[..., None]), V)
    # all-gather of D_i and LSE (disjoint row ranges: a sum of zero-padded pieces)
    lse = torch.zeros(BH, S, dtype=torch.float64)
    delta = torch.zeros(BH, S, dtype=torch.float64)
    lse[:, r0:r1] = torch.from_numpy(lse_own)
    delta[:, r0:r1] = torch.from_numpy((dO[:, r0:r1] * O_own).sum(-1))
    dist.all_reduce(lse)
    dist.all_reduce(delta)
    lse, delta = lse.numpy(), delta.numpy()
    # backward on the own key/value rows against ALL query rows
    P = np.exp(np.einsum("hqd,hkd->hqk", Q, K[:, r0:r1]) * sc - lse[..., None])
    dV_own = np.einsum("hqk,hqd->hkd", P, dO)
    dS = P * (np.einsum("hqd,hkd->hqk", dO, V[:, r0:r1]) - delta[..., None]) * sc
    dK_own = np.einsum("hqk,hqd->hkd", dS, Q)
    dQ = torch.from_numpy(np.einsum("hqk,hkd->hqd", dS, K[:, r0:r1]))          # partial: this KV range only
    dist.all_reduce(dQ)                                                      # <- the dQ reduce
    full = {n: torch.zeros(BH, S, D, dtype=torch.float64) for n in ("O", "dK", "dV")}
    for n, a in (("O", O_own), ("dK", dK_own), ("dV", dV_own)):
        full[n][:, r0:r1] = torch.from_numpy(a)
        dist.all_reduce(full[n])
    if rank == 0:
        np.savez(os.path.join(out_dir, "seq.npz"), dQ=dQ.numpy(), lse=lse, **{n: t.numpy() for n, t in full.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sequence_split_with_dq_reduce_matches_unsplit(tmp_path):
    from oracle import fa2_oracle as orc
    shape = (1, 1024, 32)                                 # one (b,h) slab on two devices: rows 0..511 / 512..1023
    port = _free_port()
    mp.spawn(_seq_worker, args=(2, port, shape, str(tmp_path)), nprocs=2, join=True)
    BH, S, D = shape
    rng = np.random.default_rng(7)
    Q, K, V, dO = (rng.standard_normal((BH, S, D)) for _ in range(4))
    tO, tL, tdQ, tdK, tdV = orc.attention_fp64(Q[None], K[None], V[None], dO[None])
    z = np.load(tmp_path / "seq.npz")
    for n, want in (("O", tO), ("lse", tL), ("dQ", tdQ), ("dK", tdK), ("dV", tdV)):
        assert np.abs(z[n] - want[0]).max() < 1e-9, n


def test_split_planner_and_row_ranges():
    import fa2_b200
    os.environ.pop("FA2_SEQ_SPLIT", None)
    assert fa2_b200.plan_split(256, 4096, 8) == (8, 1)          # config C: plenty of slabs, plain slab split
    assert fa2_b200.plan_split(16, 16384, 8) == (8, 1)          # config D: two slabs per device
    assert fa2_b200.plan_split(12, 8192, 8) == (8, 1)           # a slab per device or more: never replicate inputs over PCIe
    assert fa2_b200.plan_split(4, 16384, 8) == (4, 1)           # 64 forward items per slab: one round on one device anyway
    assert fa2_b200.plan_split(1, 16384, 8) == (1, 1)
    assert fa2_b200.plan_split(1, 131072, 8) == (1, 8)          # 512 forward / 1024 backward items: split until a share is one round
    assert fa2_b200.plan_split(2, 65536, 8) == (2, 4)
    assert fa2_b200.plan_split(1, 300, 8) == (1, 1)             # too short to split: one device
    assert fa2_b200.plan_split(3, 4096, 1) == (1, 1)
    os.environ["FA2_SEQ_SPLIT"] = "8"
    assert fa2_b200.plan_split(16, 16384, 8) == (1, 8)          # forced (tools/seq_split_bench.py measures it)
    os.environ["FA2_SEQ_SPLIT"] = "2"
    assert fa2_b200.plan_split(16, 16384, 8) == (4, 2)
    os.environ.pop("FA2_SEQ_SPLIT", None)
    for S, parts in ((16384, 8), (4096, 2), (1000, 2), (700, 2)):
        ranges = [fa2_b200.seq_range(S, parts, p) for p in range(parts)]
        assert ranges[0][0] == 0 and ranges[-1][1] == S
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))         # contiguous, disjoint, complete
        assert all(r[0] % 256 == 0 for r in ranges)
