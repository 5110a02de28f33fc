"""World-size-2 test of the host-side logic of the multi-GPU path, on CPU with the gloo backend.

The device path exchanges the 8 virtual-shard statistic vectors with one all-gather and sums them in the fixed
tree ((0+1)+(2+3))+((4+5)+(6+7)) on every rank (include/mlb200.h "Sharding").  This test replays that protocol
in numpy over torch.distributed/gloo: every rank takes the point range mlb_shard_range() gives it, forms the
per-virtual-shard sufficient statistics of one M-step, all-gathers them, applies the tree, and must end with
bit-identical results on both ranks that also equal the single-process result."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def tree8(v):
    return ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]))


def vshard_stats(data, resp, n_total, lo, hi):
    """Statistics (count, first and second moments) of the virtual shards covering [lo, hi): list of vectors."""
    from ml_b200 import cabi
    out = []
    for v in range(8):
        b, e = cabi.shard_range(n_total, 8, v)
        if b < lo or e > hi:
            continue
        x, r = data[b - lo:e - lo], resp[b - lo:e - lo]
        out.append(np.concatenate([r.sum(axis=0), (r.T @ x).ravel(), np.einsum("ik,ia,ib->kab", r, x, x).ravel()]))
    return out


def worker(rank, world, port, n_total, result_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ml_b200 import cabi
    rng = np.random.default_rng(5)
    data = rng.normal(size=(n_total, 3))
    resp = rng.dirichlet(np.ones(4), size=n_total)
    lo, hi = cabi.shard_range(n_total, world, rank)
    mine = vshard_stats(data[lo:hi], resp[lo:hi], n_total, lo, hi)
    assert len(mine) == 8 // world
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    vectors = [v for part in gathered for v in part]
    total = tree8(vectors)
    np.save(os.path.join(result_dir, f"rank{rank}.npy"), total)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_agree_bitwise_with_one(tmp_path):
    import __graft_entry__
    __graft_entry__.build()
    n_total = 20011
    port = 29500 + os.getpid() % 2000
    mp.spawn(worker, args=(2, port, n_total, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    assert np.array_equal(a, b)
    rng = np.random.default_rng(5)
    data = rng.normal(size=(n_total, 3))
    resp = rng.dirichlet(np.ones(4), size=n_total)
    single = tree8(vshard_stats(data, resp, n_total, 0, n_total))
    assert np.array_equal(a, single)
    assert abs(single[:4].sum() - n_total) < 1e-6
