"""world_size-2 test of the sharding + gather path on CPU (gloo): the gathered observables equal the single-process
result bit for bit and in the original sample order."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_forward(rows):
    """Stand-in for the per-rank solver: a deterministic row-wise map (the real one needs a GPU)."""
    return np.stack([np.sin(rows.sum(axis=1) * (j + 1)) for j in range(9)], axis=1) if len(rows) else np.zeros((0, 9))


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bayesianinferencedl_b200.dist import allreduce_moments, chain_moments, shard_bounds, sharded_map
    batch = np.random.default_rng(5).uniform(0.1, 3.5, (n_total, 9))
    full = sharded_map(_fake_forward, batch)
    lo, hi = shard_bounds(n_total, world, rank)
    cnt, mean, var = allreduce_moments(torch.from_numpy(_fake_forward(batch[lo:hi])))
    loc = _fake_forward(batch[lo:hi])                   # per-rank chain accumulators (count, sum, sum of squares)
    c2, m2, v2 = chain_moments(len(loc), loc.sum(0), (loc * loc).sum(0))
    assert c2 == cnt and np.allclose(m2, mean.numpy()) and np.allclose(v2, var.numpy(), atol=1e-12)
    dist.destroy_process_group()
    q.put((rank, full, cnt, mean.numpy(), var.numpy()))


@pytest.mark.parametrize("n_total", [11, 64, 1])
def test_two_rank_gather_is_bit_exact(n_total):
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    batch = np.random.default_rng(5).uniform(0.1, 3.5, (n_total, 9))
    ref = _fake_forward(batch)
    for rank, full, cnt, mean, var in res:
        assert full.shape == ref.shape
        assert np.array_equal(full, ref)                      # bit-exact, original order, on every rank
        assert cnt == n_total
        assert np.allclose(mean, ref.mean(axis=0)) and np.allclose(var, ref.var(axis=0), atol=1e-12)
