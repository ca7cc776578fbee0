"""Sample sharding over the GPUs of one box (one process per GPU, torch.distributed).

Samples are independent (the reference's loop ``generate_fin_dataset.py:83-100`` carries no state), so the
batch is cut into contiguous, equal, rank-major shards and the ONLY communication is the final all-gather of
the observables plus a tiny all-reduce of running statistics.  Rank-major contiguous
shards make the gathered array come out in the original sample order bit for bit.

Backend: ``nccl`` on GPUs (NVLink/NVSwitch), ``gloo`` on CPU for the tests.
"""
from __future__ import annotations

import numpy as np

__all__ = ["shard_size", "shard_bounds", "gather_rows", "allreduce_moments", "chain_moments", "sharded_map"]


def shard_size(n_total: int, world: int) -> int:
    """Rows per rank (the last ranks may own fewer / zero real rows; the gather pads to this size)."""
    return (int(n_total) + world - 1) // world


def shard_bounds(n_total: int, world: int, rank: int):
    """Half-open range [lo, hi) of the samples owned by ``rank``."""
    per = shard_size(n_total, world)
    lo = min(rank * per, n_total)
    hi = min(lo + per, n_total)
    return lo, hi


def gather_rows(local, n_total: int, group=None):
    """All-gather row blocks produced from :func:`shard_bounds` shards into the full (n_total, ...) tensor, in
    the original order, on every rank.  ``local`` is a torch tensor (device for nccl, cpu for gloo)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    per = shard_size(n_total, world)
    tail = tuple(local.shape[1:])
    if local.shape[0] != per:                       # pad the short tail shards
        pad = torch.zeros((per,) + tail, dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
        local = pad
    out = torch.empty((world * per,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out[:n_total]


def allreduce_moments(local, group=None):
    """Chain statistics: (count, sum, sum of squares) per column, summed over ranks -> (count, mean, var)."""
    import torch
    import torch.distributed as dist
    x = local.to(torch.float64)
    acc = torch.cat([torch.full((1,), float(x.shape[0]), dtype=torch.float64, device=x.device),
                     x.sum(0), (x * x).sum(0)])
    dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    c = acc[0].item()
    m = x.shape[1]
    mean = acc[1:1 + m] / c
    var = acc[1 + m:] / c - mean * mean
    return int(c), mean, var


def chain_moments(count, col_sum, col_sq, group=None, device=None):
    """Chain statistics from per-rank accumulators (count, sum, sum of squares per observable, numpy): summed over the
    ranks when a process group is initialised -> ``(count, mean, var)`` as numpy.  Single-process: no communication."""
    col_sum, col_sq = np.asarray(col_sum, dtype=np.float64), np.asarray(col_sq, dtype=np.float64)
    acc = np.concatenate([[float(count)], col_sum, col_sq])
    try:
        import torch.distributed as dist
        active = dist.is_available() and dist.is_initialized()
    except Exception:
        active = False
    if active:
        import torch
        t = torch.from_numpy(acc)
        if device is None and dist.get_backend(group) == "nccl":
            device = torch.device("cuda", torch.cuda.current_device())
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        acc = t.cpu().numpy()
    c, m = acc[0], len(col_sum)
    mean = acc[1:1 + m] / c
    return int(c), mean, acc[1 + m:] / c - mean * mean


def sharded_map(fn, batch: np.ndarray, group=None, device=None):
    """Run ``fn(local_rows) -> (n_local, n_out) ndarray`` on this rank's shard of ``batch`` (numpy, identical on
    every rank) and return the gathered (N, n_out) ndarray on every rank."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(len(batch), world, rank)
    local = np.asarray(fn(batch[lo:hi]), dtype=np.float64)
    t = torch.from_numpy(np.ascontiguousarray(local))
    if device is not None:
        t = t.to(device)
    return gather_rows(t, len(batch), group).cpu().numpy()
