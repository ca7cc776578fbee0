"""Two-GPU test of the sharded path over NCCL (skipped on boxes with fewer than two GPUs): the gathered observables and
the chain statistics of a 2-rank run equal the single-GPU result bit for bit, in the original sample order."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs(n):
    rng = np.random.default_rng(77)
    return rng.uniform(0.1, 3.5, (101, 9)), np.linalg.qr(rng.standard_normal((n, 6)))[0]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import faulthandler
    faulthandler.dump_traceback_later(100, exit=True)      # a hung collective ends the test with a traceback
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from bayesianinferencedl_b200 import AffineROMFin, Fin, get_space, make_cov_chol
    from bayesianinferencedl_b200.bayesian_inference.likelihood import PCNChains
    from bayesianinferencedl_b200.dist import shard_bounds, sharded_map
    V = get_space(40, m=1)
    theta, phi = _inputs(V.dim())
    rom = AffineROMFin(V, None, phi, device=rank)
    q_fom = sharded_map(rom.forward_nine_param_qoi, theta, device=dev)
    q_rom = sharded_map(rom.forward_reduced_qoi, theta, device=dev)
    # chains 0..7 split 4 + 4: every chain is the Philox stream of its GLOBAL index
    fin = Fin(V, device=rank)
    chol = make_cov_chol(V, "m52", 1.6)
    data = rom.forward_nine_param_qoi(np.ones(9))
    lo, hi = shard_bounds(8, world, rank)
    out = PCNChains(fin, chol, data, 0.05, seed=5).run(4, n_chains=hi - lo, beta=0.2, first_chain=lo)
    assert torch.cuda.current_device() == rank              # libtfin restores the caller's device (handle on cuda:0 above)
    summ = PCNChains.summarize(out, device=dev)             # all-reduced over the two ranks
    acc = sharded_map(lambda rows: out["accepted"][:, None].astype(np.float64), np.zeros((8, 1)), device=dev)
    qs = sharded_map(lambda rows: out["qoi_sum"], np.zeros((8, 1)), device=dev)
    dist.destroy_process_group()
    if rank == 0:
        q.put((q_fom, q_rom, acc, qs, summ))


def test_two_gpu_shards_equal_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    import queue
    res = None
    for _ in range(150):                                     # fail fast if a worker died instead of waiting out the queue
        try:
            res = q.get(timeout=1)
            break
        except queue.Empty:
            assert all(p.is_alive() or p.exitcode == 0 for p in procs), "a worker crashed"
    assert res is not None, "timed out"
    q_fom, q_rom, acc, qs, summ = res
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    from bayesianinferencedl_b200 import AffineROMFin, Fin, get_space, make_cov_chol
    from bayesianinferencedl_b200.bayesian_inference.likelihood import PCNChains
    V = get_space(40, m=1)
    theta, phi = _inputs(V.dim())
    rom = AffineROMFin(V, None, phi)
    assert np.array_equal(q_fom, rom.forward_nine_param_qoi(theta))       # bit-exact, original order
    assert np.array_equal(q_rom, rom.forward_reduced_qoi(theta))
    fin = Fin(V)
    data = rom.forward_nine_param_qoi(np.ones(9))
    one = PCNChains(fin, make_cov_chol(V, "m52", 1.6), data, 0.05, seed=5).run(4, n_chains=8, beta=0.2)
    assert np.array_equal(acc[:, 0], one["accepted"]) and np.array_equal(qs, one["qoi_sum"])
    ref = PCNChains.summarize(one)
    assert summ["count"] == ref["count"] == 32 and np.allclose(summ["qoi_mean"], ref["qoi_mean"], rtol=1e-13)
