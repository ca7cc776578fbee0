"""GPU probe: streaming PCG (K4) at different batch sizes per GPU -- how much of the 20 % gap to the HBM peak is tile
imbalance (a tile runs until its slowest sample converges; a launch until its slowest CTA finishes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesianinferencedl_b200 import get_space, _cabi
from bayesianinferencedl_b200.assembly import build_operators
V = get_space(40, m=26); ops = build_operators(V)
h = _cabi.TfinHandle(0)
h.set_operator(ops.row_ptr, ops.col_idx, ops.vals, ops.rhs); h.set_observation(*ops.obs_csr())
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); st = ts.cuda_stream
for N in (1184, 2368, 4736):
    theta = torch.tensor(np.random.default_rng(2).uniform(0.1, 10.0, (N, 9)), device="cuda")
    qoi = torch.empty((N, 9), device="cuda", dtype=torch.float64); it = torch.empty(N, device="cuda", dtype=torch.int32)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); h.fom_affine_raw(theta.data_ptr(), N, 0, 1, 1e-12, 50000, qoi=qoi.data_ptr(), iters=it.data_ptr(), stream=st); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1); its = it.cpu().numpy().astype(np.float64)
    gbs = 88.0 * ops.n * its.sum() / (ms * 1e-3) / 1e9
    tiles = its[: N // 8 * 8].reshape(-1, 8)
    print(f"N={N}: {N/ms*1e3:.1f} solves/s, {gbs:.0f} GB/s = {gbs/6454.9:.3f} of peak; iters mean {its.mean():.0f} max {its.max():.0f}; "
          f"mean over tiles of (tile max / tile mean) = {(tiles.max(1)/tiles.mean(1)).mean():.3f}; tile-max spread max/mean = {tiles.max(1).max()/tiles.max(1).mean():.3f}")
