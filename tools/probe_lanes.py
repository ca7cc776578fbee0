"""D1 throughput against (resident warps, samples per warp): whole waves only, so wave quantisation does not blur it."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesianinferencedl_b200 import get_space, _cabi
from bayesianinferencedl_b200.assembly import build_operators

lanes_list = [int(a) for a in sys.argv[1:] if "=" not in a] or [32, 27, 24, 20, 16, 12, 8]
rr_list = [int(a[3:]) for a in sys.argv[1:] if a.startswith("rr=")] or [0]
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); st = ts.cuda_stream
ops = build_operators(get_space(40, m=3))
h = _cabi.TfinHandle(0)
h.set_operator(ops.row_ptr, ops.col_idx, ops.vals, ops.rhs); h.set_observation(*ops.obs_csr())
h.set_int("fom_solver", 2)
rng = np.random.default_rng(2)
for lanes, rr in [(l, r) for l in lanes_list for r in rr_list]:
    h.set_int("frontal_ring_rows", rr)
    h.set_int("frontal_lanes", lanes)
    th = torch.tensor(rng.uniform(0.1, 3.5, (64, 9)), device="cuda")
    q = torch.empty((64, 9), device="cuda", dtype=torch.float64)
    h.fom_affine_raw(th.data_ptr(), 64, 0, 1, 1e-12, 50000, qoi=q.data_ptr(), stream=st)
    occ, occ_b = h.get_int("frontal_ctas_per_sm"), h.get_int("frontal_bsub_ctas_per_sm")
    lanes = h.get_int("frontal_lanes")          # the automatic choice when 0 was asked for
    print("ring rows", h.get_int("frontal_ring_rows"), end="  ")
    for waves in (4,):
        N = 148 * occ * lanes * waves
        theta = torch.tensor(rng.uniform(0.1, 3.5, (N, 9)), device="cuda")
        q = torch.empty((N, 9), device="cuda", dtype=torch.float64)
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            h.fom_affine_raw(theta.data_ptr(), N, 0, 1, 1e-12, 50000, qoi=q.data_ptr(), stream=st)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"lanes={lanes:2d} occ={occ} occ_b={occ_b} in_flight/SM={occ * lanes:3d} waves={waves} N={N:6d} "
              f"{best:7.3f} ms  {N / best / 1e3:6.3f} M/s", flush=True)
h.close()
