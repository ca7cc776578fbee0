"""One launch of each sparse-direct kernel for ncu: D1 on the m=3 mesh (one group per resident warp) and D2 on m=26."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesianinferencedl_b200 import get_space, _cabi
from bayesianinferencedl_b200.assembly import build_operators

which = sys.argv[1:] or ["d1", "d2"]
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); st = ts.cuda_stream
for name, m, N in (("d1", 3, 296 * 32), ("d1x6", 3, 296 * 32 * 6), ("d2", 26, 148), ("d2mid", 8, 148 * 7)):
    if name not in which:
        continue
    ops = build_operators(get_space(40, m=m))
    h = _cabi.TfinHandle(0)
    h.set_operator(ops.row_ptr, ops.col_idx, ops.vals, ops.rhs); h.set_observation(*ops.obs_csr())
    h.set_int("fom_solver", 2)
    theta = torch.tensor(np.random.default_rng(2).uniform(0.1, 3.5, (N, 9)), device="cuda")
    q = torch.empty((N, 9), device="cuda", dtype=torch.float64)
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h.fom_affine_raw(theta.data_ptr(), N, 0, 1, 1e-12, 50000, qoi=q.data_ptr(), stream=st)
        e1.record(); torch.cuda.synchronize()
    print(f"{name}: m={m} N={N} kernel={h.get_int('frontal_kernel')} threads={h.get_int('frontal_threads')} "
          f"{e0.elapsed_time(e1):.2f} ms", flush=True)
    h.close()
