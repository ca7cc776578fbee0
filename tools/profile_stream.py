"""Short fixed-iteration run of the streaming PCG kernel on the refined mesh, for ncu (traffic vs algorithmic bytes)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesianinferencedl_b200 import get_space, _cabi
from bayesianinferencedl_b200.assembly import build_operators
V = get_space(40, m=26); ops = build_operators(V)
h = _cabi.TfinHandle(0)
h.set_operator(ops.row_ptr, ops.col_idx, ops.vals, ops.rhs); h.set_observation(*ops.obs_csr())
N, maxit = int(os.environ.get("PS_N", 1184)), int(os.environ.get("PS_MAXIT", 60))
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
theta = torch.tensor(np.random.default_rng(2).uniform(0.1, 10.0, (N, 9)), device="cuda")
qoi = torch.empty((N, 9), device="cuda", dtype=torch.float64); it = torch.empty(N, device="cuda", dtype=torch.int32)
for rep in range(2):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); h.fom_affine_raw(theta.data_ptr(), N, 0, 1, 1e-12, maxit, qoi=qoi.data_ptr(), iters=it.data_ptr(), stream=ts.cuda_stream); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1); iters = int(it.sum().item())
print(json.dumps({"kernel": "pcg_stream_kernel", "n": ops.n, "N": N, "tile": h.get_int("stream_tile"), "iters_total": iters, "ms": ms,
                  "algorithmic_bytes": 88.0 * ops.n * iters, "algorithmic_GBps": 88.0 * ops.n * iters / ms / 1e6}))
