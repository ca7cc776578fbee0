"""GPU probe: optional fp32 on-chip PCG (K1f) vs the fp64 kernel on the headline workload (five-param FOM)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesianinferencedl_b200 import get_space, AffineROMFin, _cabi
V = get_space(40)
phi = np.random.default_rng(0).standard_normal((1597, 8))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
k5 = np.random.default_rng(1).uniform(0.1, 1.0, (N, 5))
theta = torch.tensor(np.concatenate([k5, k5[:, 3::-1]], axis=1), device="cuda")
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); st = ts.cuda_stream
res = {}
for prec, tol in (("fp64", 1e-12), ("fp32", 1e-6), ("fp32", 1e-8), ("fp32", 1e-9), ("fp32", 1e-10), ("fp32", 1e-12), ("fp64", 1e-8), ("fp64", 1e-9)):
    rom = AffineROMFin(V, None, phi, precision=prec)
    h = rom.handle
    q = torch.empty((N, 9), device="cuda", dtype=torch.float64); it = torch.empty(N, device="cuda", dtype=torch.int32)
    stt = torch.empty(N, device="cuda", dtype=torch.int32)
    run = lambda: h.fom_affine_raw(theta.data_ptr(), N, 0, 1, tol, 20000, qoi=q.data_ptr(), iters=it.data_ptr(), status=stt.data_ptr(), stream=st)
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); run(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    res[(prec, tol)] = q.cpu().numpy()
    print(f"{prec} tol={tol:g}: {N/ms*1e3:.3e} solves/s, mean iters {it.double().mean().item():.1f}, ok {(stt==0).all().item()}, "
          f"T={h.get_int('pcg_threads')} R={h.get_int('pcg_rows_per_thread')} occ={h.get_int('pcg_ctas_per_sm')}")
ref = res[("fp64", 1e-12)]
for k, v in res.items():
    print(k, "max rel err of observables vs fp64/1e-12:", np.max(np.abs(v - ref) / np.abs(ref)))
