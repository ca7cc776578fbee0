"""GPU probe: sparse-direct solver (D1 sample-per-thread, D2 sample-per-CTA) vs the PCG kernels on the bench workloads."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesianinferencedl_b200 import get_space, _cabi
from bayesianinferencedl_b200.assembly import build_operators

ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); st = ts.cuda_stream
def timeit(fn, reps=2, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def handle(m, cells=False):
    ops = build_operators(get_space(40, m=m))
    h = _cabi.TfinHandle(0)
    h.set_operator(ops.row_ptr, ops.col_idx, ops.vals, ops.rhs); h.set_observation(*ops.obs_csr())
    if cells: h.set_cells(ops.cells, ops.Ke)
    return ops, h

which = sys.argv[1:] or ["small", "nodal", "mid", "big"]
if "small" in which:
    ops, h = handle(3)
    N = 200000
    k5 = np.random.default_rng(1).uniform(0.1, 1.0, (N, 5))
    theta = torch.tensor(np.concatenate([k5, k5[:, 3::-1]], axis=1), device="cuda")
    q = torch.empty((N, 9), device="cuda", dtype=torch.float64)
    w = torch.empty((N, ops.n), device="cuda", dtype=torch.float64)
    print("m=3 program:", {k: h.get_int(k) for k in ("frontal_slots", "frontal_cmax", "frontal_nnz_factor", "frontal_pair_updates")})
    for solver, kern, name in ((1, 0, "PCG K1"), (2, 1, "direct D1"), (2, 2, "direct D2 qoi")):
        h.set_int("fom_solver", solver); h.set_int("frontal_kernel", kern)
        ms = timeit(lambda: h.fom_affine_raw(theta.data_ptr(), N, 0, 1, 1e-12, 20000, qoi=q.data_ptr(), stream=st))
        print(f"m=3 affine {name}: {N/ms*1e3:.3e} solves/s  ({ms:.2f} ms; kernel {h.get_int('frontal_kernel')} threads {h.get_int('frontal_threads')} "
              f"ctas/sm {h.get_int('frontal_ctas_per_sm')} smem {h.get_int('frontal_smem_bytes')})", flush=True)
    h.set_int("fom_solver", 2); h.set_int("frontal_kernel", 1)
    for split in (0, 1):
        h.set_int("frontal_split", split)
        ms = timeit(lambda: h.fom_affine_raw(theta.data_ptr(), N, 0, 1, 1e-12, 20000, qoi=q.data_ptr(), stream=st))
        print(f"m=3 affine direct D1 split={split}: {N/ms*1e3:.3e} solves/s (factor ctas/sm {h.get_int('frontal_ctas_per_sm')}, bsub ctas/sm "
              f"{h.get_int('frontal_bsub_ctas_per_sm')}, ring rows {h.get_int('frontal_ring_rows')})", flush=True)
    for lanes in (16, 32):
        h.set_int("frontal_lanes", lanes)
        ms = timeit(lambda: h.fom_affine_raw(theta.data_ptr(), N, 0, 1, 1e-12, 20000, qoi=q.data_ptr(), stream=st))
        print(f"m=3 affine direct D1 lanes={lanes}: {N/ms*1e3:.3e} solves/s (ctas/sm {h.get_int('frontal_ctas_per_sm')} smem {h.get_int('frontal_smem_bytes')} ring rows {h.get_int('frontal_ring_rows')})", flush=True)
    Nw = 50000
    ms = timeit(lambda: h.fom_affine_raw(theta.data_ptr(), Nw, 0, 1, 1e-12, 20000, qoi=q.data_ptr(), w=w.data_ptr(), stream=st))
    print(f"m=3 affine direct D1 with w out: {Nw/ms*1e3:.3e} solves/s", flush=True)
    h.close(); del w
if "nodal" in which:
    ops, h = handle(3, cells=True)
    N = 100000
    k = torch.exp(0.3 * torch.randn((N, ops.n), device="cuda", dtype=torch.float64))
    q = torch.empty((N, 9), device="cuda", dtype=torch.float64)
    for solver, kern, name in ((1, 0, "PCG K2"), (2, 1, "direct D1")):
        h.set_int("fom_solver", solver); h.set_int("frontal_kernel", kern)
        ms = timeit(lambda: h.fom_nodal_raw(k.data_ptr(), N, 1, 1e-12, 20000, qoi=q.data_ptr(), stream=st))
        print(f"m=3 nodal {name}: {N/ms*1e3:.3e} solves/s ({ms:.2f} ms)", flush=True)
    h.close(); del k
if "mid" in which:
    for m in (4, 5, 8):
        ops, h = handle(m)
        N = 20000 if m < 8 else 6000
        theta = torch.tensor(np.random.default_rng(2).uniform(0.1, 10.0, (N, 9)), device="cuda")
        q = torch.empty((N, 9), device="cuda", dtype=torch.float64)
        for solver, kern, thr in ((1, 0, 0), (2, 1, 0), (2, 2, 0), (2, 2, 64), (2, 2, 128), (2, 2, 256)):
            h.set_int("fom_solver", solver); h.set_int("frontal_kernel", kern); h.set_int("frontal_threads", thr)
            try:
                ms = timeit(lambda: h.fom_affine_raw(theta.data_ptr(), N, 0, 1, 1e-12, 50000, qoi=q.data_ptr(), stream=st), reps=1)
                print(f"m={m} n={ops.n} solver={solver} kernel={h.get_int('frontal_kernel') if solver == 2 else 0} threads={h.get_int('frontal_threads')} "
                      f"ctas/sm={h.get_int('frontal_ctas_per_sm')}: {N/ms*1e3:.3e} solves/s", flush=True)
            except Exception as e:
                print(f"m={m} solver={solver} kernel={kern} threads={thr}: {str(e)[:100]}", flush=True)
        h.close()
if "big" in which:
    ops, h = handle(26)
    print("m=26 program:", {k: h.get_int(k) for k in ("frontal_slots", "frontal_cmax", "frontal_nnz_factor", "frontal_pair_updates")})
    N = 592
    theta = torch.tensor(np.random.default_rng(2).uniform(0.1, 10.0, (N, 9)), device="cuda")
    q = torch.empty((N, 9), device="cuda", dtype=torch.float64)
    for thr in (0, 128, 256, 384):
        h.set_int("fom_solver", 2); h.set_int("frontal_threads", thr)
        ms = timeit(lambda: h.fom_affine_raw(theta.data_ptr(), N, 0, 1, 1e-12, 50000, qoi=q.data_ptr(), stream=st), reps=1, warm=0)
        print(f"m=26 direct D2 qoi threads={h.get_int('frontal_threads')} ctas/sm={h.get_int('frontal_ctas_per_sm')}: {N/ms*1e3:.1f} solves/s ({ms:.0f} ms)", flush=True)
    h.set_int("frontal_threads", 0); h.set_int("frontal_mode", 1)
    ms = timeit(lambda: h.fom_affine_raw(theta.data_ptr(), N, 0, 1, 1e-12, 50000, qoi=q.data_ptr(), stream=st), reps=1, warm=0)
    print(f"m=26 direct D2 solve mode: {N/ms*1e3:.1f} solves/s ({ms:.0f} ms)", flush=True)
    h.set_int("fom_solver", 1)
    Np = 296
    ms = timeit(lambda: h.fom_affine_raw(theta.data_ptr(), Np, 0, 1, 1e-12, 50000, qoi=q.data_ptr(), stream=st), reps=1, warm=0)
    print(f"m=26 stream PCG K4 (N={Np}): {Np/ms*1e3:.1f} solves/s", flush=True)
    h.close()
