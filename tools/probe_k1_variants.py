"""GPU probe: on-chip affine PCG (K1) variants on the headline workload, and streaming-kernel tile widths."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesianinferencedl_b200 import get_space, AffineROMFin, _cabi
from bayesianinferencedl_b200.assembly import build_operators
V = get_space(40)
rom = AffineROMFin(V, None, np.random.default_rng(0).standard_normal((1597, 8)))
h = rom.handle
N = 50000
k5 = np.random.default_rng(1).uniform(0.1, 1.0, (N, 5))
theta = torch.tensor(np.concatenate([k5, k5[:, 3::-1]], axis=1), device="cuda")
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); st = ts.cuda_stream
q = torch.empty((N, 9), device="cuda", dtype=torch.float64)
def timeit(fn, reps=2):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for R, WR in ((0, -1), (5, 0), (6, 0), (8, 0), (3, 1), (4, 1)):
    h.set_int("pcg_rows_per_thread", R); h.set_int("pcg_reg_slots", WR)
    try:
        ms = timeit(lambda: h.fom_affine_raw(theta.data_ptr(), N, 0, 1, 1e-12, 20000, qoi=q.data_ptr(), stream=st))
        print(f"K1 R={R} WR={WR}: {N/ms*1e3:.3e} solves/s  (T={h.get_int('pcg_threads')} R={h.get_int('pcg_rows_per_thread')} "
              f"occ={h.get_int('pcg_ctas_per_sm')} regslots={h.get_int('pcg_reg_slots')})")
    except Exception as e:
        print(f"K1 R={R} WR={WR}: {e}")
del rom
Vr = get_space(40, m=26); ops = build_operators(Vr)
hr = _cabi.TfinHandle(0)
hr.set_operator(ops.row_ptr, ops.col_idx, ops.vals, ops.rhs); hr.set_observation(*ops.obs_csr())
for tile, Nr in ((8, 2368), (16, 2368), (16, 4736)):
    hr.set_int("stream_tile", tile)
    th = torch.tensor(np.random.default_rng(2).uniform(0.1, 10.0, (Nr, 9)), device="cuda")
    qo = torch.empty((Nr, 9), device="cuda", dtype=torch.float64); it = torch.empty(Nr, device="cuda", dtype=torch.int32)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); hr.fom_affine_raw(th.data_ptr(), Nr, 0, 1, 1e-12, 50000, qoi=qo.data_ptr(), iters=it.data_ptr(), stream=st); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1); gbs = 88.0 * ops.n * it.double().sum().item() / (ms * 1e-3) / 1e9
    print(f"K4 tile={tile} N={Nr}: {Nr/ms*1e3:.1f} solves/s, {gbs/6454.9:.3f} of the HBM peak")
