"""Scratch probe of the streaming PCG kernel on the refined mesh (m=26, n=99 945)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesianinferencedl_b200 import get_space, _cabi
from bayesianinferencedl_b200.assembly import build_operators
m = int(os.environ.get("PROBE_M", 26))
t0 = time.time(); V = get_space(40, m=m); ops = build_operators(V); print("n", ops.n, "nnz", ops.nnz, "build", time.time() - t0)
h = _cabi.TfinHandle(0)
t0 = time.time(); h.set_operator(ops.row_ptr, ops.col_idx, ops.vals, ops.rhs); h.set_observation(*ops.obs_csr()); print("upload", time.time() - t0)
rng = np.random.default_rng(2)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); st = ts.cuda_stream
print("bandwidth after RCM", h.get_int("stream_bandwidth"))
for tile, ring, N, maxit, pad in ((8, -1, 1184, 2, 0), (4, -1, 592, 200, 0), (8, -1, 1184, 200, 0), (8, 1, 1184, 200, 0), (16, -1, 2368, 200, 0), (8, -1, 1184, 20000, 0), (8, -1, 2368, 20000, 0)):
    theta = torch.tensor(rng.uniform(0.1, 10.0, (N, 9)), device="cuda")
    qoi = torch.empty((N, 9), device="cuda", dtype=torch.float64); it = torch.empty(N, device="cuda", dtype=torch.int32)
    stt = torch.empty(N, device="cuda", dtype=torch.int32)
    h.set_int("stream_tile", tile); h.set_int("stream_ring", ring); h.set_int("stream_pad_smem", pad); print("pad smem KB", pad); h.set_int("stream_prof", 1)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); h.fom_affine_raw(theta.data_ptr(), N, 0, 1, 1e-12, maxit, qoi=qoi.data_ptr(), iters=it.data_ptr(), status=stt.data_ptr(), stream=st); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1); iters = it.double().sum().item()
    gbs = 88.0 * ops.n * iters / (ms * 1e-3) / 1e9
    print(f"tile={tile} ring={h.get_int('stream_ring')} N={N} maxit={maxit}: {ms:.1f} ms, mean iters {iters/N:.1f}, {N/ms*1e3:.2f} solves/s, algorithmic {gbs:.0f} GB/s = {gbs/6454.9:.3f} of measured HBM peak; converged {(stt==0).sum().item()}")
    pr = [h.get_int(f"stream_prof_{i}") for i in range(4)]
    if pr[3] > 0: print("   CTA0 clocks/iter: pass A %.0f  pass B %.0f  (iters %d, ell width %d); steady-state %.0f GB/s at 1.92 GHz" % (pr[0]/pr[3], pr[1]/pr[3], pr[3], h.get_int("stream_ell_width"), 88.0*ops.n*tile*148/((pr[0]+pr[1])/pr[3]/1.92e9)/1e9))
