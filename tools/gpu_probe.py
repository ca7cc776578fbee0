"""Scratch GPU probe: times K1 for each rows/thread setting and K3, prints a few lines. Not a benchmark."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesianinferencedl_b200 import get_space, AffineROMFin, Fin, _cabi
from oracle.thermal_fin_oracle import FinOracle, pod_basis

V = get_space(40)
orc = FinOracle(V.mesh().coordinates(), V.mesh().cells())
phi = pod_basis(orc)
rom = AffineROMFin(V, None, phi)
h = rom.handle
rng = np.random.default_rng(1)
N = int(os.environ.get("PROBE_N", 20000))
theta = rng.uniform(0.1, 1.0, (N, 5)); theta = np.concatenate([theta, theta[:, 3::-1]], 1)
th = torch.tensor(theta, device="cuda"); qoi = torch.empty((N, 9), device="cuda", dtype=torch.float64)
it = torch.empty(N, device="cuda", dtype=torch.int32)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); st = ts.cuda_stream; assert st != 0
for R, WR in ((0, -1), (5, 0), (6, 0), (8, 0), (3, 1), (4, 1)):
    h.set_int("pcg_rows_per_thread", R); h.set_int("pcg_reg_slots", WR)
    try:
        for rep in range(2):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            h.fom_affine_raw(th.data_ptr(), N, 0, 1, 1e-12, 20000, qoi=qoi.data_ptr(), iters=it.data_ptr(), stream=st)
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"K1 R={R} WR={WR}: T={h.get_int('pcg_threads')} R_used={h.get_int('pcg_rows_per_thread')} occ={h.get_int('pcg_ctas_per_sm')} "
              f"smem={h.get_int('pcg_smem_bytes')} W={h.get_int('ell_width')} WT={h.get_int('pcg_ell_width_compiled')} regslots={h.get_int('pcg_reg_slots')}: {ms:.2f} ms, {N/ms*1e3:.0f} solves/s, mean iters {it.float().mean().item():.1f}")
    except Exception as ex:
        print("K1 R=", R, WR, "failed:", ex)
h.set_int("pcg_reg_slots", -1)
h.set_int("pcg_rows_per_thread", 0)
NR = int(os.environ.get("PROBE_NR", 200000))
thr = torch.tensor(rng.uniform(0.1, 3.5, (NR, 9)), device="cuda"); qr = torch.empty((NR, 9), device="cuda", dtype=torch.float64)
for chunk in (0, 2048, 4096, 16384):
    h.set_int("rom_chunk", chunk)
    for rep in range(2):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); h.rom_raw(thr.data_ptr(), NR, 0, 1, qoi=qr.data_ptr(), stream=st); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"K3 chunk={chunk}: {ms:.2f} ms, {NR/ms*1e3:.0f} solves/s")
fin = Fin(V)
NN = int(os.environ.get("PROBE_NN", 10000))
k = torch.tensor(np.exp(0.3 * rng.standard_normal((NN, fin.dofs))), device="cuda"); qn = torch.empty((NN, 9), device="cuda", dtype=torch.float64)
for rep in range(2):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); fin.handle.fom_nodal_raw(k.data_ptr(), NN, 1, 1e-12, 20000, qoi=qn.data_ptr(), stream=st); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"K2: {ms:.2f} ms, {NN/ms*1e3:.0f} solves/s  T={fin.handle.get_int('pcg_threads')} R={fin.handle.get_int('pcg_rows_per_thread')} W={fin.handle.get_int('ell_width_nodal')}")
