"""GPU probe: throughput of the batched nodal-conductivity LSPG (tfin_rom_nodal) on device-resident inputs."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesianinferencedl_b200 import get_space, Fin
from bayesianinferencedl_b200.rom.pod import generate_pod_basis

V = get_space(40)
phi = generate_pod_basis(V, n_snapshots=200, basis_size=81, seed=0)
fin = Fin(V)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
rng = np.random.default_rng(1)
fin.r_fwd_no_full_qoi(np.exp(0.3 * rng.standard_normal((4, fin.dofs))), phi)    # uploads the basis
k = torch.exp(0.3 * torch.randn((N, fin.dofs), device="cuda", dtype=torch.float64))
y = torch.empty((N, 10), device="cuda", dtype=torch.float64)
h = fin.handle; lib = h._lib
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); st = ts.cuda_stream
def run():
    rc = lib.tfin_rom_nodal(h._h, k.data_ptr(), N, 1, None, None, None, y.data_ptr(), None, st)
    assert rc == 0, lib.tfin_last_error()
for _ in range(2): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
fl = 2 * 105 * 36 * fin.dofs + 2 * fin.dofs * 84 * 7 + 81 ** 3 / 3
print(f"rom_nodal: N={N} {ms:.2f} ms  {N/ms*1e3:.3e} solves/s  ({N/ms*1e3*fl/1e12:.2f} TFLOP/s fp64)")
