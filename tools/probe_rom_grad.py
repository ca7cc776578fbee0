"""GPU probe: throughput of the batched reduced gradient (tfin_rom_gradient) on device-resident inputs."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesianinferencedl_b200 import get_space, AffineROMFin, _cabi
from bayesianinferencedl_b200.rom.pod import generate_pod_basis

V = get_space(40)
phi = generate_pod_basis(V, n_snapshots=200, basis_size=81, seed=0)
rom = AffineROMFin(V, None, phi)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
rng = np.random.default_rng(1)
theta = torch.tensor(rng.uniform(0.1, 3.5, (N, 9)), device="cuda")
data = torch.tensor(rng.uniform(0.05, 0.6, (1, 9)), device="cuda")
grad = torch.empty((N, 9), device="cuda", dtype=torch.float64)
cost = torch.empty(N, device="cuda", dtype=torch.float64)
rom.set_data(np.zeros(9)); rom.grad_reduced_nine_param(theta[:4].cpu().numpy())   # builds the Gram blocks
h = rom.handle
lib = h._lib
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); st = ts.cuda_stream   # events must sit on the launch stream
def run():
    rc = lib.tfin_rom_gradient(h._h, theta.data_ptr(), N, 0, 1, data.data_ptr(), 1, 0, grad.data_ptr(), cost.data_ptr(),
                               None, None, None, st)
    assert rc == 0, lib.tfin_last_error()
for _ in range(2): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); 
for _ in range(3): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"rom_gradient: N={N} {ms:.2f} ms  {N/ms*1e3:.3e} gradients/s  ({N/ms*1e3*(558657+2*81*81*90+4*81*81)/1e12:.2f} TFLOP/s fp64)")
q = torch.empty((N, 9), device="cuda", dtype=torch.float64)
def fwd():
    rc = lib.tfin_rom(h._h, theta.data_ptr(), N, 0, 1, None, q.data_ptr(), None, st); assert rc == 0
for _ in range(2): fwd()
e0.record()
for _ in range(3): fwd()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"rom forward : N={N} {ms:.2f} ms  {N/ms*1e3:.3e} solves/s")
