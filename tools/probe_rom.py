"""ROM (nine-parameter LSPG, n_r = 81) throughput against the chunk size of the combine -> Cholesky pipeline."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesianinferencedl_b200 import get_space, AffineROMFin, _cabi

ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); st = ts.cuda_stream
from bayesianinferencedl_b200.rom.pod import generate_pod_basis
V = get_space(40)
phi = generate_pod_basis(V, 200, 81, seed=0, device=0)
model = AffineROMFin(V, None, phi, device=0)
h = model.handle
N = 1_000_000
theta = torch.tensor(np.random.default_rng(3).uniform(0.1, 3.5, (N, 9)), device="cuda")
q = torch.empty((N, 9), device="cuda", dtype=torch.float64)
stt = torch.empty(N, device="cuda", dtype=torch.int32)
for chunk in [int(a) for a in sys.argv[1:]] or [0, 9472 * 2, 9472 * 4, 9472 * 8]:
    h.set_int("rom_chunk", chunk)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h.rom_raw(theta.data_ptr(), N, _cabi.IN_PARAMS, _cabi.MEM_DEVICE, qoi=q.data_ptr(), status=stt.data_ptr(), stream=st)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"rom_chunk={chunk:7d}: {best:8.2f} ms  {N / best / 1e3:7.2f} M solves/s", flush=True)
