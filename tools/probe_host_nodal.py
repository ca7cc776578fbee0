"""GPU probe: Fin.forward_qoi from PAGEABLE numpy fields (the drop-in user path), pipelined vs single-shot staging."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianinferencedl_b200 import get_space, Fin
fin = Fin(get_space(40))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
k = np.exp(0.3 * np.random.default_rng(0).standard_normal((N, fin.dofs)))
for chunk in (0, 8192, 16384, 4096):
    fin.handle.set_int("host_chunk", chunk)
    fin.forward_qoi(k[:20000])
    t0 = time.perf_counter(); q = fin.forward_qoi(k); dt = time.perf_counter() - t0
    print(f"host_chunk={chunk}: {N/dt:.3e} solves/s from pageable numpy ({dt*1e3:.0f} ms)")
data = q[0]
NG = min(N, 50000)
for chunk in (0, 8192):
    fin.handle.set_int("host_chunk", chunk)
    fin.gradient(k[:10000], data)
    t0 = time.perf_counter(); g = fin.gradient(k[:NG], data); dt = time.perf_counter() - t0
    print(f"host_chunk={chunk}: {NG/dt:.3e} gradients/s from pageable numpy ({dt*1e3:.0f} ms)")
