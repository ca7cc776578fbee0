"""GPU probe: end-to-end rate of the device-resident data-set generator (fields, FOM, ROM, D2H of everything)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianinferencedl_b200 import get_space
from bayesianinferencedl_b200.deep_learning.generate_fin_dataset import DatasetGenerator
from bayesianinferencedl_b200.rom.pod import generate_pod_basis
V = get_space(40)
gen = DatasetGenerator(V, generate_pod_basis(V, 200, 81, seed=0))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
gen.generate(20000, seed=1)
t0 = time.perf_counter(); z_s, errs, qois = gen.generate(N, seed=2); dt = time.perf_counter() - t0
print(f"dataset: N={N} in {dt:.3f} s = {N/dt:.3e} samples/s (wall clock, includes D2H of {z_s.nbytes/1e9:.2f} GB of fields); "
      f"max |rom error| {np.abs(errs).max():.3e}")
