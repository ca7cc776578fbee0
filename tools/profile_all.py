"""One launch of every hot kernel at a modest size, for `ncu --set full` (tools/summarize_ncu.py reads the report).

    ncu --set full --clock-control none --import-source on -k regex:'pcg_|rom_|field_sample|field_normal|csr_project' \
        -o gpurun_out/prof_r1 python tools/profile_all.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesianinferencedl_b200 import get_space, AffineROMFin, Fin, _cabi
from bayesianinferencedl_b200.assembly import build_operators
from bayesianinferencedl_b200.bayesian_inference.gaussian_field import FieldSampler
from bayesianinferencedl_b200.rom.pod import generate_pod_basis

small = os.environ.get("PROFILE_SKIP_SMALL") != "1"
refined = os.environ.get("PROFILE_SKIP_STREAM") != "1"
rng = np.random.default_rng(0)
if small:
    V = get_space(40)
    phi = generate_pod_basis(V, 200, 81, seed=0)
    rom = AffineROMFin(V, None, phi)
    fin = Fin(V)
    prior = FieldSampler(V, "m52", 1.6, handle=fin.handle)
    NS = 2960                                        # 148 SMs x 2 CTAs x 10 samples
    k5 = rng.uniform(0.1, 1.0, (NS, 5))
    theta5 = np.concatenate([k5, k5[:, 3::-1]], axis=1)
    out = rom.handle.fom_affine(theta5)                                  # K1
    print("K1 affine PCG: mean iters", out["iters"].mean())
    k = prior.sample(N=NS, seed=3)                                       # F3 + F4
    out = fin.handle.fom_nodal(k)                                        # K2
    print("K2 nodal PCG: mean iters", out["iters"].mean())
    NR = 9472                                                            # one ROM chunk
    theta = rng.uniform(0.1, 3.5, (NR, 9))
    rom.handle.rom(theta)                                                # R1 + R2
    rom.set_data(np.full(9, 0.3))
    rom.grad_reduced_nine_param(theta)                                   # R1 + R2(adj) + R3
    fin.r_fwd_no_full_qoi(k[:1184], phi)                                 # R4 + R2
    fin.gradient(k[:592], np.full(9, 0.3))                               # K2 adjoint variant
if refined:
    Vr = get_space(40, m=26)
    ops = build_operators(Vr)
    h = _cabi.TfinHandle(0)
    h.set_operator(ops.row_ptr, ops.col_idx, ops.vals, ops.rhs)
    h.set_observation(*ops.obs_csr())
    N, maxit = int(os.environ.get("PS_N", 1184)), int(os.environ.get("PS_MAXIT", 60))
    out = h.fom_affine(rng.uniform(0.1, 10.0, (N, 9)), maxit=maxit)      # K4, fixed iteration count
    print("K4 streaming PCG: n", ops.n, "tile", h.get_int("stream_tile"), "iters total", int(out["iters"].sum()),
          "algorithmic bytes", 88.0 * ops.n * int(out["iters"].sum()))
print("profile_all done")
