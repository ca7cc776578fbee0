"""GPU probe: which PCG path / variant serves meshes between 2 k and 8 k dofs (affine and nodal operators)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bayesianinferencedl_b200 import get_space, AffineROMFin, Fin, _cabi
from bayesianinferencedl_b200.assembly import build_operators
for m in (4, 5, 6, 7):
    V = get_space(40, m=m); ops = build_operators(V)
    h = _cabi.TfinHandle(0)
    h.set_operator(ops.row_ptr, ops.col_idx, ops.vals, ops.rhs); h.set_observation(*ops.obs_csr())
    th = np.random.default_rng(0).uniform(0.1, 3.5, (4, 9))
    try:
        out = h.fom_affine(th); print(m, ops.n, "affine ok path", h.get_int("pcg_path"), "T", h.get_int("pcg_threads"), "R", h.get_int("pcg_rows_per_thread"), out["status"], out["iters"])
    except Exception as e: print(m, ops.n, "affine FAIL", e)
    try:
        fin = Fin(V); q = fin.forward_qoi(np.ones((2, ops.n))); print(m, "nodal ok", fin.handle.get_int("pcg_threads"), fin.handle.get_int("pcg_rows_per_thread"))
    except Exception as e: print(m, ops.n, "nodal FAIL", e)
