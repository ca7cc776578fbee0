"""POD reduced basis for the nine-parameter fin, following the (commented) recipe that produced the
reference's ``data/basis_nine_param.txt`` -- ``rom/generate_reduced_basis_nine_param.py:296-318``:
snapshots ``forward_nine_param(k)`` at ``k ~ U(0.1, 3.5)^9``, eigen-decomposition of ``Y Y^T`` and the
UNNORMALISED modes ``U_i = sum_s v[s, i] Y[s, :]``.  The snapshot solves run in the batched PCG kernel."""
from __future__ import annotations

import numpy as np

from .. import _cabi
from ..assembly import build_operators

__all__ = ["generate_pod_basis"]


def generate_pod_basis(V, n_snapshots=200, basis_size=81, seed=0, lo=0.1, hi=3.5, device=0, tol=1e-12):
    ops = build_operators(V)
    h = _cabi.TfinHandle(device)
    try:
        h.set_operator(ops.row_ptr, ops.col_idx, ops.affine_terms(), ops.rhs, True)
        rng = np.random.default_rng(seed)
        theta = rng.uniform(lo, hi, (n_snapshots, ops.affine_terms().shape[0] - 1))
        out = h.fom_affine(theta, _cabi.IN_PARAMS, tol=tol, want_w=True, want_qoi=False)
        if np.any(out["status"] != _cabi.STATUS_CONVERGED):
            raise RuntimeError("snapshot solve failed")
        Y = out["w"]
    finally:
        h.close()
    e, v = np.linalg.eigh(Y @ Y.T)            # the reference calls eig; eigh + descending sort is the same set
    order = np.argsort(e)[::-1][:basis_size]
    return np.ascontiguousarray((v[:, order].T @ Y).T)
