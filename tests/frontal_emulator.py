"""Test infrastructure: fetches the per-pivot program of the sparse-direct solver from libtfin.so (host-only entry
points tfin_frontal_*) and interprets it in numpy exactly the way the CUDA kernels of csrc/frontal.cuh do:

* ``run_solve``  -- D1 / D2 SOLVE mode: factorisation + forward elimination, backward substitution, observables;
* ``run_qoi``    -- D2 observables mode: the n_obs observation rows ride along as extra right-hand sides of the forward
                   elimination, qoi_o = (L^-1 B_obs[o])^T (L^-1 b); no factor storage, no backward substitution.

The symbolic phase is product code (C++); this interpreter lets the CPU suite pin it against the oracle without a GPU.
"""
import ctypes as C

import numpy as np

from bayesianinferencedl_b200 import _cabi

_ARRAYS = {"perm": np.int32, "piv_slot": np.uint16, "col_ptr": np.int32, "col_slot": np.uint16, "rhs": np.float64,
           "asm_ptr": np.int32, "asm_addr": np.uint32, "asm_eptr": np.int32, "ent_term": np.int32,
           "ent_coef": np.float64, "obs_ptr": np.int32, "obs_row": np.int32, "obs_val": np.float64}
_SCALARS = ("n", "nslots", "cmax", "nnzL", "pair_updates")


def fetch_program(ops, obs_csr=None, lookahead=True):
    """ops: assembly.FinOperators.  Returns a dict of numpy arrays / ints."""
    lib = _cabi.load_library()
    rp, ci = np.ascontiguousarray(ops.row_ptr, np.int32), np.ascontiguousarray(ops.col_idx, np.int32)
    vals, rhs = np.ascontiguousarray(ops.vals, np.float64), np.ascontiguousarray(ops.rhs, np.float64)
    optr, oidx, oval = ops.obs_csr() if obs_csr is None else obs_csr
    optr, oidx = np.ascontiguousarray(optr, np.int32), np.ascontiguousarray(oidx, np.int32)
    oval = np.ascontiguousarray(oval, np.float64)
    prog = C.c_void_p()
    rc = lib.tfin_frontal_analyze_ex(ops.n, ci.shape[0], rp.ctypes.data, ci.ctypes.data, vals.shape[0], vals.ctypes.data,
                                     rhs.ctypes.data, optr.shape[0] - 1, optr.ctypes.data, oidx.ctypes.data,
                                     oval.ctypes.data, 1 if lookahead else 0, C.byref(prog))
    if rc != 0:
        raise RuntimeError(lib.tfin_last_error().decode())
    out = {}
    try:
        for k in _SCALARS:
            out[k] = int(lib.tfin_frontal_array(prog, k.encode(), None, 0))
        for k, dt in _ARRAYS.items():
            nbytes = int(lib.tfin_frontal_array(prog, k.encode(), None, 0))
            a = np.empty(nbytes // np.dtype(dt).itemsize, dtype=dt)
            if nbytes:
                lib.tfin_frontal_array(prog, k.encode(), a.ctypes.data, nbytes)
            out[k] = a
    finally:
        lib.tfin_frontal_free(prog)
    return out


def _tri(s):
    return s * (s + 1) // 2


def _assemble(P, F, cvec, j):
    for pos in range(P["asm_ptr"][j], P["asm_ptr"][j + 1]):
        e0, e1 = P["asm_eptr"][pos], P["asm_eptr"][pos + 1]
        F[P["asm_addr"][pos]] += np.dot(P["ent_coef"][e0:e1], cvec[P["ent_term"][e0:e1]])


def _eliminate(P, F, j):
    """Pivot step shared by both modes: returns (slots, l, rinv) and applies the rank-1 update to the front."""
    cp0, cp1 = P["col_ptr"][j], P["col_ptr"][j + 1]
    slots = P["col_slot"][cp0:cp1].astype(np.int64)
    assert np.all(np.diff(slots) > 0), "column slots must ascend"
    p = int(P["piv_slot"][j])
    pd = _tri(p) + p
    dd = F[pd]
    F[pd] = 0.0
    assert dd > 0.0, f"non-positive pivot {dd} at step {j}"
    rinv = 1.0 / np.sqrt(dd)
    ad = np.where(slots > p, _tri(slots) + p, _tri(p) + slots)
    l = F[ad] * rinv
    F[ad] = 0.0
    return slots, l, rinv, p


def _update(F, slots, l):
    for a in range(len(slots)):
        F[_tri(slots[a]) + slots[:a + 1]] -= l[a] * l[:a + 1]


def run_solve(P, cvec, n_obs):
    """-> (w in the caller's dof order, qoi, y.y)"""
    n, ns = P["n"], P["nslots"]
    F = np.zeros(_tri(ns))
    yv = np.zeros(ns)
    L, rinvs, ys = [None] * n, np.zeros(n), np.zeros(n)
    _assemble(P, F, cvec, 0)
    for j in range(n):
        slots, l, rinv, p = _eliminate(P, F, j)
        yp = (yv[p] + P["rhs"][j]) * rinv
        yv[p] = 0.0
        yv[slots] -= l * yp
        if j + 1 < n:
            _assemble(P, F, cvec, j + 1)
        _update(F, slots, l)
        L[j], rinvs[j], ys[j] = (slots, l), rinv, yp
    assert not F.any() and not yv.any(), "front must be empty after the last pivot"
    w = np.zeros(n)
    qoi = np.zeros(n_obs)
    for j in range(n - 1, -1, -1):
        slots, l = L[j]
        wj = (ys[j] - np.dot(l, yv[slots])) * rinvs[j]
        yv[P["piv_slot"][j]] = wj
        w[P["perm"][j]] = wj
        for o in range(P["obs_ptr"][j], P["obs_ptr"][j + 1]):
            qoi[P["obs_row"][o]] += P["obs_val"][o] * wj
    return w, qoi, float(np.dot(ys, ys))


def run_qoi(P, cvec, n_obs):
    """Observables through extra right-hand sides (no backward substitution)."""
    n, ns, R = P["n"], P["nslots"], 1 + n_obs
    F = np.zeros(_tri(ns))
    yv = np.zeros((ns, R))
    acc = np.zeros(R)
    _assemble(P, F, cvec, 0)
    for j in range(n):
        slots, l, rinv, p = _eliminate(P, F, j)
        v = yv[p].copy()
        yv[p] = 0.0
        v[0] += P["rhs"][j]
        for o in range(P["obs_ptr"][j], P["obs_ptr"][j + 1]):
            v[1 + P["obs_row"][o]] += P["obs_val"][o]
        ypiv = v * rinv
        if j + 1 < n:
            _assemble(P, F, cvec, j + 1)
        acc += ypiv * ypiv[0]
        yv[slots] -= np.outer(l, ypiv)
        _update(F, slots, l)
    return acc[1:], float(acc[0])
