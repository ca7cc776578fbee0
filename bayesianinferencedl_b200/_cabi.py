"""ctypes binding of libtfin.so (include/tfin.h).  There is NO fallback: if the library is missing or no
sm_100 device is usable, every compute entry point raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtfin.so")

MEM_HOST, MEM_DEVICE = 0, 1
KERN_SQ_EXP, KERN_M52, KERN_M32 = 0, 1, 2
IN_PARAMS, IN_NODAL = 0, 1
STATUS_CONVERGED, STATUS_MAXIT, STATUS_BREAKDOWN = 0, 1, 2

_p_f64 = C.POINTER(C.c_double)
_p_i32 = C.POINTER(C.c_int32)
_handle = C.c_void_p

# name -> (restype, argtypes); must list every symbol declared in include/tfin.h
SIGNATURES = {
    "tfin_version": (C.c_int, []),
    "tfin_last_error": (C.c_char_p, []),
    "tfin_create": (C.c_int, [C.c_int, C.POINTER(_handle)]),
    "tfin_destroy": (C.c_int, [_handle]),
    "tfin_set_operator": (C.c_int, [_handle, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                    C.c_void_p, C.c_void_p, C.c_int32]),
    "tfin_set_observation": (C.c_int, [_handle, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tfin_set_averaging": (C.c_int, [_handle, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tfin_set_cells": (C.c_int, [_handle, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]),
    "tfin_set_rom": (C.c_int, [_handle, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tfin_fom_affine": (C.c_int, [_handle, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_double, C.c_int32,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tfin_fom_nodal": (C.c_int, [_handle, C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_int32,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tfin_rom": (C.c_int, [_handle, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                           C.c_void_p, C.c_void_p]),
    "tfin_set_basis": (C.c_int, [_handle, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "tfin_rom_nodal": (C.c_int, [_handle, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p]),
    "tfin_set_rom_gradient": (C.c_int, [_handle, C.c_int32, C.c_int32, C.c_void_p]),
    "tfin_rom_gradient": (C.c_int, [_handle, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                    C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p]),
    "tfin_fom_nodal_gradient": (C.c_int, [_handle, C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_int32,
                                          C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p]),
    "tfin_fom_nodal_sensitivity": (C.c_int, [_handle, C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_int32,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tfin_field_set_cov": (C.c_int, [_handle, C.c_int32, C.c_void_p, C.c_int32, C.c_double, C.c_void_p]),
    "tfin_field_set_chol": (C.c_int, [_handle, C.c_int32, C.c_void_p]),
    "tfin_field_sample": (C.c_int, [_handle, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                    C.c_void_p]),
    "tfin_pcn_chains": (C.c_int, [_handle, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_double, C.c_void_p,
                                  C.c_double, C.c_uint64, C.c_double, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tfin_subfin_avg": (C.c_int, [_handle, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "tfin_frontal_analyze": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                       C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "tfin_frontal_analyze_ex": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                          C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]),
    "tfin_frontal_array": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]),
    "tfin_frontal_free": (None, [C.c_void_p]),
    "tfin_smem_bandwidth": (C.c_int, [_handle, C.POINTER(C.c_double)]),
    "tfin_kernel_launches": (C.c_int64, [_handle]),
    "tfin_get_int": (C.c_int64, [_handle, C.c_char_p]),
    "tfin_set_int": (C.c_int, [_handle, C.c_char_p, C.c_int64]),
}

_lib = None


def load_library(path=None):
    """dlopen libtfin.so and attach prototypes.  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} not found: build it with `python -m bayesianinferencedl_b200._build` "
            "(or __graft_entry__.build()). There is no CPU fallback.")
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


class TfinError(RuntimeError):
    pass


def _check(lib, rc, what):
    if rc != 0:
        msg = lib.tfin_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise TfinError(f"{what} failed ({rc}): {msg}")


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class TfinHandle:
    """Owns one ``tfin_handle_t`` (one CUDA device).  numpy-level wrappers of the C ABI."""

    def __init__(self, device=0):
        self._lib = load_library()
        self._h = _handle()
        _check(self._lib, self._lib.tfin_create(int(device), C.byref(self._h)), "tfin_create")
        self.device = int(device)
        self.n = self.n_obs = self.n_terms = self.n_r = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.tfin_destroy(self._h)
            self._h = _handle()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- setup
    def set_operator(self, row_ptr, col_idx, vals, rhs, prune_zeros=True):
        row_ptr, col_idx, vals, rhs = _i32(row_ptr), _i32(col_idx), _f64(vals), _f64(rhs)
        n, nnz = rhs.shape[0], col_idx.shape[0]
        if vals.ndim != 2 or vals.shape[1] != nnz or row_ptr.shape[0] != n + 1:
            raise ValueError("set_operator: inconsistent shapes")
        _check(self._lib, self._lib.tfin_set_operator(self._h, n, nnz, _ptr(row_ptr), _ptr(col_idx), vals.shape[0],
                                                      _ptr(vals), _ptr(rhs), int(bool(prune_zeros))),
               "tfin_set_operator")
        self.n, self.n_terms = n, vals.shape[0]

    def set_observation(self, ptr, idx, val):
        ptr, idx, val = _i32(ptr), _i32(idx), _f64(val)
        _check(self._lib, self._lib.tfin_set_observation(self._h, ptr.shape[0] - 1, _ptr(ptr), _ptr(idx), _ptr(val)),
               "tfin_set_observation")
        self.n_obs = ptr.shape[0] - 1

    def set_averaging(self, ptr, idx, val):
        ptr, idx, val = _i32(ptr), _i32(idx), _f64(val)
        _check(self._lib, self._lib.tfin_set_averaging(self._h, ptr.shape[0] - 1, _ptr(ptr), _ptr(idx), _ptr(val)),
               "tfin_set_averaging")

    def set_cells(self, cells, Ke, prune_zeros=True):
        cells, Ke = _i32(cells), _f64(Ke)
        if cells.ndim != 2 or cells.shape[1] != 3 or Ke.shape != (cells.shape[0], 3, 3):
            raise ValueError("set_cells: cells must be (nc,3) and Ke (nc,3,3)")
        _check(self._lib, self._lib.tfin_set_cells(self._h, cells.shape[0], _ptr(cells), _ptr(Ke),
                                                   int(bool(prune_zeros))), "tfin_set_cells")

    def set_rom(self, S, G, obs_phi):
        S, G, obs_phi = _f64(S), _f64(G), _f64(obs_phi)
        n_terms, n_r = G.shape
        n_obs = obs_phi.shape[0]
        if S.shape != (n_terms * (n_terms + 1) // 2, n_r * (n_r + 1) // 2) or obs_phi.shape[1] != n_r:
            raise ValueError("set_rom: inconsistent shapes")
        _check(self._lib, self._lib.tfin_set_rom(self._h, n_r, n_terms, n_obs, _ptr(S), _ptr(G), _ptr(obs_phi)),
               "tfin_set_rom")
        self.n_r, self.rom_terms, self.rom_obs = n_r, n_terms, n_obs

    def set_basis(self, phi, out_phi):
        """phi (n, n_r); out_phi (n_out, n_r): rows projected onto the reduced solution by rom_nodal."""
        phi, out_phi = _f64(phi), _f64(out_phi)
        if phi.ndim != 2 or phi.shape[0] != self.n or out_phi.ndim != 2 or out_phi.shape[1] != phi.shape[1]:
            raise ValueError("set_basis: phi must be (n, n_r) and out_phi (n_out, n_r)")
        _check(self._lib, self._lib.tfin_set_basis(self._h, self.n, phi.shape[1], _ptr(phi), out_phi.shape[0],
                                                   _ptr(out_phi)), "tfin_set_basis")
        self.basis_n_r, self.basis_n_out = phi.shape[1], out_phi.shape[0]

    def set_rom_gradient(self, gram):
        """gram[t][q-1] = Psi_t^T Psi_q, shape (n_terms, n_terms-1, n_r, n_r)."""
        gram = _f64(gram)
        if gram.shape != (self.rom_terms, self.rom_terms - 1, self.n_r, self.n_r):
            raise ValueError("set_rom_gradient: gram must be (n_terms, n_terms-1, n_r, n_r)")
        _check(self._lib, self._lib.tfin_set_rom_gradient(self._h, self.n_r, self.rom_terms, _ptr(gram)),
               "tfin_set_rom_gradient")

    # ---- host-memory batch calls (numpy in, numpy out)
    def fom_affine(self, batch, in_kind=IN_PARAMS, tol=1e-12, maxit=20000, want_w=False, want_qoi=True,
                   want_stats=True):
        batch = _f64(batch)
        cols = self.n if in_kind == IN_NODAL else self.n_terms - 1
        if batch.ndim != 2 or batch.shape[1] != cols:
            raise ValueError(f"fom_affine: expected (N, {cols}) input, got {batch.shape}")
        return self._fom(False, batch, in_kind, tol, maxit, want_w, want_qoi, want_stats)

    def fom_nodal(self, k, tol=1e-12, maxit=20000, want_w=False, want_qoi=True, want_stats=True):
        k = _f64(k)
        if k.ndim != 2 or k.shape[1] != self.n:
            raise ValueError(f"fom_nodal: expected (N, {self.n}) input, got {k.shape}")
        return self._fom(True, k, IN_NODAL, tol, maxit, want_w, want_qoi, want_stats)

    def _fom(self, nodal, batch, in_kind, tol, maxit, want_w, want_qoi, want_stats):
        N = batch.shape[0]
        w = np.empty((N, self.n)) if want_w else None
        qoi = np.empty((N, self.n_obs)) if want_qoi else None
        iters = np.empty(N, dtype=np.int32) if want_stats else None
        status = np.empty(N, dtype=np.int32) if want_stats else None
        relres = np.empty(N) if want_stats else None
        if nodal:
            rc = self._lib.tfin_fom_nodal(self._h, _ptr(batch), N, MEM_HOST, float(tol), int(maxit), _ptr(w),
                                          _ptr(qoi), _ptr(iters), _ptr(status), _ptr(relres), None)
            _check(self._lib, rc, "tfin_fom_nodal")
        else:
            rc = self._lib.tfin_fom_affine(self._h, _ptr(batch), N, int(in_kind), MEM_HOST, float(tol), int(maxit),
                                           _ptr(w), _ptr(qoi), _ptr(iters), _ptr(status), _ptr(relres), None)
            _check(self._lib, rc, "tfin_fom_affine")
        return {"w": w, "qoi": qoi, "iters": iters, "status": status, "relres": relres}

    def fom_nodal_gradient(self, k, data, tol=1e-12, maxit=20000):
        """Batched Fin.gradient: k (N, n), data (n_obs,) or (N, n_obs) -> grad (N, n), cost (N), qoi (N, n_obs)."""
        k, data = _f64(k), _f64(data)
        if k.ndim != 2 or k.shape[1] != self.n:
            raise ValueError(f"fom_nodal_gradient: expected (N, {self.n}) input, got {k.shape}")
        N = k.shape[0]
        if data.ndim == 1:
            data = data[None, :]
        if data.shape[1] != self.n_obs or data.shape[0] not in (1, N):
            raise ValueError(f"fom_nodal_gradient: data must be ({self.n_obs},) or (N, {self.n_obs}), got {data.shape}")
        grad, cost, qoi = np.empty((N, self.n)), np.empty(N), np.empty((N, self.n_obs))
        iters, status = np.empty(N, dtype=np.int32), np.empty(N, dtype=np.int32)
        rc = self._lib.tfin_fom_nodal_gradient(self._h, _ptr(k), N, MEM_HOST, float(tol), int(maxit), _ptr(data),
                                               data.shape[0], _ptr(grad), _ptr(cost), _ptr(qoi), _ptr(iters),
                                               _ptr(status), None)
        _check(self._lib, rc, "tfin_fom_nodal_gradient")
        return {"grad": grad, "cost": cost, "qoi": qoi, "iters": iters, "status": status, "relres": None}

    def fom_nodal_sensitivity(self, k, tol=1e-12, maxit=20000):
        """Batched Fin.sensitivity: k (N, n) -> jac (N, n_obs, n), qoi (N, n_obs)."""
        k = _f64(k)
        if k.ndim != 2 or k.shape[1] != self.n:
            raise ValueError(f"fom_nodal_sensitivity: expected (N, {self.n}) input, got {k.shape}")
        N = k.shape[0]
        jac, qoi = np.empty((N, self.n_obs, self.n)), np.empty((N, self.n_obs))
        iters, status = np.empty(N, dtype=np.int32), np.empty(N, dtype=np.int32)
        rc = self._lib.tfin_fom_nodal_sensitivity(self._h, _ptr(k), N, MEM_HOST, float(tol), int(maxit), _ptr(jac),
                                                  _ptr(qoi), _ptr(iters), _ptr(status), None)
        _check(self._lib, rc, "tfin_fom_nodal_sensitivity")
        return {"jac": jac, "qoi": qoi, "iters": iters, "status": status, "relres": None}

    def rom(self, batch, in_kind=IN_PARAMS, want_wr=True, want_qoi=True):
        batch = _f64(batch)
        cols = self.n if in_kind == IN_NODAL else self.rom_terms - 1
        if batch.ndim != 2 or batch.shape[1] != cols:
            raise ValueError(f"rom: expected (N, {cols}) input, got {batch.shape}")
        N = batch.shape[0]
        wr = np.empty((N, self.n_r)) if want_wr else None
        qoi = np.empty((N, self.rom_obs)) if want_qoi else None
        status = np.empty(N, dtype=np.int32)
        rc = self._lib.tfin_rom(self._h, _ptr(batch), N, int(in_kind), MEM_HOST, _ptr(wr), _ptr(qoi), _ptr(status),
                                None)
        _check(self._lib, rc, "tfin_rom")
        return {"w_r": wr, "qoi": qoi, "status": status}

    def rom_nodal(self, k, want_system=True):
        """Batched Fin.r_fwd_no_full: k (N, n) -> A_r (N, n_r, n_r), B_r, x_r (N, n_r), y (N, n_out)."""
        k = _f64(k)
        if k.ndim != 2 or k.shape[1] != self.n:
            raise ValueError(f"rom_nodal: expected (N, {self.n}) input, got {k.shape}")
        N, nr = k.shape[0], self.basis_n_r
        Ar = np.empty((N, nr, nr)) if want_system else None
        Br = np.empty((N, nr)) if want_system else None
        xr, y = np.empty((N, nr)), np.empty((N, self.basis_n_out))
        status = np.empty(N, dtype=np.int32)
        rc = self._lib.tfin_rom_nodal(self._h, _ptr(k), N, MEM_HOST, _ptr(Ar), _ptr(Br), _ptr(xr), _ptr(y),
                                      _ptr(status), None)
        _check(self._lib, rc, "tfin_rom_nodal")
        return {"A_r": Ar, "B_r": Br, "x_r": xr, "y": y, "status": status}

    def rom_gradient(self, batch, data, in_kind=IN_PARAMS, grad_kind=IN_PARAMS, want_wr=False):
        """Batched AffineROMFin.grad_reduced: -> grad (N, n_terms-1 | n), cost (N), qoi (N, n_obs)."""
        batch, data = _f64(batch), _f64(data)
        cols = self.n if in_kind == IN_NODAL else self.rom_terms - 1
        if batch.ndim != 2 or batch.shape[1] != cols:
            raise ValueError(f"rom_gradient: expected (N, {cols}) input, got {batch.shape}")
        N = batch.shape[0]
        if data.ndim == 1:
            data = data[None, :]
        if data.shape[1] != self.rom_obs or data.shape[0] not in (1, N):
            raise ValueError(f"rom_gradient: data must be ({self.rom_obs},) or (N, {self.rom_obs}), got {data.shape}")
        gcols = self.n if grad_kind == IN_NODAL else self.rom_terms - 1
        grad, cost, qoi = np.empty((N, gcols)), np.empty(N), np.empty((N, self.rom_obs))
        wr = np.empty((N, self.n_r)) if want_wr else None
        status = np.empty(N, dtype=np.int32)
        rc = self._lib.tfin_rom_gradient(self._h, _ptr(batch), N, int(in_kind), MEM_HOST, _ptr(data), data.shape[0],
                                         int(grad_kind), _ptr(grad), _ptr(cost), _ptr(qoi), _ptr(wr), _ptr(status),
                                         None)
        _check(self._lib, rc, "tfin_rom_gradient")
        return {"grad": grad, "cost": cost, "qoi": qoi, "w_r": wr, "status": status}

    # ---- Gaussian-field sampler
    def field_set_cov(self, coords, kern_type=KERN_M52, length=1.6, want_chol=True):
        """Covariance + Cholesky on the device; returns the upper factor (n, n) like scipy.linalg.cholesky."""
        coords = _f64(coords)
        if coords.ndim != 2 or coords.shape[1] != 2:
            raise ValueError("field_set_cov: coords must be (n, 2)")
        n = coords.shape[0]
        chol = np.empty((n, n)) if want_chol else None
        _check(self._lib, self._lib.tfin_field_set_cov(self._h, n, _ptr(coords), int(kern_type), float(length),
                                                       _ptr(chol)), "tfin_field_set_cov")
        self.field_n = n
        return chol

    def field_set_chol(self, chol):
        chol = _f64(chol)
        if chol.ndim != 2 or chol.shape[0] != chol.shape[1]:
            raise ValueError("field_set_chol: chol must be square")
        _check(self._lib, self._lib.tfin_field_set_chol(self._h, chol.shape[0], _ptr(chol)), "tfin_field_set_chol")
        self.field_n = chol.shape[0]

    def field_sample(self, N=None, z=None, seed=0, subsequence=0, first_row=0, want_z=False):
        """k = exp(0.5 chol^T z): from given normals z (N, n) or from the device generator (N, seed)."""
        n = self.field_n
        if z is not None:
            z = _f64(z)
            if z.ndim != 2 or z.shape[1] != n:
                raise ValueError(f"field_sample: z must be (N, {n})")
            N = z.shape[0]
        elif N is None:
            raise ValueError("field_sample: give N or z")
        k = np.empty((int(N), n))
        z_out = np.empty((int(N), n)) if want_z else None
        _check(self._lib, self._lib.tfin_field_sample(self._h, _ptr(z), int(seed), int(subsequence), int(first_row), int(N),
                                                      MEM_HOST, _ptr(k),
                                                      _ptr(z_out), None), "tfin_field_sample")
        return (k, z_out) if want_z else k

    def subfin_avg(self, k):
        k = _f64(k)
        N = k.shape[0]
        rows = self.n_terms - 1
        out = np.empty((N, rows))
        _check(self._lib, self._lib.tfin_subfin_avg(self._h, _ptr(k), N, MEM_HOST, _ptr(out), None),
               "tfin_subfin_avg")
        return out

    # ---- raw-pointer calls (device or pinned-host buffers owned by the caller, e.g. torch tensors)
    def fom_affine_raw(self, in_ptr, N, in_kind, mem, tol, maxit, w=0, qoi=0, iters=0, status=0, relres=0,
                       stream=0):
        rc = self._lib.tfin_fom_affine(self._h, in_ptr, N, in_kind, mem, tol, maxit, w or None, qoi or None,
                                       iters or None, status or None, relres or None, stream or None)
        _check(self._lib, rc, "tfin_fom_affine")

    def fom_nodal_raw(self, in_ptr, N, mem, tol, maxit, w=0, qoi=0, iters=0, status=0, relres=0, stream=0):
        rc = self._lib.tfin_fom_nodal(self._h, in_ptr, N, mem, tol, maxit, w or None, qoi or None, iters or None,
                                      status or None, relres or None, stream or None)
        _check(self._lib, rc, "tfin_fom_nodal")

    def rom_raw(self, in_ptr, N, in_kind, mem, wr=0, qoi=0, status=0, stream=0):
        rc = self._lib.tfin_rom(self._h, in_ptr, N, in_kind, mem, wr or None, qoi or None, status or None,
                                stream or None)
        _check(self._lib, rc, "tfin_rom")

    def smem_bandwidth(self):
        """Measured aggregate shared-memory read bandwidth of the device (GB/s): roofline denominator of the on-chip kernels."""
        out = C.c_double()
        _check(self._lib, self._lib.tfin_smem_bandwidth(self._h, C.byref(out)), "tfin_smem_bandwidth")
        return float(out.value)

    # ---- introspection / tuning
    def kernel_launches(self):
        return int(self._lib.tfin_kernel_launches(self._h))

    def get_int(self, key):
        return int(self._lib.tfin_get_int(self._h, key.encode()))

    def set_int(self, key, value):
        _check(self._lib, self._lib.tfin_set_int(self._h, key.encode(), int(value)), "tfin_set_int")
