"""``AffineROMFin`` -- nine-parameter affine thermal fin and its LSPG reduced-order model, B200-native.

Mirrors the forward-map part of the reference's ``rom/averaged_affine_ROM.py`` (class ``AffineROMFin`` :52,
``forward`` :237, ``forward_reduced`` :260, ``forward_nine_param_reduced`` :278, ``qoi`` :312, ``qoi_reduced``
:323, ``subfin_avg_op`` :404, ``observation_operator`` :420, ``set_data`` :398, ``set_dl_model`` :401) plus the
scalar-parameter entry point ``forward_nine_param`` of ``rom/generate_reduced_basis_nine_param.py:178-183``.

A leading batch axis is accepted wherever the reference takes one sample.  The affine FOM runs in the batched
PCG kernel, the ROM in the Gram-tensor GEMM + warp-per-sample Cholesky kernels of libtfin.so.
"""
from __future__ import annotations

import time

import numpy as np

from .. import _cabi
from ..assembly import BIOT, build_operators, five_to_nine
from ..fom.forward_solve import DEFAULT_MAXIT, DEFAULT_TOL, _as_batch
from ..fom.thermal_fin import FinSpace, Function

__all__ = ["AffineROMFin", "rom_offline_tensors", "rom_gradient_tensors"]


def rom_offline_tensors(ops, phi, B_obs):
    """Offline tensors of the LSPG ROM (averaged_affine_ROM.py:292-304 rewritten for affine A):
    Psi_t = V_t phi (V_0 = Bi M, V_q = K_q);  S_pq = Psi_p^T Psi_q (+ transpose if p<q), packed lower
    row-major;  G_t = Psi_t^T b;  obs_phi = B_obs phi (averaged_affine_ROM.py:212)."""
    phi = np.ascontiguousarray(phi, dtype=np.float64)
    n_terms = ops.affine_terms().shape[0]
    n_r = phi.shape[1]
    Psi = [ops.csr(ops.affine_terms()[t]) @ phi for t in range(n_terms)]
    il = np.tril_indices(n_r)
    S = np.empty((n_terms * (n_terms + 1) // 2, len(il[0])))
    pq = 0
    for p in range(n_terms):
        for q in range(p, n_terms):
            M = Psi[p].T @ Psi[q]
            if q != p:
                M = M + M.T
            S[pq] = M[il]
            pq += 1
    G = np.stack([P.T @ ops.rhs for P in Psi])
    return S, G, B_obs @ phi


def rom_gradient_tensors(ops, phi):
    """Offline Gram blocks of the reduced gradient: gram[t][q-1] = Psi_t^T Psi_q with Psi_t = V_t phi, so that
    psi^T dA_dsigmak_phi[q-1] = sum_t theta_t gram[t][q-1] (averaged_affine_ROM.py:215-220, 343-348)."""
    phi = np.ascontiguousarray(phi, dtype=np.float64)
    n_terms = ops.affine_terms().shape[0]
    Psi = [ops.csr(ops.affine_terms()[t]) @ phi for t in range(n_terms)]
    return np.stack([np.stack([Psi[t].T @ Psi[q] for q in range(1, n_terms)]) for t in range(n_terms)])


class AffineROMFin:
    """Affine FOM ``A(k_s) = sum_q k_s[q] K_q + Bi M`` (averaged_affine_ROM.py:156-162) and its LSPG ROM."""

    def __init__(self, V: FinSpace, err_model, phi, external_obs=False, *, device=0, tol=None,
                 maxit=DEFAULT_MAXIT, prune_zeros=True, precision="fp64"):
        self.fwd_time = 0.0                       # averaged_affine_ROM.py:64-67 (kept for API parity)
        self.rom_grad_time = 0.0
        self.romml_grad_time = 0.0
        self.romml_grad_time_dl = 0.0
        self.num_params = 9
        if precision not in ("fp64", "fp32"):
            raise ValueError("precision must be 'fp64' or 'fp32'")
        # opt-in fp32 full-order path (csrc/pcg_f32.cuh; measured error floor 3.6e-5 on the observables)
        self.precision = precision
        self.tol = float(tol) if tol is not None else (DEFAULT_TOL if precision == "fp64" else 1e-8)
        self.maxit = int(maxit)

        self.phi = np.ascontiguousarray(phi, dtype=np.float64)
        if self.phi.ndim != 2:
            raise ValueError("phi must be (n, n_r)")
        (self.n, self.n_r) = self.phi.shape
        self.V = V
        self.ops = build_operators(V)
        self.dofs = self.ops.n
        if self.n != self.dofs:
            raise ValueError(f"basis has {self.n} rows but the space has {self.dofs} dofs")
        self.Bi = BIOT
        self.B = self.ops.rhs.copy()
        (self.fin1_A, self.fin2_A, self.fin3_A, self.fin4_A, self.fin5_A,
         self.fin6_A, self.fin7_A, self.fin8_A, self.fin9_A) = self.ops.subfin_area

        if external_obs is not False and external_obs is not None:
            from ..fom.forward_solve import Fin
            self.n_obs = 40
            b_vals = Fin._external_indices(self, external_obs)
            self.B_obs = np.zeros((self.n_obs, self.dofs))
            self.B_obs[np.arange(self.n_obs), b_vals] = 1
        else:
            self.n_obs = 9
            self.B_obs = self.observation_operator()
        self.dsigma_dk = self.observation_operator()                   # :210
        self.B_obs_phi = np.dot(self.B_obs, self.phi)                   # :212
        self.dl_model = err_model
        self.data = None

        S, G, obs_phi = rom_offline_tensors(self.ops, self.phi, self.B_obs)
        self._h = _cabi.TfinHandle(device)
        self._h.set_operator(self.ops.row_ptr, self.ops.col_idx, self.ops.affine_terms(), self.ops.rhs, prune_zeros)
        self._h.set_observation(*self.ops.obs_csr(self.B_obs))
        self._h.set_averaging(*self.ops.obs_csr(self.ops.B_obs))
        self._h.set_rom(S, G, obs_phi)
        if precision == "fp32":
            self._h.set_int("pcg_precision", 32)
        self._grad_ready = False
        self._dA_phi = None

    @property
    def handle(self):
        return self._h

    @property
    def dA_dsigmak_phi(self):
        """:215-220: ``dA_dsigmak[q] @ phi`` = K_q phi, shape (9, n, n_r) (built on first use; the batched gradient
        kernels use the Gram blocks of :func:`rom_gradient_tensors` instead)."""
        if self._dA_phi is None:   # depends on phi: kept on this object (the operators are shared per space)
            self._dA_phi = np.stack([self.ops.csr(self.ops.affine_terms()[q]) @ self.phi for q in range(1, self.num_params + 1)])
        return self._dA_phi

    # ------------------------------------------------------------------ full-order affine model
    def forward(self, k):
        """:237-258.  ``k`` nodal field (n,) | (N, n): averaged over the sub-fins, then the affine solve."""
        kb, single = _as_batch(k, self.dofs, "AffineROMFin.forward")
        out = self._h.fom_affine(kb, _cabi.IN_NODAL, tol=self.tol, maxit=self.maxit, want_w=True, want_qoi=False)
        self._check(out)
        return Function(out["w"][0] if single else out["w"])

    def forward_nine_param(self, k_s):
        """generate_reduced_basis_nine_param.py:178-183: (9,) | (N, 9) sub-fin conductivities -> w."""
        tb, single = _as_batch(k_s, self.num_params, "AffineROMFin.forward_nine_param")
        out = self._h.fom_affine(tb, _cabi.IN_PARAMS, tol=self.tol, maxit=self.maxit, want_w=True, want_qoi=False)
        self._check(out)
        return Function(out["w"][0] if single else out["w"])

    def forward_five_param(self, k_s):
        return self.forward_nine_param(five_to_nine(k_s))

    def forward_nine_param_qoi(self, k_s, return_stats=False):
        """Fused ``qoi(forward_nine_param(k_s))`` without materialising w."""
        tb, single = _as_batch(k_s, self.num_params, "AffineROMFin.forward_nine_param_qoi")
        out = self._h.fom_affine(tb, _cabi.IN_PARAMS, tol=self.tol, maxit=self.maxit, want_w=False, want_qoi=True)
        self._check(out)
        q = out["qoi"][0] if single else out["qoi"]
        return (q, out) if return_stats else q

    def _check(self, out):
        st = out["status"]
        if st is not None and np.any(st != _cabi.STATUS_CONVERGED):
            bad = np.nonzero(st != _cabi.STATUS_CONVERGED)[0]
            raise RuntimeError(f"solve failed for {len(bad)} sample(s) (first: {bad[0]}, status {int(st[bad[0]])}); "
                               "conductivities must be positive")
        rr = out.get("relres")   # true residual (PCG) / consistency of the two substitutions (direct solver)
        cap = max(1e3 * self.tol, 1e-9 if self.precision == "fp64" else 1e-3)   # the fp32 path floors at ~1e-5
        if rr is not None and len(rr) and not np.all(rr <= cap):
            bad = np.nonzero(~(rr <= cap))[0]
            raise RuntimeError(f"true residual {rr[bad[0]]:.3e} of sample {bad[0]} is far above tol = {self.tol:g}")

    # ------------------------------------------------------------------ reduced-order model
    def forward_reduced(self, k):
        """:260-276.  ``k`` nodal field (n,) | (N, n) -> w_r (n_r,) | (N, n_r)."""
        t_i = time.time()
        kb, single = _as_batch(k, self.dofs, "AffineROMFin.forward_reduced")
        out = self._h.rom(kb, _cabi.IN_NODAL, want_wr=True, want_qoi=False)
        self._check(out)
        self.fwd_time += time.time() - t_i
        return out["w_r"][0] if single else out["w_r"]

    def forward_nine_param_reduced(self, k_s):
        """:278-310.  (9,) | (N, 9) -> w_r."""
        t_i = time.time()
        tb, single = _as_batch(k_s, self.num_params, "AffineROMFin.forward_nine_param_reduced")
        out = self._h.rom(tb, _cabi.IN_PARAMS, want_wr=True, want_qoi=False)
        self._check(out)
        self.fwd_time += time.time() - t_i
        return out["w_r"][0] if single else out["w_r"]

    def forward_reduced_qoi(self, k_or_ks):
        """Fused ``qoi_reduced(forward_[nine_param_]reduced(.))``; the input kind is inferred from the width."""
        a = np.asarray(k_or_ks, dtype=np.float64)
        kind = _cabi.IN_PARAMS if a.shape[-1] == self.num_params and self.num_params != self.dofs else _cabi.IN_NODAL
        ab, single = _as_batch(a, self.num_params if kind == _cabi.IN_PARAMS else self.dofs, "forward_reduced_qoi")
        out = self._h.rom(ab, kind, want_wr=False, want_qoi=True)
        self._check(out)
        return out["qoi"][0] if single else out["qoi"]

    # ------------------------------------------------------------------ observation
    def qoi(self, w):
        """:312-321."""
        w = np.asarray(w, dtype=np.float64)
        return w @ self.B_obs.T if w.ndim == 2 else np.dot(self.B_obs, w)

    def qoi_reduced(self, w_r):
        """:323-333."""
        t_i = time.time()
        w_r = np.asarray(w_r, dtype=np.float64)
        q = w_r @ self.B_obs_phi.T if w_r.ndim == 2 else np.dot(self.B_obs_phi, w_r)
        self.fwd_time += time.time() - t_i
        return q

    def subfin_avg_op(self, k):
        """:404-418, batched on the GPU."""
        kb, single = _as_batch(k, self.dofs, "AffineROMFin.subfin_avg_op")
        out = self._h.subfin_avg(kb)
        return out[0] if single else out

    def observation_operator(self):
        """:420-445."""
        return self.ops.B_obs.copy()

    def set_data(self, data):
        self.data = data

    def set_dl_model(self, model):
        self.dl_model = model

    # ------------------------------------------------------------------ 'next' rows
    def _rom_gradient(self, batch, in_kind, grad_kind, data):
        data = self.data if data is None else data
        if data is None:
            raise ValueError("grad_reduced: call set_data(data) first (averaged_affine_ROM.py:398)")
        if not self._grad_ready:   # 90 Gram blocks of n_r x n_r, built on first use
            self._h.set_rom_gradient(rom_gradient_tensors(self.ops, self.phi))
            self._grad_ready = True
        t_i = time.time()
        out = self._h.rom_gradient(batch, data, in_kind, grad_kind)
        self._check(out)
        self.rom_grad_time += time.time() - t_i
        return out

    def grad_reduced(self, k, data=None):
        """:335-356.  ``k`` nodal field (n,) | (N, n) -> ``(dJ_dk, J)`` with dJ_dk (n,) | (N, n); the observation
        vector is ``set_data``'s (one for all samples) unless ``data`` (n_obs,) | (N, n_obs) is given."""
        kb, single = _as_batch(k, self.dofs, "AffineROMFin.grad_reduced")
        out = self._rom_gradient(kb, _cabi.IN_NODAL, _cabi.IN_NODAL, data)
        return (out["grad"][0], float(out["cost"][0])) if single else (out["grad"], out["cost"])

    def grad_reduced_nine_param(self, k_s, data=None):
        """The same gradient with respect to the nine sub-fin conductivities (the factor ``psi_v_r^T A_phi_w_r`` of
        :349-351 before it is multiplied by ``dsigma_dk``): (9,) | (N, 9) -> ``(g, J)``."""
        tb, single = _as_batch(k_s, self.num_params, "AffineROMFin.grad_reduced_nine_param")
        out = self._rom_gradient(tb, _cabi.IN_PARAMS, _cabi.IN_PARAMS, data)
        return (out["grad"][0], float(out["cost"][0])) if single else (out["grad"], out["cost"])

    def grad_romml(self, k):
        raise NotImplementedError("ROM+NN gradient (averaged_affine_ROM.py:358-396) is out of scope (TensorFlow model)")
