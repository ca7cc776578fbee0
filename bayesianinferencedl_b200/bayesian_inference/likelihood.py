"""Batched squared-misfit likelihood and the many-chain pCN driver built on it.

Reference: ``bayesian_inference/pymc_func_bayes_inverse.py`` -- ``SqError`` (:25-104: ``err_grad_FOM`` :68-78,
``err_grad_ROM`` :80-90), the Theano ops that call it once per sampler proposal (:106-167; ``inference.py:21-57``),
and the potential ``sq_err(nodal_vals)[0] / sigma / sigma`` (:201).  The samplers themselves (PyMC3 NUTS, MUQ) stay out of
scope; what is rebuilt is the part that bounds their throughput -- the likelihood -- evaluated for MANY conductivity
fields (chains / proposals) per call, plus a minimal device-resident sampler that shows the throughput end to end:
preconditioned Crank-Nicolson Metropolis over the Gaussian-field coordinates, every chain advanced by the same five
kernels per step (csrc/chains.cuh), statistics accumulated on the device and reduced across ranks (dist.py).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _cabi
from ..fom.forward_solve import Fin, _as_batch
from ..rom.averaged_affine_ROM import AffineROMFin
from .gaussian_field import FieldSampler

__all__ = ["SqError", "PCNChains"]


class SqError:
    """``SqError(V, chol, randobs)`` (:30-66) with the data-dependent pieces as keyword arguments, because the
    reference hard-codes files of its own mesh (``res_x.npy``, ``../data/basis_nine_param.txt``):
    ``k_true`` nodal truth field (default: ``exp(0.5 chol^T z)`` for the Philox draw (seed, row 0)), ``phi`` reduced
    basis (default: no ROM).  ``err_grad_*`` accept one field (n,) like the reference or a batch (N, n)."""

    def __init__(self, V, chol, randobs=False, *, k_true=None, phi=None, seed=0, device=0):
        self._V = V
        self._solver = Fin(V, randobs, device=device)
        self.prior = FieldSampler(chol=chol, handle=self._solver.handle)
        if k_true is None:
            k_true = self.prior.sample(N=1, seed=seed)[0]
        self.k_true = np.asarray(k_true, dtype=np.float64)
        self.obs_data = self._solver.forward_qoi(self.k_true)             # :52-55
        self.phi = None
        self._solver_r = None
        if phi is not None:
            self.phi = np.asarray(phi, dtype=np.float64)
            self._solver_r = AffineROMFin(V, None, self.phi, randobs, device=device)   # :63-65
            self._solver_r.set_data(self.obs_data)

    def err_grad_FOM(self, pred_k):
        """:68-78: ``(0.5 ||qoi - obs||^2, dJ/dk)`` from one fused forward + adjoint kernel."""
        kb, single = _as_batch(pred_k, self._solver.dofs, "SqError.err_grad_FOM")
        grad, err = self._solver.gradient(kb, self.obs_data, return_cost=True)
        return (float(err[0]), grad[0]) if single else (err, grad)

    def err_grad_ROM(self, pred_k):
        """:80-90."""
        if self._solver_r is None:
            raise RuntimeError("SqError.err_grad_ROM: no reduced basis was given (phi=...)")
        grad, err = self._solver_r.grad_reduced(pred_k)
        return err, grad

    def err_grad_ROMML(self, pred_k):
        raise NotImplementedError("ROM + neural-network correction (:92-104) needs the Keras error model: out of scope")


class PCNChains:
    """Many-chain pCN Metropolis for ``pi(z) ~ N(0, I) exp(-0.5 ||qoi(k(z)) - data||^2 / sigma^2)``,
    ``k = exp(0.5 chol^T z)``.  ``model``: a ``Fin`` ('fom', nodal full-order likelihood) or an ``AffineROMFin``
    ('rom', sub-fin-averaged reduced likelihood)."""

    def __init__(self, solver, chol, data, sigma, *, seed=0):
        self.solver = solver
        self.model = 1 if isinstance(solver, AffineROMFin) else 0
        self.prior = FieldSampler(chol=chol, handle=solver.handle)
        self.data = np.ascontiguousarray(data, dtype=np.float64)
        if self.data.shape != (solver.n_obs,):
            raise ValueError(f"data must have shape ({solver.n_obs},)")
        self.sigma, self.seed = float(sigma), int(seed)
        self.n, self.n_obs = solver.dofs, solver.n_obs
        self.steps_done = 0
        self.z = None
        self.first_chain = 0

    def run(self, n_steps, n_chains=None, beta=0.2, z0=None, first_chain=0, want_k_mean=False):
        """Advance all chains by ``n_steps``.  The first call starts them (``z0`` (C, n) or ``n_chains`` draws from
        the prior); later calls continue.  Returns per-chain statistics of THIS call: accepted counts, final misfit and
        observables, sums / sums of squares of the observables over the steps, optionally the summed field."""
        h = self.solver.handle
        if self.z is None:
            if z0 is not None:
                self.z = np.ascontiguousarray(z0, dtype=np.float64).copy()
                if self.z.ndim != 2 or self.z.shape[1] != self.n:
                    raise ValueError(f"z0 must be (C, {self.n})")
                init = 0
            else:
                if not n_chains:
                    raise ValueError("give n_chains or z0 on the first call")
                self.z = np.empty((int(n_chains), self.n))
                init = 1
            self.first_chain = int(first_chain)
        else:
            init = 0
        Cn = self.z.shape[0]
        out = {"accepted": np.zeros(Cn, dtype=np.int64), "misfit": np.empty(Cn), "qoi": np.empty((Cn, self.n_obs)),
               "qoi_sum": np.empty((Cn, self.n_obs)), "qoi_sq": np.empty((Cn, self.n_obs)),
               "k_sum": np.empty((Cn, self.n)) if want_k_mean else None}
        p = _cabi._ptr
        rc = h._lib.tfin_pcn_chains(h._h, self.model, Cn, self.first_chain, int(n_steps), self.steps_done, float(beta),
                                    p(self.data), self.sigma, self.seed, self.solver.tol, self.solver.maxit,
                                    _cabi.MEM_HOST, p(self.z), init, p(out["misfit"]), p(out["accepted"]),
                                    p(out["qoi"]), p(out["qoi_sum"]), p(out["qoi_sq"]), p(out["k_sum"]), None)
        _cabi._check(h._lib, rc, "tfin_pcn_chains")
        self.steps_done += int(n_steps)
        out["z"] = self.z
        out["n_steps"] = int(n_steps)
        return out

    @staticmethod
    def summarize(out, group=None, device=None):
        """Posterior summaries over chains and steps: mean / std of the observables and the acceptance rate, reduced
        over all ranks of ``group`` when torch.distributed is initialised (dist.chain_moments)."""
        from ..dist import chain_moments
        count = out["qoi_sum"].shape[0] * out["n_steps"]
        cnt, mean, var = chain_moments(count, out["qoi_sum"].sum(0), out["qoi_sq"].sum(0), group=group, device=device)
        acc = out["accepted"].sum() / max(count, 1)
        return {"count": cnt, "qoi_mean": mean, "qoi_std": np.sqrt(np.maximum(var, 0.0)), "accept_rate": float(acc)}
