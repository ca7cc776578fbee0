"""Greedy (model-constrained adaptive sampling) construction of the trial basis.

Reference: ``rom/model_constr_adaptive_sampling.py`` -- ``sample`` (:4-50) repeatedly asks an optimiser for the
parameter with the largest ROM error, solves the full model there and appends the snapshot with ``enrich`` (:52-68).
The control flow is host work in the reference too; what changes here is where the time goes: ``solver.forward`` runs
on the GPU, and :func:`batched_error_optimizer` replaces the black-box scipy search by ONE batched FOM + ROM evaluation
of a whole candidate cloud per greedy step (the reference's optimiser callback signature is kept, so a scipy-based
callback still plugs in).
"""
from __future__ import annotations

import time

import numpy as np

__all__ = ["sample", "enrich", "batched_error_optimizer"]


def enrich(basis, w):
    """Append the snapshot ``w`` (n,) | (n, 1) to ``basis`` (n, k), orthogonalised and normalised.

    Faithful to :52-68 including its loop bound: the new column is orthogonalised (modified Gram-Schmidt, one
    column at a time) against columns ``0 .. k-2`` only -- the reference's ``range(0, k-1)`` skips the last one."""
    basis = np.asarray(basis, dtype=np.float64)
    v = np.array(w, dtype=np.float64).reshape(-1)
    k = basis.shape[1]
    for j in range(k - 1):
        u = basis[:, j]
        v -= (v @ u) / (u @ u) * u
    v /= np.sqrt(v @ v)
    return np.column_stack([basis, v])


def sample(basis, random_initial, optimizer, solver, tol=1.0e-14, maxiter=80, verbose=True):
    """:4-50.  ``optimizer(z_0, basis, solver) -> (z_star, g_z_star)`` returns the parameter with the largest ROM
    error and that error; ``solver.forward(z_star)`` must return the 5-tuple of ``Fin.forward``."""
    g_z_star, iterations = 1e30, 0
    while g_z_star > tol and iterations < maxiter:
        t0 = time.time()
        previous = g_z_star
        z_star, g_z_star = optimizer(random_initial(), basis, solver)
        iterations += 1
        snapshot = np.asarray(solver.forward(z_star)[0], dtype=np.float64)
        basis = enrich(basis, snapshot)
        if verbose:
            print(f"greedy step {iterations}: error {g_z_star:.3e} (improvement {previous - g_z_star:.3e}), "
                  f"{time.time() - t0:.2f} s")
    if verbose:
        print(f"sampling completed after {iterations} iterations")
    return basis


def batched_error_optimizer(n_candidates=4096, draw=None, seed=0):
    """Optimiser callback for :func:`sample` that evaluates a cloud of candidates in one batch: for every candidate
    field the full-order observables (``solver.forward_qoi``) and the LSPG observables on the CURRENT basis
    (``solver.r_fwd_no_full_qoi``) are computed on the GPU; the candidate with the largest discrepancy wins.
    ``draw(rng, n_candidates, z_0) -> (n_candidates, n)`` proposes candidates (default: log-normal perturbations of
    ``z_0``)."""
    rng = np.random.default_rng(seed)

    def optimizer(z_0, basis, solver):
        z_0 = np.asarray(z_0, dtype=np.float64)
        if draw is not None:
            cand = draw(rng, n_candidates, z_0)
        else:
            cand = z_0[None, :] * np.exp(0.5 * rng.standard_normal((n_candidates, 1)) +
                                         0.25 * rng.standard_normal((n_candidates, z_0.shape[0])))
        q = solver.forward_qoi(cand)
        q_r = solver.r_fwd_no_full_qoi(cand, basis)
        err = 0.5 * np.sum((q - q_r) ** 2, axis=1)
        best = int(np.argmax(err))
        return cand[best], float(err[best])

    return optimizer
