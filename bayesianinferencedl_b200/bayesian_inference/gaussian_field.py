"""Gaussian-random-field prior of the nodal conductivity, on the device.

Reference: ``bayesian_inference/gaussian_field.py:9-31`` (``make_cov_chol``: pairwise dof distances -> squared
exponential / Matern-5/2 / Matern-3/2 kernel -> upper Cholesky factor) and the draw
``exp(0.5 * chol.T @ randn)`` of ``deep_learning/generate_fin_dataset.py:87-88``.

Here the covariance, its blocked Cholesky factorisation and the batched draws (fp64 triangular GEMM with the exp
fused, optional on-device Philox normals) all run in libtfin.so (csrc/field.cuh); there is no host path.
``make_cov_chol`` still RETURNS the upper factor as a numpy array, so reference call sites (``chol.T @ norm``,
``len(chol)``) keep working unchanged.
"""
from __future__ import annotations

import numpy as np

from .. import _cabi

__all__ = ["make_cov_chol", "FieldSampler", "sample_fields"]

_KERNELS = {"sq_exp": _cabi.KERN_SQ_EXP, "m52": _cabi.KERN_M52}


class FieldSampler:
    """Device-resident prior: owns the Cholesky factor and draws conductivity fields in batches."""

    def __init__(self, V=None, kern_type="m52", length=1.6, *, chol=None, device=0, handle=None):
        self._h = handle if handle is not None else _cabi.TfinHandle(device)
        if chol is not None:
            self.chol = np.ascontiguousarray(chol, dtype=np.float64)
            self._h.field_set_chol(self.chol)
        else:
            # the reference falls through to Matern-3/2 for any name other than 'sq_exp' / 'm52' (:25-28)
            kind = _KERNELS.get(kern_type, _cabi.KERN_M32)
            pts = V.tabulate_dof_coordinates().reshape((-1, 2))[V.dofmap().dofs(), :]
            self.chol = self._h.field_set_cov(pts, kind, length)
        self.n = self.chol.shape[0]

    @property
    def handle(self):
        return self._h

    def sample(self, N=None, z=None, seed=0, subsequence=0, first_row=0, return_z=False):
        """``exp(0.5 * chol.T @ z)`` row by row.  ``z`` (n,) | (N, n) standard normals, or ``N`` draws from the
        device generator (Philox4x32-10 + Box-Muller; row r is the stream of (seed, first_row + r, subsequence))."""
        if z is not None:
            z = np.asarray(z, dtype=np.float64)
            if z.ndim == 1:
                return self._h.field_sample(z=z[None, :])[0]
            return self._h.field_sample(z=z)
        return self._h.field_sample(N=N, seed=seed, subsequence=subsequence, first_row=first_row, want_z=return_z)


def make_cov_chol(V, kern_type="m52", length=1.6):
    """Upper Cholesky factor of the prior covariance over the dofs of ``V`` (computed on the device)."""
    return FieldSampler(V, kern_type, length).chol


def sample_fields(chol, z, device=0):
    """Batched ``exp(0.5 * chol.T @ z)`` for a given upper factor: z (n,) | (N, n)."""
    return FieldSampler(chol=chol, device=device).sample(z=z)
