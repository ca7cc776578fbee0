"""Matern / squared-exponential covariance factor over the dof coordinates.

Mirrors ``bayesian_inference/gaussian_field.py:9-31`` (``make_cov_chol``): pairwise dof distances, kernel,
UPPER Cholesky factor; conductivity samples are ``exp(0.5 * chol.T @ z)`` (generate_fin_dataset.py:87-88).
One-time O(n^3) set-up on the host, exactly like the reference; only the per-sample use is on the hot path.
"""
from __future__ import annotations

import numpy as np
from scipy import linalg, spatial

__all__ = ["make_cov_chol", "sample_fields"]


def make_cov_chol(V, kern_type="m52", length=1.6):
    Wdofs_x = V.tabulate_dof_coordinates().reshape((-1, 2))
    V0_dofs = V.dofmap().dofs()
    points = Wdofs_x[V0_dofs, :]
    dists = spatial.distance.squareform(spatial.distance.pdist(points))
    if kern_type == "sq_exp":
        alpha = 1 / (2 * length ** 2)
        noise_var = 1e-5
        cov = np.exp(-alpha * dists ** 2) + np.eye(len(points)) * noise_var
    elif kern_type == "m52":
        tmp = np.sqrt(5) * dists / length
        cov = (1 + tmp + tmp * tmp / 3) * np.exp(-tmp)
    else:
        tmp = np.sqrt(3) * dists / length
        cov = (1 + tmp) * np.exp(-tmp)
    return linalg.cholesky(cov)


def sample_fields(chol, z):
    """``exp(0.5 * chol.T @ z)`` for one draw (n,) or a batch (N, n) of standard normals."""
    z = np.asarray(z, dtype=np.float64)
    return np.exp(0.5 * (z @ chol))
