"""``Fin`` with the log-conductivity parametrisation ``exp(k)`` -- the variant ``bayesian_inference/inference.py:18``
imports (``fom/forward_solve_exp.py``: ``_F = inner(exp(k) grad w, grad v) dx + Bi w v ds``, :160-161).

Same class as :mod:`.forward_solve` with one switch in the kernels (``nodal_coef_mode = 1``): the per-cell coefficient
of the in-kernel assembly is the quadrature of ``exp(k)`` dolfin's form compiler would use (6-point degree-3 Strang-Fix
rule) instead of the vertex mean, and the gradient form ``k_hat exp(k) grad z . grad v`` of ``gradient`` (:277-310) and
``sensitivity`` (:312-342) uses the degree-4 rule.  The Hessian / Fisher actions (:344-395) are not built.
"""
from __future__ import annotations

from .forward_solve import Fin as _Fin

__all__ = ["Fin"]


class Fin(_Fin):
    def __init__(self, V, external_obs=False, **kw):
        super().__init__(V, external_obs, **kw)
        self._h.set_int("nodal_coef_mode", 1)
