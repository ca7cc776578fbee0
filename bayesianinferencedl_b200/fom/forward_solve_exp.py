"""``Fin`` with the log-conductivity parametrisation ``exp(k)`` -- the variant ``bayesian_inference/inference.py:18``
imports (``fom/forward_solve_exp.py``: ``_F = inner(exp(k) grad w, grad v) dx + Bi w v ds``, :160-161).

Same class as :mod:`.forward_solve` with one switch in the kernels: the per-cell coefficient of the in-kernel assembly
is the quadrature of ``exp(k)`` dolfin's form compiler would use (6-point degree-3 Strang-Fix rule) instead of the
vertex mean.  ``forward`` / ``forward_qoi`` / ``r_fwd_no_full`` and the pCN likelihood run in this mode; the adjoint
``gradient`` / ``sensitivity`` of the exp form (:277-342, a degree-4 rule inside the gradient form) are not built.
"""
from __future__ import annotations

from .forward_solve import Fin as _Fin

__all__ = ["Fin"]


class Fin(_Fin):
    def __init__(self, V, external_obs=False, **kw):
        super().__init__(V, external_obs, **kw)
        self._h.set_int("nodal_coef_mode", 1)

    def gradient(self, k, data, return_cost=False):
        raise NotImplementedError("adjoint gradient of the exp(k) form (forward_solve_exp.py:277-310) is not built")

    def sensitivity(self, k):
        raise NotImplementedError("sensitivity of the exp(k) form (forward_solve_exp.py:312-342) is not built")
