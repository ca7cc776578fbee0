"""Sampler-facing operators with Theano's ``Op.perform(node, inputs, outputs)`` calling convention.

Reference call-ins: ``ParamToObsFOM`` (``bayesian_inference/inference.py:21-57``: qoi and the sensitivity
matrix of the ``exp(k)`` full-order model), ``SqErrorOpFOM`` / ``SqErrorOpROM``
(``bayesian_inference/pymc_func_bayes_inverse.py:106-151``: misfit value and gradient).  ``perform`` receives the inputs
as a list and writes each result into ``outputs[i][0]`` -- exactly what ``theano.Op`` requires -- so with Theano
installed these classes become real ops by ``class Op(theano.Op, ParamToObsFOM): itypes = ...; otypes = ...``; without
it (this image) they are plain callables.  The one extension: an input with a leading batch axis is evaluated in a
single batched kernel launch (many chains / many proposals at once).
"""
from __future__ import annotations

import numpy as np

from ..fom.forward_solve_exp import Fin as FinExp
from .likelihood import SqError

__all__ = ["ParamToObsFOM", "SqErrorOpFOM", "SqErrorOpROM"]


class ParamToObsFOM:
    """inference.py:21-57: ``pred_k`` (log-conductivity, n) -> ``[qoi (n_obs), sensitivity (n_obs, n)]``."""
    __props__ = ()

    def __init__(self, V, randobs, **kw):
        self._V = V
        self._solver = FinExp(V, randobs, **kw)

    def perform(self, node, inputs, outputs):
        pred_k = np.asarray(inputs[0], dtype=np.float64)
        single = pred_k.ndim == 1
        kb = pred_k[None, :] if single else pred_k
        out = self._solver.handle.fom_nodal_sensitivity(kb, tol=self._solver.tol, maxit=self._solver.maxit)
        self._solver._raise_on_failure(out)
        outputs[0][0] = out["qoi"][0] if single else out["qoi"]
        outputs[1][0] = out["jac"][0] if single else out["jac"]

    def __call__(self, pred_k):
        outputs = [[None], [None]]
        self.perform(None, [pred_k], outputs)
        return outputs[0][0], outputs[1][0]

    def vjp(self, pred_k, output_gradient):
        """The contraction ``grad`` builds symbolically (:53-56): ``sensitivity^T @ output_gradient``."""
        _, jac = self(pred_k)
        g = np.asarray(output_gradient, dtype=np.float64)
        return jac.T @ g if jac.ndim == 2 else np.einsum("son,so->sn", jac, g)


class _SqErrorOp:
    __props__ = ()
    _method = None

    def __init__(self, V, chol, randobs, **kw):
        self._error_op = SqError(V, chol, randobs, **kw)

    def perform(self, node, inputs, outputs):
        value, grad = getattr(self._error_op, self._method)(inputs[0])
        outputs[0][0] = np.asarray(value)
        outputs[1][0] = grad

    def __call__(self, pred_k):
        outputs = [[None], [None]]
        self.perform(None, [pred_k], outputs)
        return outputs[0][0], outputs[1][0]


class SqErrorOpFOM(_SqErrorOp):
    """pymc_func_bayes_inverse.py:130-151."""
    _method = "err_grad_FOM"


class SqErrorOpROM(_SqErrorOp):
    """pymc_func_bayes_inverse.py:106-128 (needs ``phi=...``)."""
    _method = "err_grad_ROM"
