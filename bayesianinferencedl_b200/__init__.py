"""B200-native batched parameter-to-observable forward map of the thermal-fin problem.

Drop-in for the hot path of sheroze1123/BayesianInferenceDL (fom/forward_solve.py, rom/averaged_affine_ROM.py,
fom/thermal_fin.py, bayesian_inference/gaussian_field.py): same class / method names, numpy in, numpy out, CUDA
(sm_100a) kernels behind the C ABI of include/tfin.h.  See DESIGN.md.
"""
from .fom.thermal_fin import get_space, FinSpace, FinMesh, Function
from .fom.forward_solve import Fin
from .rom.averaged_affine_ROM import AffineROMFin
from .bayesian_inference.gaussian_field import make_cov_chol, sample_fields

__all__ = ["get_space", "FinSpace", "FinMesh", "Function", "Fin", "AffineROMFin", "make_cov_chol", "sample_fields"]
__version__ = "0.1.0"
