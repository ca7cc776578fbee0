"""B200-native batched parameter-to-observable forward map of the thermal-fin problem.

Drop-in for the hot path of sheroze1123/BayesianInferenceDL (fom/forward_solve.py, fom/forward_solve_exp.py,
rom/averaged_affine_ROM.py, fom/thermal_fin.py, bayesian_inference/gaussian_field.py and the batched call sites in
deep_learning/generate_fin_dataset.py and bayesian_inference/pymc_func_bayes_inverse.py): same class / method names,
numpy in, numpy out, CUDA (sm_100a) kernels behind the C ABI of include/tfin.h.  See DESIGN.md.

Sub-modules that are imported on demand (they pull in torch for device buffers or are sampler-facing):
``deep_learning.generate_fin_dataset`` (gen_affine_avg_rom_dataset), ``bayesian_inference.likelihood`` (SqError,
PCNChains), ``bayesian_inference.ops`` (perform-protocol ops), ``fom.forward_solve_exp`` (exp(k) model),
``rom.pod`` / ``rom.model_constr_adaptive_sampling`` (basis construction), ``dist`` (sharding over GPUs).
"""
from .fom.thermal_fin import get_space, FinSpace, FinMesh, Function
from .fom.forward_solve import Fin
from .rom.averaged_affine_ROM import AffineROMFin
from .bayesian_inference.gaussian_field import FieldSampler, make_cov_chol, sample_fields

__all__ = ["get_space", "FinSpace", "FinMesh", "Function", "Fin", "AffineROMFin", "make_cov_chol", "sample_fields",
           "FieldSampler"]
__version__ = "0.1.0"
