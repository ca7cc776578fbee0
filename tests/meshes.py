"""Test-side alias of the package's mshr stand-in: an UNSTRUCTURED, non-conforming triangulation of the fin (Delaunay of
jittered points); vertex degrees vary (ELL width > 7) and cells straddle x = 2.5 / 3.5, so they keep marker 0 exactly like
SURVEY Q-1 describes."""
from bayesianinferencedl_b200.fom.thermal_fin import unstructured_fin_mesh


def unstructured_fin(h=0.125, seed=0, jitter=0.3):
    return unstructured_fin_mesh(h=h, seed=seed, jitter=jitter)
