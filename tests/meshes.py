"""Test-only mesh generators: an UNSTRUCTURED, non-conforming triangulation of the fin (Delaunay of jittered points),
standing in for the reference's mshr mesh: vertex degrees vary (ELL width > 7) and cells straddle x = 2.5 / 3.5, so
they keep marker 0 exactly like SURVEY Q-1 describes."""
import numpy as np
from scipy.spatial import Delaunay


def _inside(p):
    x, y = p[:, 0], p[:, 1]
    post = (x >= 2.5) & (x <= 3.5) & (y >= 0) & (y <= 4)
    band = np.zeros(len(p), dtype=bool)
    for yb in (0.75, 1.75, 2.75, 3.75):
        band |= (y >= yb) & (y <= yb + 0.25)
    fins = band & (x >= 0) & (x <= 6)
    return post | fins


def unstructured_fin(h=0.125, seed=0, jitter=0.3):
    rng = np.random.default_rng(seed)
    nx, ny = int(round(6 / h)), int(round(4 / h))
    X, Y = np.meshgrid(np.arange(nx + 1) * h, np.arange(ny + 1) * h)
    pts = np.stack([X.ravel(), Y.ravel()], axis=1)
    pts = pts[_inside(pts)]
    # jitter interior points only (points whose whole h-neighbourhood is inside), but NOT along x = 2.5/3.5 lines,
    # and shift a column of post points so that no mesh edge lies on x = 2.5 / 3.5 inside the fins' bands
    eps = 1e-9
    interior = np.ones(len(pts), dtype=bool)
    for dx, dy in ((h, 0), (-h, 0), (0, h), (0, -h), (h, h), (-h, -h), (h, -h), (-h, h)):
        interior &= _inside(pts + np.array([dx, dy]) * (1 - eps))
    pts = pts.copy()
    pts[interior] += rng.uniform(-jitter * h, jitter * h, (interior.sum(), 2))
    tri = Delaunay(pts)
    cells = tri.simplices
    cen = pts[cells].mean(axis=1)
    mids = [(pts[cells[:, a]] + pts[cells[:, b]]) / 2 for a, b in ((0, 1), (1, 2), (2, 0))]
    keep = _inside(cen)
    for m in mids:
        keep &= _inside(m)
    cells = cells[keep]
    # orient counter-clockwise, drop slivers
    a, b, c = pts[cells[:, 0]], pts[cells[:, 1]], pts[cells[:, 2]]
    det = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])
    cells = cells[np.abs(det) > 1e-10]
    det = det[np.abs(det) > 1e-10]
    cells[det < 0] = cells[det < 0][:, [0, 2, 1]]
    used = np.unique(cells)
    remap = -np.ones(len(pts), dtype=np.int64)
    remap[used] = np.arange(len(used))
    return pts[used], remap[cells].astype(np.int32)
