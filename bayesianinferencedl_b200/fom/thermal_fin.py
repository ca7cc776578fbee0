"""Thermal-fin geometry, mesh and P1 space factory.

Drop-in for the reference's ``fom/thermal_fin.py:4-20`` (``get_space(resolution)``).  The reference
builds the union of nine rectangles with mshr/CGAL and wraps it in a dolfin ``FunctionSpace``; neither
library exists here and the mshr mesh is not shipped, so this module builds a *conforming structured*
triangulation of exactly the same geometry (every rectangle corner of ``thermal_fin.py:7-15`` is a mesh
vertex) and a light-weight space object exposing the handful of ``FunctionSpace`` methods the hot path
touches (``dim``, ``mesh``, ``dofmap().dofs()``, ``tabulate_dof_coordinates``).

dof index == vertex index of the mesh.  An externally produced mesh (e.g. the
exported mshr mesh) enters through :meth:`FinSpace.from_mesh`.
"""
from __future__ import annotations

import numpy as np

__all__ = ["get_space", "get_space_unstructured", "unstructured_fin_mesh", "FinSpace", "FinMesh", "Function",
           "resolution_to_m"]

# Geometry of fom/thermal_fin.py:7-15 in units of 0.25 (all corners are multiples of 0.25).
_POST_X = (10, 14)          # x in [2.5, 3.5]
_POST_Y = (0, 16)           # y in [0, 4]
_FIN_BANDS = (3, 7, 11, 15)  # sub-fins occupy y in [0.75,1], [1.75,2], [2.75,3], [3.75,4]
_LEFT_X = (0, 10)           # x in [0, 2.5]
_RIGHT_X = (14, 24)         # x in [3.5, 6]


def resolution_to_m(resolution) -> int:
    """Map the reference's mshr ``resolution`` to the structured refinement level ``m`` (h = 0.25/m).

    mshr resolution 40 gives 1446 dofs in the reference (data/B_obs.txt has 1446 columns); m = 3 gives
    1597 dofs, the closest structured size, so 40 -> 3 and the mapping is linear around it.
    """
    return max(1, int(round(float(resolution) * 3.0 / 40.0)))


class FinMesh:
    """Triangle mesh: ``coordinates()`` (nv,2) float64 and ``cells()`` (nc,3) int32, like dolfin.Mesh."""

    def __init__(self, coords, cells):
        self._x = np.ascontiguousarray(coords, dtype=np.float64)
        self._c = np.ascontiguousarray(cells, dtype=np.int32)
        if self._x.ndim != 2 or self._x.shape[1] != 2:
            raise ValueError("coords must have shape (n_vertices, 2)")
        if self._c.ndim != 2 or self._c.shape[1] != 3:
            raise ValueError("cells must have shape (n_cells, 3)")
        if self._c.size and (self._c.min() < 0 or self._c.max() >= len(self._x)):
            raise ValueError("cell vertex index out of range")

    def coordinates(self):
        return self._x

    def cells(self):
        return self._c

    def num_vertices(self):
        return self._x.shape[0]

    def num_cells(self):
        return self._c.shape[0]


class _DofMap:
    def __init__(self, n):
        self._n = n

    def dofs(self):
        return np.arange(self._n, dtype=np.int64)


class FinSpace:
    """P1 ('CG', 1) space on a :class:`FinMesh`; dof i lives on vertex i."""

    def __init__(self, mesh: FinMesh, m=None):
        self._mesh = mesh
        self.m = m

    @classmethod
    def from_mesh(cls, coords, cells):
        return cls(FinMesh(coords, cells), m=None)

    def mesh(self):
        return self._mesh

    def dim(self):
        return self._mesh.num_vertices()

    def dofmap(self):
        return _DofMap(self.dim())

    def tabulate_dof_coordinates(self):
        return self._mesh.coordinates()


class Function(np.ndarray):
    """Nodal P1 field.  An ndarray of shape (n,) or (N, n) that also answers the dolfin idioms the
    reference's callers use on forward-solve results: ``w.vector()[:]``, ``w.vector().get_local()``,
    ``w.vector().set_local(v)``, ``w.assign(other)``."""

    def __new__(cls, V_or_values, values=None):
        if values is None and isinstance(V_or_values, FinSpace):
            arr = np.zeros(V_or_values.dim(), dtype=np.float64)
        else:
            arr = np.asarray(V_or_values if values is None else values, dtype=np.float64)
        return arr.view(cls)

    def vector(self):
        return self

    def get_local(self):
        return np.asarray(self)

    def set_local(self, v):
        self[...] = np.asarray(v, dtype=np.float64).reshape(self.shape)

    def assign(self, other):
        self.set_local(np.asarray(other))


def _build_structured(m: int):
    """Conforming structured triangulation with h = 0.25/m.

    Vertices are numbered sub-domain by sub-domain (post row-major first, then each sub-fin row-major)
    so that the dofs of one sub-domain are contiguous; squares left of x=3 are cut by the '\\' diagonal
    and squares right of it by '/', which makes the mesh mirror-symmetric about the post axis.
    """
    if m < 1:
        raise ValueError("m must be >= 1")
    nx, ny = 24 * m, 16 * m
    inside = np.zeros((ny, nx), dtype=bool)           # inside[j, i]: square with lower-left (i, j)
    inside[_POST_Y[0] * m:_POST_Y[1] * m, _POST_X[0] * m:_POST_X[1] * m] = True
    for b in _FIN_BANDS:
        inside[b * m:(b + 1) * m, _LEFT_X[0] * m:_LEFT_X[1] * m] = True
        inside[b * m:(b + 1) * m, _RIGHT_X[0] * m:_RIGHT_X[1] * m] = True

    vid = -np.ones((ny + 1, nx + 1), dtype=np.int64)
    order = []

    def claim(j0, j1, i0, i1):
        for j in range(j0, j1 + 1):
            row = np.arange(i0, i1 + 1)
            new = row[vid[j, row] < 0]
            vid[j, new] = len(order) + np.arange(len(new))
            order.extend((j, i) for i in new)

    claim(_POST_Y[0] * m, _POST_Y[1] * m, _POST_X[0] * m, _POST_X[1] * m)
    for b in _FIN_BANDS:
        claim(b * m, (b + 1) * m, _LEFT_X[0] * m, _LEFT_X[1] * m)
    for b in _FIN_BANDS:
        claim(b * m, (b + 1) * m, _RIGHT_X[0] * m, _RIGHT_X[1] * m)

    ji = np.asarray(order, dtype=np.int64)
    # (i * 0.25) / m is exact whenever the quotient is representable, so the rectangle corners of
    # thermal_fin.py:7-15 (2.5, 3.5, 0.75, ...) come out bit-exact.
    coords = np.stack([(ji[:, 1] * 0.25) / m, (ji[:, 0] * 0.25) / m], axis=1)

    jj, ii = np.nonzero(inside)
    v00, v10 = vid[jj, ii], vid[jj, ii + 1]
    v01, v11 = vid[jj + 1, ii], vid[jj + 1, ii + 1]
    left = ii < 12 * m
    # '\' : (v00,v10,v01) + (v10,v11,v01);   '/' : (v00,v10,v11) + (v00,v11,v01); all counter-clockwise
    t1 = np.where(left[:, None], np.stack([v00, v10, v01], 1), np.stack([v00, v10, v11], 1))
    t2 = np.where(left[:, None], np.stack([v10, v11, v01], 1), np.stack([v00, v11, v01], 1))
    cells = np.empty((2 * len(jj), 3), dtype=np.int32)
    cells[0::2] = t1
    cells[1::2] = t2
    return coords, cells


def get_space(resolution, m=None):
    """``get_space(resolution)`` of fom/thermal_fin.py:4-20.  ``m`` (keyword) overrides the mapping
    resolution -> refinement level; dofs = 144 m^2 + 100 m + 1 (m=3: 1597, m=26: 99 945)."""
    if m is None:
        m = resolution_to_m(resolution)
    coords, cells = _build_structured(int(m))
    return FinSpace(FinMesh(coords, cells), m=int(m))


def _inside_fin(p):
    x, y = p[:, 0], p[:, 1]
    post = (x >= 2.5) & (x <= 3.5) & (y >= 0) & (y <= 4)
    band = np.zeros(len(p), dtype=bool)
    for yb in (0.75, 1.75, 2.75, 3.75):
        band |= (y >= yb) & (y <= yb + 0.25)
    return post | (band & (x >= 0) & (x <= 6))


def unstructured_fin_mesh(h=0.125, seed=0, jitter=0.3):
    """UNSTRUCTURED, non-conforming triangulation of the fin (Delaunay of jittered lattice points): the stand-in for the
    reference's mshr mesh (``mshr.generate_mesh(geometry, 40)``, fom/thermal_fin.py:17, not shipped and not reproducible
    without CGAL).  Like that mesh it has varying vertex degrees (6-7 non-zeros per row on average, up to 9-10) and cells
    that straddle x = 2.5 / 3.5, which keep marker 0 under dolfin's ``SubDomain.mark`` (SURVEY Q-1).  h = 0.0925 gives
    1439 dofs (the reference mesh has 1446).  Returns (coords, cells)."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    nx, ny = int(round(6 / h)), int(round(4 / h))
    X, Y = np.meshgrid(np.arange(nx + 1) * (6.0 / nx), np.arange(ny + 1) * (4.0 / ny))
    pts = np.stack([X.ravel(), Y.ravel()], axis=1)
    # lattice lines rarely hit the band edges y = 0.75 + 0.25 i exactly: add the geometry's own boundary points
    hx, hy = 6.0 / nx, 4.0 / ny
    extra = []
    for yb in (0.75, 1.0, 1.75, 2.0, 2.75, 3.0, 3.75, 4.0):
        xs = np.concatenate([np.arange(0, 2.5 + 1e-9, hx), np.arange(3.5, 6 + 1e-9, hx)])
        extra.append(np.stack([xs, np.full_like(xs, yb)], axis=1))
    for xb in (0.0, 2.5, 3.5, 6.0):
        ys = np.arange(0, 4 + 1e-9, hy)
        extra.append(np.stack([np.full_like(ys, xb), ys], axis=1))
    pts = np.concatenate([pts] + extra)
    pts = pts[_inside_fin(pts)]
    # drop lattice points that crowd a boundary point (closer than 0.4 h), keeping the boundary ones
    pts = np.unique(np.round(pts, 12), axis=0)
    on_edge = np.zeros(len(pts), dtype=bool)
    for yb in (0.75, 1.0, 1.75, 2.0, 2.75, 3.0, 3.75, 4.0, 0.0):
        on_edge |= np.abs(pts[:, 1] - yb) < 1e-12
    for xb in (0.0, 2.5, 3.5, 6.0):
        on_edge |= np.abs(pts[:, 0] - xb) < 1e-12
    from scipy.spatial import cKDTree
    tree = cKDTree(pts[on_edge])
    d, _ = tree.query(pts, k=1)
    pts = pts[on_edge | (d > 0.4 * min(hx, hy))]
    eps = 1e-9
    interior = np.ones(len(pts), dtype=bool)
    for dx, dy in ((hx, 0), (-hx, 0), (0, hy), (0, -hy), (hx, hy), (-hx, -hy), (hx, -hy), (-hx, hy)):
        interior &= _inside_fin(pts + np.array([dx, dy]) * (1 - eps))
    pts = pts.copy()
    pts[interior] += rng.uniform(-jitter, jitter, (int(interior.sum()), 2)) * np.array([hx, hy])
    cells = Delaunay(pts).simplices
    cen = pts[cells].mean(axis=1)
    keep = _inside_fin(cen)
    for a, b in ((0, 1), (1, 2), (2, 0)):
        keep &= _inside_fin((pts[cells[:, a]] + pts[cells[:, b]]) / 2)
    cells = cells[keep]
    a, b, c = pts[cells[:, 0]], pts[cells[:, 1]], pts[cells[:, 2]]
    det = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])
    cells, det = cells[np.abs(det) > 1e-10], det[np.abs(det) > 1e-10]
    cells[det < 0] = cells[det < 0][:, [0, 2, 1]]
    used = np.unique(cells)
    remap = -np.ones(len(pts), dtype=np.int64)
    remap[used] = np.arange(len(used))
    return pts[used], remap[cells].astype(np.int32)


def get_space_unstructured(h=0.0925, seed=0):
    """P1 space on :func:`unstructured_fin_mesh` -- a reference-like (mshr-like) mesh of about 1446 dofs by default."""
    return FinSpace.from_mesh(*unstructured_fin_mesh(h=h, seed=seed))
