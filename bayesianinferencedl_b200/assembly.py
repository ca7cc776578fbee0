"""Host-side P1 operator builder for the thermal fin (vectorised numpy; runs once per space).

Produces everything the CUDA path needs as flat arrays on ONE shared CSR pattern:

* ``vals[0]``      Bi * M_Gamma (Robin mass on the exterior boundary)   forward_solve.py:160-161
* ``vals[q]``      K_q, stiffness over cells with marker q = 1..9       averaged_affine_ROM.py:156-162
* ``rhs``          b_j = int_{root} phi_j                               forward_solve.py:162-163
* ``Ke``           per-element 3x3 stiffness (nodal-conductivity path)  forward_solve.py:160
* ``B_obs``        9 x n sub-fin averaging operator                      forward_solve.py:488-511
* ``C``            whole-domain averaging operator                      forward_solve.py:396-406

Marking follows dolfin's ``SubDomain.mark`` (an entity is marked iff all its vertices AND its midpoint
satisfy ``inside``; later marks overwrite earlier ones) with ``between``/``near`` tolerances of
DOLFIN_EPS, see SURVEY.md appendix A.1.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

DOLFIN_EPS = 3.0e-16
BIOT = 0.1                      # forward_solve.py:112, averaged_affine_ROM.py:83
NUM_SUBDOMAINS = 9

# (y_b, is_left) of fin1..fin9 as instantiated in forward_solve.py:125-133 (None = centre post)
_SUBFIN_DEFS = ((0.75, True), (1.75, True), (2.75, True), (3.75, True), None,
                (3.75, False), (2.75, False), (1.75, False), (0.75, False))


def _between(x, lo, hi):
    return (x >= lo - DOLFIN_EPS) & (x <= hi + DOLFIN_EPS)


def _inside_subdomain(q, x, y):
    """Predicate of SubFin.inside / CenterFin.inside (forward_solve.py:11-15, 31-33) for id q (1..9)."""
    d = _SUBFIN_DEFS[q - 1]
    if d is None:
        return _between(x, 2.5, 3.5)
    y_b, is_left = d
    xr = (0.0, 2.5) if is_left else (3.5, 6.0)
    return _between(y, y_b, y_b + 0.75) & _between(x, *xr)


def mark_cells(coords, cells):
    """Cell markers 0..9 (forward_solve.py:134-144 == averaged_affine_ROM.py:101-112)."""
    X = coords[cells]                                   # (nc, 3, 2)
    mid = X.mean(axis=1)
    out = np.zeros(len(cells), dtype=np.int8)
    for q in range(1, NUM_SUBDOMAINS + 1):
        ok = _inside_subdomain(q, mid[:, 0], mid[:, 1])
        for a in range(3):
            ok &= _inside_subdomain(q, X[:, a, 0], X[:, a, 1])
        out[ok] = q
    return out


def boundary_facets(cells):
    """Edges that belong to exactly one cell, as an (nf, 2) vertex array."""
    e = np.concatenate([cells[:, [0, 1]], cells[:, [1, 2]], cells[:, [2, 0]]]).astype(np.int64)
    e.sort(axis=1)
    nv = int(cells.max()) + 1
    key = e[:, 0] * nv + e[:, 1]
    uniq, counts = np.unique(key, return_counts=True)
    b = uniq[counts == 1]
    return np.stack([b // nv, b % nv], axis=1)


def classify_facets(coords, facets):
    """Facet markers: 2 = root (all of the facet on y=0), 1 = exterior/Robin (no point of it on y=0),
    0 = neither (forward_solve.py:147-152)."""
    y0, y1 = coords[facets[:, 0], 1], coords[facets[:, 1], 1]
    ym = 0.5 * (y0 + y1)
    near = lambda y: np.abs(y) < DOLFIN_EPS
    mark = np.zeros(len(facets), dtype=np.int8)
    mark[~near(y0) & ~near(y1) & ~near(ym)] = 1
    mark[near(y0) & near(y1) & near(ym)] = 2
    return mark


def classify_facets_affine(coords, facets):
    """Exterior-facet markers of ``AffineROMFin`` exactly as rom/averaged_affine_ROM.py:116-138 applies them: sub-domain
    objects 1..9 in order (1 = SubFinBoundary of sub-fin 1, 5 = CenterFinBoundary, the others the plain SubFin boxes),
    then ``bottom`` = 10; a facet is marked iff both vertices and the midpoint satisfy ``inside``.  Markers 1..9 carry the
    Robin term (:156-162), 10 the root flux (:163), 0 nothing.  Differs from :func:`classify_facets` only on boundary facets
    that straddle x = 2.5 / 3.5, which no sub-domain box contains."""
    p0, p1 = coords[facets[:, 0]], coords[facets[:, 1]]
    pts = (p0, p1, 0.5 * (p0 + p1))
    near0 = lambda y: np.abs(y) < DOLFIN_EPS
    mark = np.zeros(len(facets), dtype=np.int8)
    for q in range(1, NUM_SUBDOMAINS + 1):
        ok = np.ones(len(facets), dtype=bool)
        for p in pts:
            ok &= _inside_subdomain(q, p[:, 0], p[:, 1])
            if q == 5:
                ok &= ~near0(p[:, 1])
        mark[ok] = q
    bottom = near0(pts[0][:, 1]) & near0(pts[1][:, 1]) & near0(pts[2][:, 1])
    mark[bottom] = 10
    return mark


@dataclass
class FinOperators:
    n: int
    n_cells: int
    coords: np.ndarray
    cells: np.ndarray
    cell_markers: np.ndarray
    row_ptr: np.ndarray            # (n+1,) int32
    col_idx: np.ndarray            # (nnz,) int32, sorted within each row
    vals: np.ndarray               # (10, nnz) float64: [Bi*M_Gamma, K_1 .. K_9]; M_Gamma with Fin's facet markers
    robin_affine: np.ndarray       # (nnz,) Bi*M_Gamma with AffineROMFin's own facet markers (averaged_affine_ROM.py:116-138)
    k_unmarked: np.ndarray         # (nnz,) stiffness of marker-0 cells (SURVEY Q-1; zero on conforming meshes)
    rhs: np.ndarray                # (n,)
    Ke: np.ndarray                 # (n_cells, 3, 3)
    cell_area: np.ndarray          # (n_cells,)
    subfin_area: np.ndarray        # (9,)
    B_obs: np.ndarray              # (9, n) dense
    C: np.ndarray                  # (n,)
    domain_measure: float
    boundary_dofs: np.ndarray      # dofs on exterior facets with no point on y=0 (external_obs support)
    _cache: dict = field(default_factory=dict, repr=False)

    @property
    def nnz(self):
        return int(self.col_idx.shape[0])

    def csr(self, values):
        import scipy.sparse as sp
        return sp.csr_matrix((values, self.col_idx, self.row_ptr), shape=(self.n, self.n))

    def affine_values(self, theta):
        """CSR values of A(theta) = sum_q theta_q K_q + Bi M (host helper, not the product path)."""
        theta = np.asarray(theta, dtype=np.float64)
        return self.robin_affine + theta @ self.vals[1:]

    def affine_terms(self):
        """(10, nnz) terms of the AFFINE model: its own Robin matrix, then K_1 .. K_9."""
        if np.array_equal(self.robin_affine, self.vals[0]):
            return self.vals
        return np.concatenate([self.robin_affine[None, :], self.vals[1:]], axis=0)

    def stiffness_values(self):
        """CSR values of the full stiffness K (all cells, marker 0 included)."""
        return self.vals[1:].sum(axis=0) + self.k_unmarked

    def mass_matrix(self):
        """Consistent P1 mass matrix ``assemble(inner(w, v) * dx)`` (forward_solve.py:172) as scipy CSR: per cell
        |e| / 12 * [[2, 1, 1], [1, 2, 1], [1, 1, 2]]."""
        import scipy.sparse as sp
        loc = (np.ones((3, 3)) + np.eye(3)) / 12.0
        r = np.repeat(self.cells, 3, axis=1).ravel()
        c = np.tile(self.cells, (1, 3)).ravel()
        v = (self.cell_area[:, None, None] * loc[None, :, :]).ravel()
        return sp.coo_matrix((v, (r, c)), shape=(self.n, self.n)).tocsr()

    def obs_csr(self, B=None):
        """(ptr, idx, val) CSR of an observation matrix (default: B_obs)."""
        B = self.B_obs if B is None else np.asarray(B, dtype=np.float64)
        rows, cols = np.nonzero(B)
        ptr = np.zeros(B.shape[0] + 1, dtype=np.int32)
        np.add.at(ptr, rows + 1, 1)
        return np.cumsum(ptr).astype(np.int32), cols.astype(np.int32), B[rows, cols].copy()


def build_operators(V) -> FinOperators:
    """Assemble all operators of the hot path for space ``V`` (cached on the space object)."""
    cached = getattr(V, "_fin_operators", None)
    if cached is not None:
        return cached
    mesh = V.mesh()
    coords, cells = mesh.coordinates(), mesh.cells()
    n, nc = coords.shape[0], cells.shape[0]
    c64 = cells.astype(np.int64)

    # --- element stiffness, SURVEY appendix A.2:  K_e[i,j] = (b_i b_j + c_i c_j) / (4|e|)
    X = coords[cells]
    x, y = X[:, :, 0], X[:, :, 1]
    b = np.stack([y[:, 1] - y[:, 2], y[:, 2] - y[:, 0], y[:, 0] - y[:, 1]], axis=1)
    c = np.stack([x[:, 2] - x[:, 1], x[:, 0] - x[:, 2], x[:, 1] - x[:, 0]], axis=1)
    det = b[:, 0] * c[:, 1] - b[:, 1] * c[:, 0]
    if np.any(det == 0.0):
        raise ValueError("degenerate (zero-area) cell in mesh")
    area = 0.5 * np.abs(det)
    Ke = (b[:, :, None] * b[:, None, :] + c[:, :, None] * c[:, None, :]) / (4.0 * area)[:, None, None]

    markers = mark_cells(coords, cells)

    # --- boundary terms
    facets = boundary_facets(cells)
    fmark = classify_facets(coords, facets)
    flen = np.linalg.norm(coords[facets[:, 0]] - coords[facets[:, 1]], axis=1)
    robin, root = facets[fmark == 1], facets[fmark == 2]
    lrob, lroot = flen[fmark == 1], flen[fmark == 2]
    rhs = np.zeros(n)
    np.add.at(rhs, root[:, 0], 0.5 * lroot)
    np.add.at(rhs, root[:, 1], 0.5 * lroot)

    # --- shared CSR pattern = cell connectivity (the Robin edges are cell edges, so no new entries)
    ri = np.repeat(c64, 3, axis=1).ravel()             # local (a, b) -> row = cells[:, a]
    ci = np.tile(c64, (1, 3)).ravel()
    key = ri * n + ci
    ukey, inv = np.unique(key, return_inverse=True)
    rows_u, cols_u = ukey // n, ukey % n
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(row_ptr, rows_u + 1, 1)
    row_ptr = np.cumsum(row_ptr)
    nnz = len(ukey)

    vals = np.zeros((NUM_SUBDOMAINS + 1, nnz))
    k0 = np.zeros(nnz)
    inv_e = inv.reshape(nc, 9)
    Kflat = Ke.reshape(nc, 9)
    for q in range(0, NUM_SUBDOMAINS + 1):
        sel = markers == q
        if not sel.any():
            continue
        tgt = vals[q] if q > 0 else k0
        np.add.at(tgt, inv_e[sel].ravel(), Kflat[sel].ravel())

    def pos(i, j):
        p = np.searchsorted(ukey, i * n + j)
        assert np.all(ukey[p] == i * n + j)
        return p

    def robin_values(edges, lengths):
        out = np.zeros(nnz)
        e0, e1 = edges[:, 0], edges[:, 1]
        np.add.at(out, pos(e0, e0), BIOT * lengths / 3.0)
        np.add.at(out, pos(e1, e1), BIOT * lengths / 3.0)
        np.add.at(out, pos(e0, e1), BIOT * lengths / 6.0)
        np.add.at(out, pos(e1, e0), BIOT * lengths / 6.0)
        return out

    vals[0] = robin_values(robin, lrob)
    fmark_aff = classify_facets_affine(coords, facets)
    sel_aff = (fmark_aff >= 1) & (fmark_aff <= 9)
    assert np.array_equal(fmark_aff == 10, fmark == 2)
    robin_affine = vals[0] if np.array_equal(sel_aff, fmark == 1) else robin_values(facets[sel_aff], flen[sel_aff])

    # --- observation / averaging operators, appendix A.2
    subfin_area = np.array([area[markers == q].sum() for q in range(1, NUM_SUBDOMAINS + 1)])
    B_obs = np.zeros((NUM_SUBDOMAINS, n))
    for q in range(1, NUM_SUBDOMAINS + 1):
        sel = markers == q
        if subfin_area[q - 1] > 0:
            np.add.at(B_obs[q - 1], c64[sel].ravel(), np.repeat(area[sel] / 3.0, 3))
            B_obs[q - 1] /= subfin_area[q - 1]
    C = np.zeros(n)
    np.add.at(C, c64.ravel(), np.repeat(area / 3.0, 3))
    domain_measure = float(area.sum())
    C /= domain_measure

    ops = FinOperators(
        n=n, n_cells=nc, coords=coords, cells=cells, cell_markers=markers,
        row_ptr=row_ptr.astype(np.int32), col_idx=cols_u.astype(np.int32), vals=vals, robin_affine=robin_affine,
        k_unmarked=k0,
        rhs=rhs, Ke=Ke, cell_area=area, subfin_area=subfin_area, B_obs=B_obs, C=C,
        domain_measure=domain_measure, boundary_dofs=np.unique(robin.ravel()))
    try:
        V._fin_operators = ops
    except AttributeError:
        pass
    return ops


def five_to_nine(k5):
    """[k1..k5] -> [k1,k2,k3,k4,k5,k4,k3,k2,k1]: forward_solve_petsc.py:243-260 expressed in the sub-fin
    numbering of forward_solve.py:125-133 (left fins bottom-up are 1..4, right fins top-down are 6..9)."""
    k5 = np.asarray(k5, dtype=np.float64)
    return np.concatenate([k5[..., :5], k5[..., 3::-1]], axis=-1)


def nine_param_nodal(coords, theta):
    """Nodal values of ``interpolate(SubfinValExpr(k_s), V)`` (forward_solve.py:61-91, 482-486).
    ``theta``: (9,) or (N, 9) -> (n,) or (N, n)."""
    theta = np.asarray(theta, dtype=np.float64)
    x, y = coords[:, 0], coords[:, 1]
    bands = [_between(y, 0.75, 1.0), _between(y, 1.75, 2.0), _between(y, 2.75, 3.0), _between(y, 3.75, 4.0)]
    sel = np.full(len(coords), -1, dtype=np.int64)      # index into theta, -1 -> value 0
    centre = _between(x, 2.5, 3.5)
    leftside = ~centre & (x <= 2.5)
    rightside = ~centre & ~leftside
    for bi in (3, 2, 1, 0):                               # first matching band wins (if/elif chain)
        sel[leftside & bands[bi]] = bi                    # k1..k4
        sel[rightside & bands[bi]] = 8 - bi               # k9..k6
    sel[centre] = 4
    padded = np.concatenate([theta, np.zeros(theta.shape[:-1] + (1,))], axis=-1)
    return padded[..., sel]
