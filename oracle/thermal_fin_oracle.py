"""CPU ORACLE -- test infrastructure, NOT product code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs
may import this module.  The product package (``bayesianinferencedl_b200``) never does.

PARITY UNPINNED.  The reference (sheroze1123/BayesianInferenceDL) does all arithmetic of this path
inside FEniCS 2018.1 / mshr / PETSc / LAPACK, none of which is installed or installable here, and it
ships neither its mesh nor a single forward-solve output.  This file
is a restatement, in numpy + scipy, of what those libraries compute at the reference's call sites; each
function cites the reference lines it follows.  What CAN be pinned from the reference's shipped data
(``data/B_obs.txt`` row sums, ``B_obs @ phi``, shapes/conditioning of the bases) is pinned in
``tests/golden`` (see ``tests/golden/make_golden.py``).

Deliberately written independently of ``bayesianinferencedl_b200/assembly.py`` (per-element gradient
formulation, COO accumulation through scipy, loops over sub-domains) so that the two check each other.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from scipy import spatial

DOLFIN_EPS = 3.0e-16


def between(x, lo, hi):
    """dolfin ``between(x, (lo, hi))``: lo - eps <= x <= hi + eps."""
    return (x >= lo - DOLFIN_EPS) and (x <= hi + DOLFIN_EPS)


def near(x, a):
    """dolfin ``near(x, a)``: |x - a| < eps."""
    return abs(x - a) < DOLFIN_EPS


# ------------------------------------------------------------------------------------------------
# sub-domain predicates: fom/forward_solve.py:5-37 (duplicated in rom/averaged_affine_ROM.py:18-50)
# ------------------------------------------------------------------------------------------------
def subfin_inside(p, y_b, is_left):
    """SubFin.inside, forward_solve.py:11-15."""
    if is_left:
        return between(p[1], y_b, y_b + 0.75) and between(p[0], 0.0, 2.5)
    return between(p[1], y_b, y_b + 0.75) and between(p[0], 3.5, 6.0)


def centerfin_inside(p):
    """CenterFin.inside, forward_solve.py:31-33."""
    return between(p[0], 2.5, 3.5)


# order of fin1..fin9, forward_solve.py:125-133
SUBDOMAINS = [(0.75, True), (1.75, True), (2.75, True), (3.75, True), None,
              (3.75, False), (2.75, False), (1.75, False), (0.75, False)]


def inside(q, p):
    d = SUBDOMAINS[q - 1]
    return centerfin_inside(p) if d is None else subfin_inside(p, d[0], d[1])


def mark_cells(coords, cells):
    """``finq.mark(domains, q)`` for q = 1..9 in order (forward_solve.py:134-144): a cell gets q iff all
    vertices and the midpoint are inside; later ids overwrite; unmatched cells keep 0."""
    out = np.zeros(len(cells), dtype=np.int64)
    for e, tri in enumerate(cells):
        pts = [coords[v] for v in tri] + [coords[tri].mean(axis=0)]
        for q in range(1, 10):
            if all(inside(q, p) for p in pts):
                out[e] = q
    return out


def exterior_facets(cells):
    """Facets (edges) with exactly one adjacent cell -- what dolfin's ``on_boundary`` / ``ds`` see."""
    count = {}
    for tri in cells:
        for a, b in ((0, 1), (1, 2), (2, 0)):
            k = (min(tri[a], tri[b]), max(tri[a], tri[b]))
            count[k] = count.get(k, 0) + 1
    return sorted(k for k, c in count.items() if c == 1)


def mark_facets(coords, cells):
    """forward_solve.py:147-152: exterior = "!near(x[1],0) && on_boundary" -> 1, bottom = "near(x[1],0)
    && on_boundary" -> 2, both tested on the two vertices and the midpoint.  Returns (robin, root) lists.
    (``AffineROMFin`` marks its facets differently: mark_facets_affine.)"""
    robin, root = [], []
    for (a, b) in exterior_facets(cells):
        ys = [coords[a][1], coords[b][1], 0.5 * (coords[a][1] + coords[b][1])]
        if all(not near(y, 0.0) for y in ys):
            robin.append((a, b))
        elif all(near(y, 0.0) for y in ys):
            root.append((a, b))
    return robin, root


def mark_facets_affine(coords, cells):
    """Facet markers of ``AffineROMFin`` as written, rom/averaged_affine_ROM.py:116-138: nine sub-domain objects mark the
    facet function in order 1..9, then ``bottom`` marks 10; ``ds(i)`` only ever integrates over EXTERIOR facets, so only
    those are listed.  dolfin's ``SubDomain.mark`` marks a facet iff all its vertices AND its midpoint satisfy ``inside``:
      1      SubFinBoundary([0.75, True])   on_boundary and inside the (over-tall) box of sub-fin 1      (:30-42, :118)
      2-4, 6-9  SubFin(...)                 inside the box of the sub-fin -- note: the plain SubFin, not ...Boundary (:119-126)
      5      CenterFinBoundary()            on_boundary and 2.5 <= x <= 3.5 and not near(y, 0)            (:48-50, :122)
      10     bottom                         near(y, 0) and on_boundary                                     (:117, :138)
    Returns (robin, root): exterior facets with a marker in 1..9 (they carry the Bi term, :156-162) and with marker 10.
    This differs from ``Fin``'s single exterior marker (mark_facets) exactly on boundary facets that straddle x = 2.5 or
    x = 3.5 (e.g. on the top edge y = 4 of a non-conforming mesh): no sub-domain box contains them, they keep marker 0 and
    the affine model has NO Robin term there."""
    robin, root = [], []
    for (a, b) in exterior_facets(cells):
        pts = [coords[a], coords[b], 0.5 * (coords[a] + coords[b])]
        marker = 0
        for q in range(1, 10):
            if q == 5:
                ok = all(centerfin_inside(p) and not near(p[1], 0.0) for p in pts)
            else:
                ok = all(inside(q, p) for p in pts)        # on_boundary is true for every point of an exterior facet
            if ok:
                marker = q
        if all(near(p[1], 0.0) for p in pts):
            marker = 10
        if 1 <= marker <= 9:
            robin.append((a, b))
        elif marker == 10:
            root.append((a, b))
    return robin, root


# ------------------------------------------------------------------------------------------------
# P1 assembly (what dolfin.assemble produces for the forms of forward_solve.py:160-163)
# ------------------------------------------------------------------------------------------------
class FinOracle:
    """Restatement of ``Fin`` (fom/forward_solve.py:93-265) + ``AffineROMFin`` forward map
    (rom/averaged_affine_ROM.py:52-333) on a given triangle mesh."""

    Bi = 0.1   # forward_solve.py:112

    def __init__(self, coords, cells):
        self.coords = np.asarray(coords, dtype=np.float64)
        self.cells = np.asarray(cells, dtype=np.int64)
        self.n = len(self.coords)
        nc = len(self.cells)
        # gradients of the barycentric basis: G = inv([[1,x0,y0],[1,x1,y1],[1,x2,y2]])[1:, :]
        P = np.ones((nc, 3, 3))
        P[:, :, 1:] = self.coords[self.cells]
        Pinv = np.linalg.inv(P)
        grads = Pinv[:, 1:, :]                          # (nc, 2, 3): d phi_a / d(x,y)
        self.area = 0.5 * np.abs(np.linalg.det(P))
        self.Ke = np.einsum("eda,edb->eab", grads, grads) * self.area[:, None, None]
        self.markers = mark_cells(self.coords, self.cells)
        self.robin, self.root = mark_facets(self.coords, self.cells)

        # Bi * int_{exterior} w v ds   (forward_solve.py:161)
        rows, cols, vals = [], [], []
        for a, b in self.robin:
            L = np.linalg.norm(self.coords[a] - self.coords[b])
            rows += [a, b, a, b]
            cols += [a, b, b, a]
            vals += [L / 3.0, L / 3.0, L / 6.0, L / 6.0]
        self.M_robin = sp.coo_matrix((vals, (rows, cols)), shape=(self.n, self.n)).tocsr()
        # the affine model's own facet markers (averaged_affine_ROM.py:116-138); identical to M_robin unless a boundary
        # facet straddles x = 2.5 / 3.5
        self.robin_affine, root_affine = mark_facets_affine(self.coords, self.cells)
        assert sorted(root_affine) == sorted(self.root)
        rows, cols, vals = [], [], []
        for a, b in self.robin_affine:
            L = np.linalg.norm(self.coords[a] - self.coords[b])
            rows += [a, b, a, b]
            cols += [a, b, b, a]
            vals += [L / 3.0, L / 3.0, L / 6.0, L / 6.0]
        self.M_robin_affine = sp.coo_matrix((vals, (rows, cols)), shape=(self.n, self.n)).tocsr()

        # B = assemble(v * ds(2))   (forward_solve.py:162-163)
        self.B = np.zeros(self.n)
        for a, b in self.root:
            L = np.linalg.norm(self.coords[a] - self.coords[b])
            self.B[a] += 0.5 * L
            self.B[b] += 0.5 * L

        # K_q = assemble(inner(grad w, grad v) * dx(q))   (averaged_affine_ROM.py:156-162, 217)
        self.K_q = [self._stiffness(np.where(self.markers == q, 1.0, 0.0)) for q in range(1, 10)]

        # sub-fin areas and B_obs   (forward_solve.py:205-213, 488-511)
        self.fin_area = np.array([self.area[self.markers == q].sum() for q in range(1, 10)])
        self.B_obs = np.zeros((9, self.n))
        for e, tri in enumerate(self.cells):
            q = self.markers[e]
            if q > 0:
                for v in tri:
                    self.B_obs[q - 1, v] += self.area[e] / 3.0
        self.B_obs /= self.fin_area[:, None]
        self.n_obs = 9

        # averaging operator C   (forward_solve.py:396-406)
        self.domain_measure = self.area.sum()
        self.C = np.zeros(self.n)
        for e, tri in enumerate(self.cells):
            for v in tri:
                self.C[v] += self.area[e] / 3.0
        self.C /= self.domain_measure

    def _stiffness(self, cell_coeff):
        """sum_e cell_coeff[e] * K_e as CSR."""
        r = np.repeat(self.cells, 3, axis=1).ravel()
        c = np.tile(self.cells, (1, 3)).ravel()
        v = (self.Ke * cell_coeff[:, None, None]).ravel()
        return sp.coo_matrix((v, (r, c)), shape=(self.n, self.n)).tocsr()

    # ---------------- full-order models ----------------
    def matrix_nodal(self, k):
        """A(k) of ``Fin._F`` (forward_solve.py:160-161) for a P1 conductivity with nodal values k:
        int k grad w . grad v over a cell = mean(k at the 3 vertices) * K_e exactly."""
        kbar = np.asarray(k, dtype=np.float64)[self.cells].mean(axis=1)
        return (self._stiffness(kbar) + self.Bi * self.M_robin).tocsc()

    def matrix_nodal_exp(self, k):
        """A of ``forward_solve_exp.py:160-161``: ``exp(k) grad w . grad v``.  The form compiler integrates it with the
        rule UFL/FIAT select: degree(exp(P1)) is estimated as 1 + 2 = 3 and FIAT's degree-3 triangle scheme is the
        6-point Strang-Fix rule (equal weights, barycentric permutations of 0.659..., 0.231..., 0.109...), so the cell
        coefficient is the mean of exp(k) over those six points.  (dolfin/FFC 2018.1 are not installable here: the
        rule is restated from FIAT's quadrature_schemes, parity unpinned like the rest of this file.)"""
        kv = np.asarray(k, dtype=np.float64)[self.cells]                       # (n_cells, 3)
        a, b, c = 0.659027622374092, 0.231933368553031, 0.109039009072877
        pts = np.array([[a, b, c], [a, c, b], [b, a, c], [b, c, a], [c, a, b], [c, b, a]])
        coeff = np.exp(kv @ pts.T).mean(axis=1)
        return (self._stiffness(coeff) + self.Bi * self.M_robin).tocsc()

    def forward_exp(self, k):
        """``Fin.forward`` of forward_solve_exp.py:252-275."""
        return spla.splu(self.matrix_nodal_exp(k)).solve(self.B)

    def grad_form_exp(self, k, z, v):
        """assemble(k_hat * exp(k) * inner(grad z, grad v) * dx) (forward_solve_exp.py:299, 328).  Estimated degree
        1 + 3 = 4 -> FIAT's degree-4 triangle scheme: 6-point Strang-Fix rule, orbits (0.8168..., 0.0915..., 0.0915...)
        with weight 0.10995... and (0.1081..., 0.4459..., 0.4459...) with weight 0.22338... (normalised to area 1)."""
        a1, b1, w1 = 0.816847572980459, 0.091576213509771, 0.109951743655322
        a2, b2, w2 = 0.108103018168070, 0.445948490915965, 0.223381589678011
        pts = np.array([[a1, b1, b1], [b1, a1, b1], [b1, b1, a1], [a2, b2, b2], [b2, a2, b2], [b2, b2, a2]])
        wts = np.array([w1, w1, w1, w2, w2, w2])
        kv = np.asarray(k, dtype=np.float64)[self.cells]                      # (n_cells, 3)
        ek = np.exp(kv @ pts.T)                                                # exp(k) at the 6 points
        vert_w = (ek * wts) @ pts                                              # (n_cells, 3): int lambda_a exp(k) / |e|
        ge = np.einsum("ea,eab,eb->e", z[self.cells], self.Ke, v[self.cells])  # |e| grad z . grad v
        out = np.zeros(self.n)
        for a in range(3):
            np.add.at(out, self.cells[:, a], ge * vert_w[:, a])
        return out

    def gradient_exp(self, k, data):
        """``Fin.gradient`` of forward_solve_exp.py:277-310."""
        A = self.matrix_nodal_exp(k)
        lu = spla.splu(A)
        z = lu.solve(self.B)
        adj_rhs = -np.dot((self.B_obs @ z - data).T, self.B_obs)
        return self.grad_form_exp(k, z, lu.solve(adj_rhs))

    def sensitivity_exp(self, k):
        """``Fin.sensitivity`` of forward_solve_exp.py:312-342."""
        lu = spla.splu(self.matrix_nodal_exp(k))
        z = lu.solve(self.B)
        V = lu.solve(np.ascontiguousarray(-self.B_obs.T))
        return np.stack([self.grad_form_exp(k, z, V[:, o]) for o in range(self.B_obs.shape[0])])

    def forward(self, k):
        """``Fin.forward(k)`` (forward_solve.py:270-291): dolfin ``solve`` = sparse direct LU."""
        return spla.splu(self.matrix_nodal(k)).solve(self.B)

    def matrix_affine(self, theta):
        """A(theta) of ``AffineROMFin._F`` (averaged_affine_ROM.py:156-162)."""
        A = self.Bi * self.M_robin_affine
        for q in range(9):
            A = A + float(theta[q]) * self.K_q[q]
        return A.tocsc()

    def forward_nine_param(self, theta):
        """``forward_nine_param(k_s)`` (generate_reduced_basis_nine_param.py:178-183)."""
        return spla.splu(self.matrix_affine(theta)).solve(self.B)

    def forward_affine(self, k):
        """``AffineROMFin.forward(k)`` (averaged_affine_ROM.py:237-258)."""
        return self.forward_nine_param(self.subfin_avg_op(k))

    def forward_five_param_affine(self, k5):
        return self.forward_nine_param(five_param_to_nine(k5))

    # ---------------- observation ----------------
    def subfin_avg_op(self, k):
        """forward_solve.py:466-480 / averaged_affine_ROM.py:404-418: assemble(k*dx(q))/area_q."""
        return self.B_obs @ np.asarray(k, dtype=np.float64)

    def qoi_operator(self, w):
        """forward_solve.py:408-412."""
        return np.dot(self.B_obs, w)

    def nine_param_to_function(self, theta):
        """interpolate(SubfinValExpr(k_s)) (forward_solve.py:61-91, 482-486), evaluated per vertex."""
        k1, k2, k3, k4, k5, k6, k7, k8, k9 = [float(t) for t in theta]
        out = np.zeros(self.n)
        for i, (x, y) in enumerate(self.coords):
            if between(x, 2.5, 3.5):
                v = k5
            elif x <= 2.5:
                v = (k1 if between(y, 0.75, 1.0) else k2 if between(y, 1.75, 2.0) else
                     k3 if between(y, 2.75, 3.0) else k4 if between(y, 3.75, 4.0) else 0.0)
            else:
                v = (k9 if between(y, 0.75, 1.0) else k8 if between(y, 1.75, 2.0) else
                     k7 if between(y, 2.75, 3.0) else k6 if between(y, 3.75, 4.0) else 0.0)
            out[i] = v
        return out

    # ---------------- reduced-order model (LSPG) ----------------
    def forward_nine_param_reduced(self, theta, phi):
        """averaged_affine_ROM.py:278-310, literally: psi = A phi; A_r = psi^T psi; B_r = psi^T B;
        w_r = np.linalg.solve(A_r, B_r)."""
        A = self.matrix_affine(theta).tocsr()
        psi = A @ phi
        A_r = psi.T @ psi
        B_r = psi.T @ self.B
        return np.linalg.solve(A_r, B_r)

    def forward_reduced(self, k, phi):
        """averaged_affine_ROM.py:260-276."""
        return self.forward_nine_param_reduced(self.subfin_avg_op(k), phi)

    def qoi_reduced(self, w_r, phi):
        """averaged_affine_ROM.py:212, 323-333: (B_obs phi) w_r."""
        return np.dot(np.dot(self.B_obs, phi), w_r)

    def r_fwd_no_full(self, k, phi):
        """forward_solve.py:421-464 (nodal-conductivity LSPG with the legacy scalar QoI C)."""
        A = self.matrix_nodal(k).tocsr()
        psi = A @ phi
        A_r = psi.T @ (A @ phi)
        B_r = psi.T @ self.B
        C_r = self.C @ phi
        x_r = np.linalg.solve(A_r, B_r)
        return A_r, B_r, C_r, x_r, float(C_r @ x_r)


    # ---------------- adjoint gradients ----------------
    def grad_form(self, z, v):
        """assemble(k_hat * inner(grad z, grad v) * dx) for the P1 test function k_hat (forward_solve.py:313-314,
        180): entry i = sum over cells containing vertex i of (|e|/3) grad z . grad v = (1/3) z_e^T K_e v_e."""
        ge = np.einsum("ea,eab,eb->e", z[self.cells], self.Ke, v[self.cells]) / 3.0
        out = np.zeros(self.n)
        for a in range(3):
            np.add.at(out, self.cells[:, a], ge)
        return out

    def gradient(self, k, data):
        """``Fin.gradient(k, data)`` (forward_solve.py:293-322): forward solve, adjoint solve with
        rhs = -(B_obs z - data)^T B_obs (the reference uses a DENSE np.linalg.solve, :310), gradient form."""
        A = self.matrix_nodal(k)
        lu = spla.splu(A)
        z = lu.solve(self.B)
        pred = self.B_obs @ z
        adj_rhs = -np.dot((pred - data).T, self.B_obs)
        v = np.linalg.solve(A.toarray(), adj_rhs) if self.n <= 2000 else lu.solve(adj_rhs)
        return self.grad_form(z, v)

    def sensitivity(self, k):
        """``Fin.sensitivity(k)`` (forward_solve.py:324-342): Jacobian of the observables w.r.t. the nodal
        conductivity, (n_obs, n): adjoint solves with rhs = -B_obs^T, then grad_assembled @ v."""
        A = self.matrix_nodal(k)
        lu = spla.splu(A)
        z = lu.solve(self.B)
        V = lu.solve(np.ascontiguousarray(-self.B_obs.T))
        return np.stack([self.grad_form(z, V[:, o]) for o in range(self.B_obs.shape[0])])

    def grad_reduced(self, k, data, phi):
        """``AffineROMFin.grad_reduced(k)`` (averaged_affine_ROM.py:335-356), literally, with
        dA_dsigmak_phi[q] = K_q phi (:215-220) and dsigma_dk = B_obs (:210).  Returns (dJ_dk, J)."""
        theta = self.subfin_avg_op(k)
        A = self.matrix_affine(theta).tocsr()
        psi = A @ phi
        A_r = psi.T @ psi
        w_r = np.linalg.solve(A_r, psi.T @ self.B)
        B_obs_phi = self.B_obs @ phi
        reduced_fwd_obs = B_obs_phi @ w_r
        reduced_adj_rhs = B_obs_phi.T @ (data - reduced_fwd_obs)
        v_r = np.linalg.solve(A_r.T, reduced_adj_rhs)
        psi_v_r = psi @ v_r
        A_phi_w_r = np.stack([(Kq @ phi) @ w_r for Kq in self.K_q]).T      # (n, 9)
        g_theta = psi_v_r @ A_phi_w_r                                       # (9,)  = dJ / d theta
        dJ_dk = g_theta @ self.B_obs                                        # psi_v_r^T (A_phi_w_r dsigma_dk)
        J = 0.5 * np.linalg.norm(data - reduced_fwd_obs) ** 2
        return dJ_dk, J, g_theta

    def reduced_forward(self, A, B, C, psi, phi):
        """``Fin.reduced_forward`` (forward_solve.py:421-452), literally (dense A)."""
        A_r = np.dot(psi.T, np.dot(A, phi))
        B_r = np.dot(psi.T, B)
        C_r = np.dot(C, phi)
        x_r = np.linalg.solve(A_r, B_r)
        return A_r, B_r, C_r, x_r, np.dot(C_r, x_r)


def five_param_to_nine(k5):
    """forward_solve_petsc.py:243-260: k1..k4 are the y-bands 0.75/1.75/2.75/3.75 on BOTH sides, k5 the
    post.  In the numbering of forward_solve.py:125-133 (right fins counted top-down) this is
    [k1,k2,k3,k4,k5,k4,k3,k2,k1]."""
    k1, k2, k3, k4, k5 = [float(v) for v in k5]
    return np.array([k1, k2, k3, k4, k5, k4, k3, k2, k1])


def make_cov_chol(coords, kern_type="m52", length=1.6):
    """bayesian_inference/gaussian_field.py:9-31 on the dof coordinates."""
    d = spatial.distance.squareform(spatial.distance.pdist(np.asarray(coords)))
    if kern_type == "sq_exp":
        cov = np.exp(-d ** 2 / (2 * length ** 2)) + np.eye(len(d)) * 1e-5
    elif kern_type == "m52":
        t = np.sqrt(5) * d / length
        cov = (1 + t + t * t / 3) * np.exp(-t)
    else:
        t = np.sqrt(3) * d / length
        cov = (1 + t) * np.exp(-t)
    return sla.cholesky(cov)


def sample_field(chol, z):
    """deep_learning/generate_fin_dataset.py:87-88: nodal_vals = exp(0.5 * chol.T @ norm)."""
    return np.exp(0.5 * chol.T @ z)


def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon et al., SC'11; Random123): counter = 4 uint64 arrays holding 32-bit words, key = 2 ints.
    Pinned against the Random123 known-answer vectors in tests/test_oracle.py."""
    m32 = np.uint64(0xFFFFFFFF)
    c = [np.asarray(x, dtype=np.uint64) for x in counter]
    k0, k1 = np.uint64(key[0] & 0xFFFFFFFF), np.uint64(key[1] & 0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c[0]
        p1 = np.uint64(0xCD9E8D57) * c[2]
        n0 = (p1 >> np.uint64(32)) ^ c[1] ^ k0
        n2 = (p0 >> np.uint64(32)) ^ c[3] ^ k1
        c = [n0, p1 & m32, n2, p0 & m32]
        k0 = (k0 + np.uint64(0x9E3779B9)) & m32
        k1 = (k1 + np.uint64(0xBB67AE85)) & m32
    return c


def philox_uniforms(seed, rows, pairs, subsequence=0):
    """Two 53-bit uniforms per (row, pair): counter (pair, row_lo, row_hi, subsequence), key ``seed``.
    ``rows`` / ``pairs``: integer arrays (broadcast together).  u1 in (0, 1], u2 in [0, 1)."""
    rows, pairs = np.broadcast_arrays(np.asarray(rows, dtype=np.uint64), np.asarray(pairs, dtype=np.uint64))
    m32 = np.uint64(0xFFFFFFFF)
    c = philox4x32_10([pairs & m32, rows & m32, rows >> np.uint64(32),
                       np.full(rows.shape, subsequence & 0xFFFFFFFF, np.uint64)], (seed, seed >> 32))
    u1 = ((((c[0] << np.uint64(32)) | c[1]) >> np.uint64(11)).astype(np.float64) + 1.0) / 9007199254740992.0
    u2 = (((c[2] << np.uint64(32)) | c[3]) >> np.uint64(11)).astype(np.float64) / 9007199254740992.0
    return u1, u2


def philox_normals(seed, n_rows, n, first_row=0, subsequence=0):
    """Restatement of the product's device generator (csrc/field.cuh, F3) -- NOT part of the reference, which draws
    ``np.random.randn`` unseeded (generate_fin_dataset.py:87): one Philox4x32-10 stream per row, entries (2p, 2p+1) of
    global row g = Box-Muller of counter (p, g, subsequence) under key ``seed``.  Returns (n_rows, n) normals."""
    ppr = (n + 1) // 2
    g = (first_row + np.arange(n_rows))[:, None]
    u1, u2 = philox_uniforms(seed, g, np.arange(ppr)[None, :], subsequence)
    r = np.sqrt(-2.0 * np.log(u1))
    out = np.empty((n_rows, 2 * ppr))
    out[:, 0::2] = r * np.cos(2.0 * np.pi * u2)
    out[:, 1::2] = r * np.sin(2.0 * np.pi * u2)
    return out[:, :n]


def pcn_chains(qoi_fn, chol, data, sigma, seed, n_chains, n_steps, beta, first_chain=0, z0=None):
    """Restatement of the product's many-chain pCN driver (csrc/chains.cuh; not in the reference, whose samplers are
    PyMC3 / MUQ): misfit 0.5 ||qoi_fn(k) - data||^2 / sigma^2 (pymc_func_bayes_inverse.py:76, 201) with
    k = exp(0.5 chol^T z) (generate_fin_dataset.py:88); proposal z' = sqrt(1 - beta^2) z + beta xi; accept when
    log u < Phi - Phi'.  xi / u of chain g at step t are the Philox draws (seed, g, t)."""
    n = chol.shape[0]
    z = philox_normals(seed, n_chains, n, first_chain, 0) if z0 is None else np.array(z0, dtype=np.float64)
    misfit = lambda q: 0.5 * np.sum((q - data) ** 2) / sigma ** 2
    q = np.stack([qoi_fn(sample_field(chol, z[c])) for c in range(n_chains)])
    phi = np.array([misfit(q[c]) for c in range(n_chains)])
    acc = np.zeros(n_chains, dtype=np.int64)
    q_sum, q_sq = np.zeros_like(q), np.zeros_like(q)
    k_sum = np.zeros((n_chains, n))
    for t in range(1, n_steps + 1):
        xi = philox_normals(seed, n_chains, n, first_chain, t)
        u, _ = philox_uniforms(seed, first_chain + np.arange(n_chains), 0xFFFFFFFF, t)
        zp = np.sqrt(1.0 - beta * beta) * z + beta * xi
        for c in range(n_chains):
            qp = qoi_fn(sample_field(chol, zp[c]))
            pp = misfit(qp)
            if np.log(u[c]) < phi[c] - pp:
                z[c], q[c], phi[c] = zp[c], qp, pp
                acc[c] += 1
        q_sum += q
        q_sq += q * q
        k_sum += np.exp(0.5 * (z @ chol))
    return {"z": z, "qoi": q, "misfit": phi, "accepted": acc, "qoi_sum": q_sum, "qoi_sq": q_sq, "k_sum": k_sum}


def pod_basis(oracle: FinOracle, n_snapshots=200, basis_size=81, seed=0, lo=0.1, hi=3.5):
    """POD recipe of rom/generate_reduced_basis_nine_param.py:296-318 (commented script that produced
    data/basis_nine_param.txt): snapshots of forward_nine_param at k ~ U(0.1, 3.5)^9, eigenvectors of
    Y Y^T, UNNORMALISED modes U_i = sum_s v[s,i] Y[s,:].  (eigh + descending sort instead of eig.)"""
    rng = np.random.default_rng(seed)
    Y = np.stack([oracle.forward_nine_param(rng.uniform(lo, hi, 9)) for _ in range(n_snapshots)])
    e, v = np.linalg.eigh(Y @ Y.T)
    order = np.argsort(e)[::-1][:basis_size]
    return (v[:, order].T @ Y).T
