"""CPU tests of the oracle: pinned against everything the reference's shipped data pins, against its own
committed golden outputs, and against the analytic invariants of the problem."""
import os

import numpy as np
import pytest

from conftest import relerr

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(GOLD, "reference_data.npz"))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "oracle_m1.npz"))


def _shipped_B_obs(ref):
    B = np.zeros(tuple(ref["B_obs_shape"]))
    B[ref["B_obs_rows"], ref["B_obs_cols"]] = ref["B_obs_vals"]
    return B


def test_reference_B_obs_properties(ref, oracle_m3):
    """data/B_obs.txt (observation_operator on the unshipped mshr mesh): 9 rows, each a partition of unity
    average -> row sums 1, non-negative; the oracle's B_obs has the same properties."""
    B = _shipped_B_obs(ref)
    assert B.shape == (9, 1446)
    assert np.allclose(B.sum(axis=1), 1.0, atol=1e-12)
    assert B.min() >= 0
    Bo = oracle_m3.B_obs
    assert Bo.shape == (9, oracle_m3.n)
    assert np.allclose(Bo.sum(axis=1), 1.0, atol=1e-13)
    assert Bo.min() >= 0
    # a dof belongs to at most two sub-domains (post / sub-fin interface) in both
    assert (B > 0).sum(axis=0).max() <= 2 and (Bo > 0).sum(axis=0).max() <= 2


def test_reference_B_obs_phi_identity(ref):
    """averaged_affine_ROM.py:212: B_obs_phi = np.dot(B_obs, phi) on the shipped files (first 8 columns)."""
    B = _shipped_B_obs(ref)
    for name in ("five", "nine"):
        assert tuple(ref[f"phi_{name}_shape"]) == (1446, 81)
        got = np.dot(B, ref[f"phi_{name}_head"])
        assert np.allclose(got, ref[f"B_obs_phi_{name}_head"], rtol=1e-13, atol=1e-16)


def test_oracle_golden_regression(gold, oracle_m1):
    o = oracle_m1
    for s, t in enumerate(gold["theta"]):
        assert relerr(o.qoi_operator(o.forward_nine_param(t)), gold["qoi_affine"][s]) < 1e-12
        wr = o.forward_nine_param_reduced(t, gold["phi"])
        assert relerr(o.qoi_reduced(wr, gold["phi"]), gold["qoi_rom"][s]) < 1e-9
    assert np.allclose(o.forward_nine_param(gold["theta"][0]), gold["w_affine0"], rtol=1e-12, atol=1e-15)
    for s, k in enumerate(gold["k5"]):
        assert relerr(o.qoi_operator(o.forward_five_param_affine(k)), gold["qoi_five"][s]) < 1e-12
    for s, k in enumerate(gold["k_nodal"]):
        assert relerr(o.qoi_operator(o.forward(k)), gold["qoi_nodal"][s]) < 1e-12
        assert relerr(o.subfin_avg_op(k), gold["theta_of_k"][s]) < 1e-13
    assert np.array_equal(o.nine_param_to_function(gold["theta"][0]), gold["nine_to_fn"])


def test_invariants(oracle_m2):
    o = oracle_m2
    ones = np.ones(o.n)
    for K in o.K_q:                                   # stiffness annihilates constants, symmetric
        assert np.abs(K @ ones).max() < 1e-12
        assert abs(K - K.T).max() < 1e-14
    assert abs(o.B.sum() - 1.0) < 1e-14               # |Gamma_root| = 1
    rng = np.random.default_rng(0)
    theta = rng.uniform(0.1, 3.5, 9)
    w = o.forward_nine_param(theta)
    assert abs(o.Bi * ones @ (o.M_robin @ w) - 1.0) < 1e-11     # energy balance
    assert w.min() > 0
    # five-parameter problem is mirror symmetric
    q = o.qoi_operator(o.forward_five_param_affine([0.4, 0.6, 0.8, 1.0, 0.2]))
    assert np.allclose(q[:4], q[:4:-1], rtol=1e-10)
    # markers: conforming mesh -> no unmarked cell; areas of the nine rectangles
    assert o.markers.min() == 1
    assert np.allclose(o.fin_area, [0.625] * 4 + [4.0] + [0.625] * 4)
    # C averages: sums to one
    assert abs(o.C.sum() - 1.0) < 1e-13 and abs(o.domain_measure - 9.0) < 1e-12


def test_f1_vs_f2(oracle_m2):
    """SURVEY Q-2: nodal interpolation of a piecewise constant field (F1) differs from the affine model (F2), but
    a CONSTANT field gives identical operators."""
    o = oracle_m2
    A1 = o.matrix_nodal(np.full(o.n, 1.7))
    A2 = o.matrix_affine(np.full(9, 1.7))
    assert abs(A1 - A2).max() < 1e-12
    theta = np.array([0.5, 1.0, 1.5, 2.0, 2.5, 3.0, 0.3, 0.8, 1.2])
    w1 = o.forward(o.nine_param_to_function(theta))
    w2 = o.forward_nine_param(theta)
    assert 1e-4 < np.abs(w1 - w2).max()


def test_rom_reproduces_snapshots(oracle_m1):
    """LSPG with a basis containing the exact solution returns it (residual zero)."""
    o = oracle_m1
    rng = np.random.default_rng(1)
    theta = rng.uniform(0.5, 2.0, 9)
    w = o.forward_nine_param(theta)
    phi = np.column_stack([w, rng.standard_normal((o.n, 3))])
    wr = o.forward_nine_param_reduced(theta, phi)
    assert np.allclose(phi @ wr, w, rtol=1e-8, atol=1e-10)


def test_cov_chol(oracle_m1):
    from oracle.thermal_fin_oracle import make_cov_chol, sample_field
    for kern in ("m52", "m32", "sq_exp"):
        chol = make_cov_chol(oracle_m1.coords, kern, 1.6)
        assert np.allclose(chol, np.triu(chol))                       # scipy returns the UPPER factor
        cov = chol.T @ chol
        assert np.allclose(np.diag(cov), 1.0 + (1e-5 if kern == "sq_exp" else 0.0), atol=1e-10)
    k = sample_field(chol, np.zeros(oracle_m1.n))
    assert np.array_equal(k, np.ones(oracle_m1.n))


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10 (the generator of the device field sampler)."""
    from oracle.thermal_fin_oracle import philox4x32_10, philox_normals
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox4x32_10([np.array([v], np.uint64) for v in ctr], key)
        assert tuple(int(g[0]) for g in got) == want
    z = philox_normals(12345, 400, 501)
    assert z.shape == (400, 501) and np.all(np.isfinite(z))
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    # one stream per row: a row depends only on (seed, global row, subsequence)
    assert np.array_equal(philox_normals(12345, 3, 501, first_row=7), z[7:10])
    assert not np.array_equal(philox_normals(12345, 3, 501, first_row=7, subsequence=1), z[7:10])


def test_exp_quadrature_rule(oracle_m1):
    """The 6-point degree-3 rule of the exp(k) restatement: exact for cubics, e^c for constants, close to the analytic
    cell mean of exp(linear)."""
    o = oracle_m1
    A1 = o.matrix_nodal_exp(np.full(o.n, 0.7))
    A2 = o.matrix_nodal(np.full(o.n, np.exp(0.7)))
    assert abs(A1 - A2).max() <= 1e-13 * abs(A2).max()
    a, b, c = 0.659027622374092, 0.231933368553031, 0.109039009072877
    pts = np.array([[a, b, c], [a, c, b], [b, a, c], [b, c, a], [c, a, b], [c, b, a]])
    assert np.allclose(pts.sum(1), 1.0, atol=1e-14)
    # mean over the reference triangle of l1^p l2^q l3^r is 2 p! q! r! / (p+q+r+2)!
    from math import factorial as f
    for p_, q_, r_ in [(1, 0, 0), (2, 0, 0), (1, 1, 0), (3, 0, 0), (2, 1, 0), (1, 1, 1)]:
        exact = 2.0 * f(p_) * f(q_) * f(r_) / f(p_ + q_ + r_ + 2)
        assert abs((pts[:, 0] ** p_ * pts[:, 1] ** q_ * pts[:, 2] ** r_).mean() - exact) <= 1e-14
    # smooth field: the exp model tends to the plain model at exp(k) under refinement of the data (here: small slope)
    k = 0.05 * o.coords[:, 0]
    w_exp, w_plain = o.forward_exp(k), o.forward(np.exp(k))
    assert np.max(np.abs(w_exp - w_plain)) <= 1e-5 * np.max(np.abs(w_plain))


def test_oracle_golden_regression_next_rows(gold, oracle_m1):
    """Regression pins of the oracle restatements added after the forward map (gradients, nodal LSPG, exp(k), pCN)."""
    from oracle.thermal_fin_oracle import make_cov_chol, pcn_chains, philox_normals
    o = oracle_m1
    k, phi, data = gold["k_nodal"], gold["phi"], gold["data"]
    assert np.allclose(o.gradient(k[1], data), gold["grad_fom"][1], rtol=1e-9, atol=1e-16)
    assert np.allclose(o.sensitivity(k[0]), gold["sens_fom0"], rtol=1e-9, atol=1e-16)
    dJ, J, g = o.grad_reduced(k[2], data, phi)
    assert np.allclose(g, gold["gtheta_rom"][2], rtol=1e-8, atol=1e-16) and abs(J - gold["cost_rom"][2]) <= 1e-10 * J
    A_r, B_r, C_r, x_r, y_r = o.r_fwd_no_full(k[0], phi)
    assert np.allclose(A_r, gold["lspg_Ar0"], rtol=1e-11, atol=1e-14) and abs(y_r - gold["lspg_y"][0]) <= 1e-9 * abs(y_r)
    assert relerr(o.qoi_operator(o.forward_exp(gold["logk"][1])), gold["qoi_exp"][1]) < 1e-12
    assert np.array_equal(philox_normals(2026, 5, o.n, first_row=3, subsequence=2), gold["philox_z"])
    chol = make_cov_chol(o.coords, "m52", 1.6)
    assert np.max(np.abs(chol - gold["chol_m52"])) <= 1e-9
    pcn = pcn_chains(lambda kk: o.qoi_operator(o.forward(kk)), gold["chol_m52"], gold["qoi_nodal"][0], 0.05, 11, 4, 6, 0.2,
                     first_chain=3)
    assert np.array_equal(pcn["accepted"], gold["pcn_accepted"]) and np.allclose(pcn["z"], gold["pcn_z"], atol=1e-12)


def test_exp_gradient_rule(oracle_m1):
    """Degree-4 Strang-Fix rule of the exp(k) gradient form: exact for quartics; the adjoint gradient is consistent with
    finite differences up to the quadrature mismatch; for constant k it reduces to e^c times the plain gradient."""
    from math import factorial as f
    o = oracle_m1
    a1, b1, w1 = 0.816847572980459, 0.091576213509771, 0.109951743655322
    a2, b2, w2 = 0.108103018168070, 0.445948490915965, 0.223381589678011
    pts = np.array([[a1, b1, b1], [b1, a1, b1], [b1, b1, a1], [a2, b2, b2], [b2, a2, b2], [b2, b2, a2]])
    wts = np.array([w1, w1, w1, w2, w2, w2])
    assert abs(wts.sum() - 1.0) < 1e-14
    for p_, q_, r_ in [(1, 0, 0), (2, 1, 0), (4, 0, 0), (2, 2, 0), (2, 1, 1), (3, 1, 0)]:
        exact = 2.0 * f(p_) * f(q_) * f(r_) / f(p_ + q_ + r_ + 2)
        assert abs(np.sum(wts * pts[:, 0] ** p_ * pts[:, 1] ** q_ * pts[:, 2] ** r_) - exact) <= 1e-14
    rng = np.random.default_rng(12)
    data = rng.uniform(0.1, 0.5, 9)
    c = 0.4
    g_exp = o.gradient_exp(np.full(o.n, c), data)
    g_plain = o.gradient(np.full(o.n, np.exp(c)), data)
    assert np.allclose(g_exp, np.exp(c) * g_plain, rtol=1e-10, atol=1e-16)
    k = 0.3 * rng.standard_normal(o.n)
    d = rng.standard_normal(o.n)
    cost = lambda kk: 0.5 * np.sum((o.qoi_operator(o.forward_exp(kk)) - data) ** 2)
    fd = (cost(k + 1e-6 * d) - cost(k - 1e-6 * d)) / 2e-6
    assert abs(fd - o.gradient_exp(k, data) @ d) <= 5e-2 * abs(fd)


def _mesh_with_straddling_top_facet():
    """Unstructured fin mesh whose top edge y = 4 has a boundary facet crossing x = 2.5 (the vertex at the junction is
    slid along the edge), which the reference's mshr mesh may well have."""
    from meshes import unstructured_fin
    coords, cells = unstructured_fin(h=0.2, seed=1)
    coords = coords.copy()
    j = int(np.argmin(np.abs(coords[:, 0] - 2.5) + np.abs(coords[:, 1] - 4.0)))
    assert abs(coords[j, 0] - 2.5) < 1e-12 and abs(coords[j, 1] - 4.0) < 1e-12
    coords[j, 0] = 2.43
    return coords, cells


def test_affine_model_uses_its_own_facet_markers():
    """rom/averaged_affine_ROM.py:116-138 marks the Robin boundary with nine sub-domain objects; a boundary facet that
    straddles x = 2.5 lies in none of their boxes and carries no Robin term in the AFFINE model, while Fin's single
    exterior marker (forward_solve.py:147-152) still covers it.  Oracle and product assembly must agree on both."""
    from oracle.thermal_fin_oracle import FinOracle, mark_facets, mark_facets_affine
    from bayesianinferencedl_b200 import FinSpace
    from bayesianinferencedl_b200.assembly import build_operators
    coords, cells = _mesh_with_straddling_top_facet()
    robin, root = mark_facets(coords, cells)
    robin_a, root_a = mark_facets_affine(coords, cells)
    assert sorted(root) == sorted(root_a)
    missing = sorted(set(robin) - set(robin_a))
    assert len(missing) >= 1 and set(robin_a) <= set(robin)
    for a, b in missing:                                  # exactly the facets crossing x = 2.5 / 3.5
        xs = sorted([coords[a][0], coords[b][0]])
        assert (xs[0] < 2.5 < xs[1]) or (xs[0] < 3.5 < xs[1])
    o = FinOracle(coords, cells)
    ops = build_operators(FinSpace.from_mesh(coords, cells))
    assert abs(ops.csr(ops.vals[0]) - o.Bi * o.M_robin).max() < 1e-15
    assert abs(ops.csr(ops.robin_affine) - o.Bi * o.M_robin_affine).max() < 1e-15
    assert abs(ops.csr(ops.robin_affine) - ops.csr(ops.vals[0])).max() > 1e-4
    theta = np.random.default_rng(5).uniform(0.1, 3.5, 9)
    assert abs(ops.csr(ops.affine_values(theta)) - o.matrix_affine(theta)).max() < 1e-13
    # on the conforming structured mesh the two marker sets coincide
    from bayesianinferencedl_b200 import get_space
    ops2 = build_operators(get_space(40, m=1))
    assert np.array_equal(ops2.robin_affine, ops2.vals[0])
