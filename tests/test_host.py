"""CPU tests of the host layer: mesh, marking, assembly (vs the independent oracle), C-ABI surface, sharding."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_mesh_sizes():
    from bayesianinferencedl_b200 import get_space
    for m in (1, 2, 3, 5):
        V = get_space(40, m=m)
        assert V.dim() == 144 * m * m + 100 * m + 1
        assert V.mesh().num_cells() == 2 * 144 * m * m
        x = V.tabulate_dof_coordinates()
        assert x.min() == 0.0 and x[:, 0].max() == 6.0 and x[:, 1].max() == 4.0
    assert get_space(40).dim() == 1597              # resolution 40 -> m = 3
    # every rectangle corner of thermal_fin.py:7-15 is a vertex, bit-exactly
    xs = {tuple(p) for p in get_space(40).tabulate_dof_coordinates()}
    for c in [(2.5, 0.0), (3.5, 4.0), (0.0, 0.75), (2.5, 1.0), (6.0, 3.75), (3.5, 2.0), (0.0, 4.0)]:
        assert c in xs


def test_assembly_matches_oracle(space_m2, oracle_m2):
    from bayesianinferencedl_b200.assembly import build_operators, nine_param_nodal, five_to_nine
    ops = build_operators(space_m2)
    o = oracle_m2
    assert np.array_equal(ops.cell_markers, o.markers)
    assert np.allclose(ops.rhs, o.B, atol=1e-16)
    assert np.allclose(ops.B_obs, o.B_obs, atol=1e-15)
    assert np.allclose(ops.C, o.C, atol=1e-15)
    rng = np.random.default_rng(4)
    theta = rng.uniform(0.1, 3.5, 9)
    assert abs(ops.csr(ops.affine_values(theta)) - o.matrix_affine(theta)).max() < 1e-13
    k = np.exp(rng.standard_normal(o.n))
    kbar = k[ops.cells].mean(axis=1)
    A = sum(kbar[e] * 0 for e in range(0)) if False else None
    import scipy.sparse as sp
    r = np.repeat(ops.cells, 3, axis=1).ravel(); c = np.tile(ops.cells, (1, 3)).ravel()
    A = sp.coo_matrix(((ops.Ke * kbar[:, None, None]).ravel(), (r, c)), shape=(o.n, o.n)).tocsr() + ops.csr(ops.vals[0])
    assert abs(A - o.matrix_nodal(k)).max() < 1e-12
    assert np.array_equal(nine_param_nodal(ops.coords, theta), o.nine_param_to_function(theta))
    assert np.array_equal(five_to_nine([1, 2, 3, 4, 5]), [1, 2, 3, 4, 5, 4, 3, 2, 1])
    # batched nodal interpolation
    th2 = rng.uniform(0.1, 3.5, (3, 9))
    kn = nine_param_nodal(ops.coords, th2)
    assert kn.shape == (3, o.n) and np.array_equal(kn[1], o.nine_param_to_function(th2[1]))
    # pattern: sorted columns, diagonal present, symmetric values
    for i in range(0, ops.n, 97):
        cols = ops.col_idx[ops.row_ptr[i]:ops.row_ptr[i + 1]]
        assert np.all(np.diff(cols) > 0) and i in cols


def test_nonconforming_mesh_markers():
    """SURVEY Q-1: cells straddling x = 2.5 / 3.5 keep marker 0 (SubDomain.mark semantics)."""
    from bayesianinferencedl_b200.assembly import mark_cells
    from oracle.thermal_fin_oracle import mark_cells as mark_ref
    coords = np.array([[2.4, 0.8], [2.6, 0.8], [2.4, 0.95], [2.6, 0.95], [2.0, 0.8], [2.0, 0.95]])
    cells = np.array([[0, 1, 2], [1, 3, 2], [4, 0, 5], [0, 2, 5]], dtype=np.int32)
    got = mark_cells(coords, cells)
    assert np.array_equal(got, mark_ref(coords, cells))
    assert list(got) == [0, 0, 1, 1]


def test_function_wrapper():
    from bayesianinferencedl_b200 import Function, get_space
    V = get_space(40, m=1)
    f = Function(V)
    assert f.shape == (V.dim(),) and np.all(f == 0)
    f.vector().set_local(np.arange(V.dim(), dtype=float))
    assert f.vector()[:][5] == 5.0 and f.vector().get_local().sum() == np.arange(V.dim()).sum()
    g = Function(V)
    g.assign(f)
    assert np.array_equal(g, f)


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "tfin.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(tfin_[a-z_0-9]+)\s*\(", hdr)))


def test_cabi_exports_every_declared_symbol():
    """libtfin.so loads on a CPU-only box and exports every symbol include/tfin.h declares; the ctypes table
    lists exactly the same set.  (No compute calls without a GPU.)"""
    from bayesianinferencedl_b200 import _build, _cabi
    _build.build()
    lib = _cabi.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 15
    assert sorted(_cabi.SIGNATURES) == declared
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.tfin_version() >= 100


def test_no_cpu_fallback():
    """Without a CUDA device the product path fails loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from bayesianinferencedl_b200 import AffineROMFin, Fin, _cabi, get_space
    V = get_space(40, m=1)
    with pytest.raises(_cabi.TfinError, match="no CUDA device"):
        Fin(V)
    with pytest.raises(_cabi.TfinError):
        AffineROMFin(V, None, np.eye(V.dim())[:, :4])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "bayesianinferencedl_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), os.path.join(dp, f)


def test_rom_offline_tensors(space_m1, oracle_m1):
    """Gram-tensor form of the LSPG system equals the literal psi^T psi of averaged_affine_ROM.py:292-304."""
    from bayesianinferencedl_b200.assembly import build_operators
    from bayesianinferencedl_b200.rom.averaged_affine_ROM import rom_offline_tensors
    ops = build_operators(space_m1)
    rng = np.random.default_rng(3)
    phi = np.linalg.qr(rng.standard_normal((ops.n, 7)))[0]
    S, G, obs_phi = rom_offline_tensors(ops, phi, ops.B_obs)
    theta = rng.uniform(0.1, 3.5, 9)
    th = np.concatenate([[1.0], theta])
    coef = np.array([th[p] * th[q] for p in range(10) for q in range(p, 10)])
    il = np.tril_indices(7)
    A_r = np.zeros((7, 7)); A_r[il] = coef @ S; A_r = A_r + np.tril(A_r, -1).T
    psi = oracle_m1.matrix_affine(theta) @ phi
    assert np.allclose(A_r, psi.T @ psi, rtol=1e-11, atol=1e-13)
    assert np.allclose(th @ G, psi.T @ oracle_m1.B, rtol=1e-11, atol=1e-14)
    assert np.allclose(obs_phi, oracle_m1.B_obs @ phi)


def test_rom_gradient_tensors(space_m1, oracle_m1):
    """Gram-block form of the reduced gradient equals the literal formula of averaged_affine_ROM.py:335-356."""
    from bayesianinferencedl_b200.assembly import build_operators
    from bayesianinferencedl_b200.rom.averaged_affine_ROM import rom_gradient_tensors
    ops = build_operators(space_m1)
    rng = np.random.default_rng(4)
    phi = np.linalg.qr(rng.standard_normal((ops.n, 6)))[0]
    NG = rom_gradient_tensors(ops, phi)
    assert NG.shape == (10, 9, 6, 6)
    k = np.exp(0.3 * rng.standard_normal(ops.n))
    data = rng.uniform(0.1, 0.5, 9)
    dJ, J, g = oracle_m1.grad_reduced(k, data, phi)
    th = np.concatenate([[1.0], oracle_m1.subfin_avg_op(k)])
    psi = oracle_m1.matrix_affine(th[1:]) @ phi
    A_r = psi.T @ psi
    w_r = np.linalg.solve(A_r, psi.T @ oracle_m1.B)
    obs_phi = oracle_m1.B_obs @ phi
    v_r = np.linalg.solve(A_r, obs_phi.T @ (data - obs_phi @ w_r))
    g_gram = np.array([sum(th[t] * v_r @ NG[t, q] @ w_r for t in range(10)) for q in range(9)])
    assert np.allclose(g_gram, g, rtol=1e-10, atol=1e-14 * np.max(np.abs(g)))
    assert np.allclose(g_gram @ ops.B_obs, dJ, rtol=1e-10, atol=1e-14 * np.max(np.abs(dJ)))


def test_oracle_gradients_finite_difference(oracle_m1):
    """The oracle's adjoint gradient / sensitivity restatements (forward_solve.py:293-342) against central differences."""
    o = oracle_m1
    rng = np.random.default_rng(8)
    k = np.exp(0.3 * rng.standard_normal(o.n))
    data = rng.uniform(0.1, 0.5, 9)
    cost = lambda kk: 0.5 * np.sum((o.qoi_operator(o.forward(kk)) - data) ** 2)
    g, Jac = o.gradient(k, data), o.sensitivity(k)
    d = rng.standard_normal(o.n)
    eps = 1e-6
    fd = (cost(k + eps * d) - cost(k - eps * d)) / (2 * eps)
    assert abs(fd - g @ d) <= 1e-6 * abs(fd)
    fdq = (o.qoi_operator(o.forward(k + eps * d)) - o.qoi_operator(o.forward(k - eps * d))) / (2 * eps)
    assert np.allclose(Jac @ d, fdq, rtol=1e-6, atol=1e-12)


def test_oracle_pcn_chains_invariants(oracle_m1):
    """The pCN restatement: beta -> tiny accepts (almost) everything, chains are independent of their neighbours."""
    from oracle.thermal_fin_oracle import make_cov_chol, pcn_chains
    o = oracle_m1
    chol = make_cov_chol(o.coords, "m52", 1.6)
    data = o.qoi_operator(o.forward(np.ones(o.n)))
    f = lambda kk: o.qoi_operator(o.forward(kk))
    a = pcn_chains(f, chol, data, 0.05, 3, 4, 5, 0.2, first_chain=0)
    b = pcn_chains(f, chol, data, 0.05, 3, 2, 5, 0.2, first_chain=2)
    assert np.array_equal(a["accepted"][2:], b["accepted"]) and np.array_equal(a["z"][2:], b["z"])
    assert np.all(a["accepted"] <= 5) and np.all(a["misfit"] >= 0)
    tiny = pcn_chains(f, chol, data, 10.0, 3, 3, 4, 1e-6)
    assert np.all(tiny["accepted"] >= 3)
    assert np.allclose(tiny["qoi_sum"] / 4, tiny["qoi"], rtol=1e-4)


def test_shard_bounds():
    from bayesianinferencedl_b200.dist import shard_bounds, shard_size
    for n, w in [(10, 3), (8, 8), (5, 8), (0, 2), (1000001, 8)]:
        cover = []
        for r in range(w):
            lo, hi = shard_bounds(n, w, r)
            assert 0 <= lo <= hi <= n and hi - lo <= shard_size(n, w)
            cover += list(range(lo, hi)) if n < 100 else []
        if n < 100:
            assert cover == list(range(n))
        assert sum(shard_bounds(n, w, r)[1] - shard_bounds(n, w, r)[0] for r in range(w)) == n


def test_unstructured_nonconforming_mesh_assembly():
    """External (unstructured, non-conforming) mesh through FinSpace.from_mesh: marker-0 cells (SURVEY Q-1), varying
    vertex degree; product assembly == oracle assembly."""
    from meshes import unstructured_fin
    from bayesianinferencedl_b200 import FinSpace
    from bayesianinferencedl_b200.assembly import build_operators
    from oracle.thermal_fin_oracle import FinOracle
    coords, cells = unstructured_fin()
    ops = build_operators(FinSpace.from_mesh(coords, cells))
    o = FinOracle(coords, cells)
    assert (ops.cell_markers == 0).sum() > 0 and np.array_equal(ops.cell_markers, o.markers)
    assert np.diff(ops.row_ptr).max() >= 9
    theta = np.random.default_rng(0).uniform(0.1, 3.5, 9)
    assert abs(ops.csr(ops.affine_values(theta)) - o.matrix_affine(theta)).max() < 1e-12
    assert np.allclose(ops.rhs, o.B) and np.allclose(ops.B_obs, o.B_obs, atol=1e-15)
    # marker-0 cells carry no conductivity in the affine model but do in the nodal model
    assert np.abs(ops.k_unmarked).max() > 0


def test_enrich_matches_reference_recipe():
    """model_constr_adaptive_sampling.py:52-68: Gram-Schmidt against columns 0..k-2 (the reference's loop bound)."""
    from bayesianinferencedl_b200.rom.model_constr_adaptive_sampling import enrich
    rng = np.random.default_rng(6)
    basis = np.linalg.qr(rng.standard_normal((30, 4)))[0]
    w = rng.standard_normal((30, 1))
    U = enrich(basis, w)
    assert U.shape == (30, 5) and np.array_equal(U[:, :4], basis)
    assert abs(np.linalg.norm(U[:, 4]) - 1.0) < 1e-14
    assert np.allclose(U[:, :3].T @ U[:, 4], 0.0, atol=1e-14)          # orthogonal to all but the last column
    ref = w[:, 0].copy()                                                  # literal restatement
    for j in range(0, 4 - 1):
        ref = ref - (ref @ basis[:, j]) / (basis[:, j] @ basis[:, j]) * basis[:, j]
    assert np.allclose(U[:, 4], ref / np.sqrt(ref @ ref), atol=1e-15)


def test_mass_and_stiffness_matrices(space_m1, oracle_m1):
    """Fin.M / Fin.K (forward_solve.py:172-173) from the host operators: partition of unity and polynomial exactness."""
    from bayesianinferencedl_b200.assembly import build_operators
    ops = build_operators(space_m1)
    M = ops.mass_matrix().toarray()
    K = ops.csr(ops.stiffness_values()).toarray()
    one = np.ones(ops.n)
    assert abs(one @ M @ one - oracle_m1.domain_measure) <= 1e-12 * oracle_m1.domain_measure     # int 1 dx = |domain|
    assert np.allclose(M, M.T) and np.allclose(K, K.T) and np.allclose(K @ one, 0.0, atol=1e-12)
    x, y = ops.coords[:, 0], ops.coords[:, 1]
    assert abs(x @ K @ x - oracle_m1.domain_measure) <= 1e-10 * oracle_m1.domain_measure        # int |grad x|^2 = |domain|
    assert np.allclose(K, sum(Kq.toarray() for Kq in oracle_m1.K_q), atol=1e-13)                # conforming mesh: all cells marked
    assert np.allclose(M @ one, oracle_m1.C * oracle_m1.domain_measure, rtol=1e-12)             # row sums = int phi_i
