"""CPU tests of the sparse-direct solver's symbolic phase (C++ in libtfin.so, csrc/frontal_host.h): the per-pivot program
is fetched through the host-only tfin_frontal_* entry points and interpreted in numpy (tests/frontal_emulator.py, the same
steps the CUDA kernels take), then compared with the oracle's sparse LU.  No GPU needed."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from frontal_emulator import fetch_program, run_qoi, run_solve  # noqa: E402
from conftest import relerr  # noqa: E402


def _cvec(theta):
    return np.concatenate([[1.0], theta])


@pytest.mark.parametrize("lookahead", [True, False])
@pytest.mark.parametrize("m", [1, 2])
def test_program_solves_like_the_oracle(m, lookahead, request):
    from bayesianinferencedl_b200.assembly import build_operators
    V = request.getfixturevalue(f"space_m{m}")
    oracle = request.getfixturevalue(f"oracle_m{m}")
    ops = build_operators(V)
    P = fetch_program(ops, lookahead=lookahead)
    assert sorted(P["perm"]) == list(range(ops.n))                      # a permutation of the dofs
    assert P["col_ptr"][-1] == P["nnzL"] and P["cmax"] <= P["nslots"]
    rng = np.random.default_rng(7)
    for theta in rng.uniform(0.1, 10.0, (3, 9)):
        w_ref = oracle.forward_nine_param(theta)
        q_ref = oracle.qoi_operator(w_ref)
        w, q, yy = run_solve(P, _cvec(theta), 9)
        assert np.abs(w - w_ref).max() <= 1e-12 * np.abs(w_ref).max()
        assert relerr(q, q_ref) < 1e-11
        assert abs(yy - ops.rhs @ w_ref) < 1e-12 * abs(yy)              # b.w = y.y (energy identity)
        q2, yy2 = run_qoi(P, _cvec(theta), 9)
        assert relerr(q2, q_ref) < 1e-11 and abs(yy2 - yy) < 1e-13 * yy


def test_program_on_unstructured_nonconforming_mesh():
    from bayesianinferencedl_b200.assembly import build_operators
    from bayesianinferencedl_b200.fom.thermal_fin import FinSpace
    from oracle.thermal_fin_oracle import FinOracle
    from meshes import unstructured_fin
    coords, cells = unstructured_fin(h=0.125, seed=3)
    V = FinSpace.from_mesh(coords, cells)
    ops = build_operators(V)
    oracle = FinOracle(coords, cells)
    theta = np.random.default_rng(1).uniform(0.1, 3.5, 9)
    w_ref = oracle.forward_nine_param(theta)
    for lookahead in (True, False):
        P = fetch_program(ops, lookahead=lookahead)
        w, q, _ = run_solve(P, _cvec(theta), 9)
        assert np.abs(w - w_ref).max() <= 1e-11 * np.abs(w_ref).max()
        assert relerr(q, oracle.qoi_operator(w_ref)) < 1e-10
        q2, _ = run_qoi(P, _cvec(theta), 9)
        assert relerr(q2, oracle.qoi_operator(w_ref)) < 1e-10


def test_front_stays_as_narrow_as_the_strips(space_m3):
    """The ordering (reverse BFS from the root + elimination-tree postorder) must keep the active front at the width of
    the strip it sweeps: 4 m + 2 nodes in the post plus the pending junction rows, not the level-set width."""
    from bayesianinferencedl_b200.assembly import build_operators
    P = fetch_program(build_operators(space_m3))
    assert P["n"] == 1597
    assert P["cmax"] <= 24 and P["nslots"] <= 28
    assert P["nnzL"] < 16000 and P["pair_updates"] < 1.1e5
    # the sample-per-thread kernel's program recycles the pivot's slot before the next column's nodes arrive: the front
    # then never holds more than the widest column plus its pivot (+1), same factor
    P1 = fetch_program(build_operators(space_m3), lookahead=False)
    assert P1["nslots"] < P["nslots"] and P1["nslots"] <= P1["cmax"] + 2
    assert P1["nnzL"] == P["nnzL"] and P1["cmax"] == P["cmax"]
    print("slots with / without lookahead:", P["nslots"], P1["nslots"], "cmax", P["cmax"])
