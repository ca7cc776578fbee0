#!/usr/bin/env python
"""Export golden vectors FROM THE REFERENCE ITSELF -- run this inside the reference's FEniCS 2018.1 docker
(/root/reference/README.md:16-23: FEniCS 2018.1 + mshr + petsc4py + TensorFlow 1.x), never here.

    cd <checkout of sheroze1123/BayesianInferenceDL>/rom        # the reference uses '../data', '../bayesian_inference' paths
    python /path/to/export_reference_fenics.py reference_fenics.npz

then copy the file to  tests/golden/reference_fenics.npz  of this repository.  tests/test_reference_fenics.py consumes it
when present and turns "parity unpinned"  into parity pinned by the reference: the oracle restatement
(CPU suite) and the CUDA path (-m gpu) are both checked against the stored outputs of dolfin / PETSc / numpy at 1e-10.

What is stored (every array in dolfin DOF order -- what ``u.vector()[:]`` returns -- so that row i of
data/basis_nine_param.txt and data/B_obs.txt is dof i):
    dof_coords (n, 2), cells_dof (nc, 3)   the mshr mesh of get_space(40) (fom/thermal_fin.py:4-20) with cell vertices
                                           renumbered vertex -> dof (vertex_to_dof_map)
    B (n,), B_obs (9, n), C (n,), domain_measure
                                           Fin.__init__ state (fom/forward_solve.py:125-165, 205-231)
    theta (S, 9), w_affine (S, n), qoi_affine (S, 9)
                                           AffineROMFin affine FOM, averaged_k_s[i] = theta_i then solve(_F == _a)
                                           (rom/averaged_affine_ROM.py:251-256), qoi (:312-321)
    w_r (S, n_r), qoi_r (S, 9), phi (n, n_r)
                                           forward_nine_param_reduced / qoi_reduced (:278-333) with data/basis_nine_param.txt
    k_nodal (T, n), w_nodal (T, n), qoi_nodal (T, 9), theta_of_k (T, 9)
                                           Fin.forward + qoi_operator (fom/forward_solve.py:270-291, 408-412) and
                                           subfin_avg_op (:466-480) on Matern-5/2 fields exp(0.5 chol^T z)
                                           (bayesian_inference/gaussian_field.py:9-31, deep_learning/generate_fin_dataset.py:87-88)
    k_nine (9 -> n), w_nine_fn             nine_param_to_function(theta[0]) and Fin.forward of it (SURVEY Q-2)
    grad_k0, data0                         Fin.gradient(k_nodal[0], data0) (fom/forward_solve.py:293-322)
    z_true (5,), first row of theta        bayesian_inference/muq_old/bayes_inv.py:29 mapped to nine parameters
This script only CALLS the reference; it contains none of its code.
"""
import os
import sys

import numpy as np


def main(out_path):
    here = os.getcwd()
    sys.path.append(os.path.join(here, ".."))
    import dolfin as dl
    from fom.thermal_fin import get_space
    from fom.forward_solve import Fin
    from bayesian_inference.gaussian_field import make_cov_chol

    dl.set_log_level(40)
    V = get_space(40)
    mesh = V.mesh()
    v2d = dl.vertex_to_dof_map(V)
    n = V.dim()
    out = {
        "dof_coords": V.tabulate_dof_coordinates().reshape((-1, 2))[V.dofmap().dofs(), :],
        "cells_dof": v2d[mesh.cells()].astype(np.int64),
        "vertex_to_dof": np.asarray(v2d, dtype=np.int64),
        "vertex_coords": mesh.coordinates().copy(),
    }

    fin = Fin(V)
    out["B"] = np.asarray(fin.B, dtype=np.float64)
    out["B_obs"] = np.asarray(fin.B_obs, dtype=np.float64)
    out["C"] = np.asarray(fin.C, dtype=np.float64).ravel()
    out["domain_measure"] = float(fin.domain_measure)

    # ---- affine model (needs TensorFlow 1.x + a Keras model for the constructor's placeholders, :211-235)
    phi = np.loadtxt("../data/basis_nine_param.txt", delimiter=",")
    out["phi"] = phi
    rom = None
    try:
        from tensorflow.keras.models import Sequential
        from tensorflow.keras.layers import Dense
        from rom.averaged_affine_ROM import AffineROMFin
        err_model = Sequential([Dense(9, input_shape=(9,))])
        rom = AffineROMFin(V, err_model, phi)
    except Exception as exc:                      # TF-free third copy of the class (no error model)
        print("averaged_affine_ROM.AffineROMFin unavailable (%r); using generate_reduced_basis_nine_param's copy" % (exc,))
        from rom.generate_reduced_basis_nine_param import AffineROMFin as AffineNoNN
        rom = AffineNoNN(V)
        rom.phi = phi

    rng = np.random.RandomState(0)
    z_true = np.array([0.41126864, 0.61789679, 0.75873243, 0.96527541, 0.22348076])
    theta = np.vstack([np.concatenate([z_true, z_true[3::-1]]), rng.uniform(0.1, 3.5, (10, 9))])
    out["z_true"], out["theta"] = z_true, theta
    w_aff, q_aff, w_r, q_r = [], [], [], []
    for th in theta:
        for i in range(9):
            rom.averaged_k_s[i].assign(float(th[i]))
        w = dl.Function(V)
        dl.solve(rom._F == rom._a, w)                       # what AffineROMFin.forward does after averaging, :256
        w_aff.append(w.vector()[:].copy())
        q_aff.append(np.dot(out["B_obs"], w_aff[-1]))
        if hasattr(rom, "forward_nine_param_reduced"):
            wr = rom.forward_nine_param_reduced(th)
            w_r.append(np.asarray(wr, dtype=np.float64).ravel())
            q_r.append(np.asarray(rom.qoi_reduced(wr), dtype=np.float64).ravel())
    out["w_affine"], out["qoi_affine"] = np.array(w_aff), np.array(q_aff)
    if w_r:
        out["w_r"], out["qoi_r"] = np.array(w_r), np.array(q_r)

    # ---- nodal model on Gaussian-field conductivities
    chol = make_cov_chol(V, length=1.6)
    rs = np.random.RandomState(3)
    k_fields, w_nod, q_nod, th_k = [], [], [], []
    for _ in range(6):
        nodal = np.exp(0.5 * np.dot(chol.T, rs.randn(n)))
        k = dl.Function(V)
        k.vector().set_local(nodal)
        w = fin.forward(k)[0]
        k_fields.append(nodal)
        w_nod.append(w.vector()[:].copy())
        q_nod.append(np.asarray(fin.qoi_operator(w), dtype=np.float64).ravel())
        th_k.append(np.asarray(fin.subfin_avg_op(k), dtype=np.float64).ravel())
    out["k_nodal"], out["w_nodal"] = np.array(k_fields), np.array(w_nod)
    out["qoi_nodal"], out["theta_of_k"] = np.array(q_nod), np.array(th_k)

    k9 = fin.nine_param_to_function(theta[0])
    out["k_nine"] = k9.vector()[:].copy()
    out["w_nine_fn"] = fin.forward(k9)[0].vector()[:].copy()

    try:
        data0 = out["qoi_nodal"][1]
        k0 = dl.Function(V)
        k0.vector().set_local(k_fields[0])
        g = fin.gradient(k0, data0)
        out["grad_k0"] = np.asarray(g, dtype=np.float64).ravel()
        out["data0"] = data0
    except Exception as exc:
        print("Fin.gradient not exported: %r" % (exc,))

    np.savez_compressed(out_path, **out)
    print("wrote %s: n = %d dofs, %d cells, %d affine samples, %d nodal samples"
          % (out_path, n, mesh.num_cells(), len(theta), len(k_fields)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "reference_fenics.npz")
