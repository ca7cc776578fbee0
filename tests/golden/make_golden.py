"""Generates the committed golden fixtures.  Run from the repo root IN THE BUILD CONTAINER (it reads
/root/reference, which does not exist on the GPU box):   python tests/golden/make_golden.py

1. reference_data.npz -- everything the reference's shipped data pins for this path:
   * data/B_obs.txt (9 x 1446) stored sparse; its row sums are 1 (observation_operator, forward_solve.py:488-511)
   * the first 8 columns of data/basis_{five,nine}_param.txt and the matching columns of B_obs @ phi
     (the one hot-path quantity reproducible from data alone, averaged_affine_ROM.py:212)
   * column norms of both shipped bases (conditioning fixtures for the ROM Cholesky, SURVEY Q-4)
   * the five-parameter "truth" of bayesian_inference/muq_old/bayes_inv.py:29
   The reference ships NO forward-solve output, so nothing else can be pinned (parity unpinned).
2. oracle_m1.npz -- outputs of oracle/thermal_fin_oracle.py on the m=1 mesh for fixed seeds (regression pins of
   the oracle itself; they are NOT reference outputs).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def reference_data():
    B = np.loadtxt(os.path.join(REF, "data", "B_obs.txt"), delimiter=",")
    assert np.array_equal(B, np.loadtxt(os.path.join(REF, "rom", "B_obs.txt"), delimiter=","))
    r, c = np.nonzero(B)
    out = {"B_obs_shape": np.array(B.shape), "B_obs_rows": r.astype(np.int32), "B_obs_cols": c.astype(np.int32),
           "B_obs_vals": B[r, c], "z_true_five": np.array([0.41126864, 0.61789679, 0.75873243, 0.96527541, 0.22348076])}
    for name in ("five", "nine"):
        phi = np.loadtxt(os.path.join(REF, "data", f"basis_{name}_param.txt"), delimiter=",")
        out[f"phi_{name}_shape"] = np.array(phi.shape)
        out[f"phi_{name}_head"] = phi[:, :8].copy()
        out[f"B_obs_phi_{name}_head"] = np.dot(B, phi)[:, :8]
        out[f"phi_{name}_colnorm"] = np.linalg.norm(phi, axis=0)
    np.savez_compressed(os.path.join(OUT, "reference_data.npz"), **out)


def oracle_m1():
    from bayesianinferencedl_b200 import get_space
    from oracle.thermal_fin_oracle import FinOracle, five_param_to_nine, make_cov_chol, sample_field
    V = get_space(40, m=1)
    o = FinOracle(V.mesh().coordinates(), V.mesh().cells())
    rng = np.random.default_rng(2024)
    theta = rng.uniform(0.1, 3.5, (6, 9))
    k5 = rng.uniform(0.1, 1.0, (3, 5))
    chol = make_cov_chol(o.coords, "m52", 1.6)
    knod = np.stack([sample_field(chol, rng.standard_normal(o.n)) for _ in range(3)])
    phi = np.linalg.qr(rng.standard_normal((o.n, 10)))[0]
    out = {
        "theta": theta, "qoi_affine": np.stack([o.qoi_operator(o.forward_nine_param(t)) for t in theta]),
        "w_affine0": o.forward_nine_param(theta[0]),
        "k5": k5, "qoi_five": np.stack([o.qoi_operator(o.forward_five_param_affine(k)) for k in k5]),
        "k_nodal": knod, "qoi_nodal": np.stack([o.qoi_operator(o.forward(k)) for k in knod]),
        "theta_of_k": np.stack([o.subfin_avg_op(k) for k in knod]),
        "phi": phi, "qoi_rom": np.stack([o.qoi_reduced(o.forward_nine_param_reduced(t, phi), phi) for t in theta]),
        "nine_to_fn": o.nine_param_to_function(theta[0]),
    }
    # rows added for the 'next' components (gradients, nodal LSPG, exp(k), prior, Philox, pCN)
    from oracle.thermal_fin_oracle import pcn_chains, philox_normals
    data = rng.uniform(0.05, 0.6, 9)
    gr = [o.grad_reduced(k, data, phi) for k in knod]
    rf = [o.r_fwd_no_full(k, phi) for k in knod]
    logk = 0.5 * rng.standard_normal((3, o.n))
    pcn = pcn_chains(lambda kk: o.qoi_operator(o.forward(kk)), chol, out["qoi_nodal"][0], 0.05, 11, 4, 6, 0.2, first_chain=3)
    out.update({
        "data": data, "grad_fom": np.stack([o.gradient(k, data) for k in knod]),
        "sens_fom0": o.sensitivity(knod[0]),
        "grad_rom": np.stack([g[0] for g in gr]), "cost_rom": np.array([g[1] for g in gr]),
        "gtheta_rom": np.stack([g[2] for g in gr]),
        "lspg_Ar0": rf[0][0], "lspg_Br": np.stack([r[1] for r in rf]), "lspg_y": np.array([r[4] for r in rf]),
        "logk": logk, "qoi_exp": np.stack([o.qoi_operator(o.forward_exp(k)) for k in logk]),
        "chol_m52": chol, "philox_z": philox_normals(2026, 5, o.n, first_row=3, subsequence=2),
        "pcn_accepted": pcn["accepted"], "pcn_z": pcn["z"], "pcn_qoi_sum": pcn["qoi_sum"], "pcn_misfit": pcn["misfit"],
    })
    np.savez_compressed(os.path.join(OUT, "oracle_m1.npz"), **out)


if __name__ == "__main__":
    reference_data()
    oracle_m1()
    for f in ("reference_data.npz", "oracle_m1.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")
