/*
 * tfin.h -- C ABI of the B200-native thermal-fin batched forward map (libtfin.so).
 *
 * The reference (sheroze1123/BayesianInferenceDL) has no FFI: its hot path is Python methods calling
 * FEniCS/PETSc/LAPACK one conductivity sample at a time.  Every entry point below therefore names the
 * reference METHOD (file:line under /root/reference) whose per-sample arithmetic it performs for a whole
 * batch; the Python facade in bayesianinferencedl_b200/ binds them with ctypes.
 *
 * Conventions
 *   - all functions return 0 on success, <0 on error; tfin_last_error() gives the message (thread local)
 *   - all floating point data is IEEE fp64, row-major, C-contiguous; indices are int32
 *   - the caller owns every buffer; the library owns only the handle and its device workspaces
 *   - `mem` says where the batch buffers (inputs AND outputs of a solve call) live:
 *       TFIN_MEM_HOST   host pointers (pageable or pinned); the call copies H2D/D2H on its stream and
 *                       returns after the outputs are complete
 *       TFIN_MEM_DEVICE device pointers on the handle's device; the call only enqueues work on `stream`
 *   - `stream` is a cudaStream_t passed as void* (NULL = the handle's own stream)
 *   - operator/setup arrays passed to tfin_set_* are always HOST pointers and are copied
 *   - there is no CPU fallback: every solve call fails if no sm_100 device is usable
 */
#ifndef TFIN_H
#define TFIN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tfin_ctx* tfin_handle_t;

#define TFIN_MEM_HOST 0
#define TFIN_MEM_DEVICE 1

/* what the per-sample input rows of a solve call are */
#define TFIN_IN_PARAMS 0 /* (N, n_terms-1) sub-domain conductivities theta            */
#define TFIN_IN_NODAL 1  /* (N, n) nodal conductivity fields k                         */

/* per-sample status */
#define TFIN_STATUS_CONVERGED 0
#define TFIN_STATUS_MAXIT 1     /* PCG hit maxit before reaching tol                    */
#define TFIN_STATUS_BREAKDOWN 2 /* non-positive curvature / non-SPD pivot / NaN         */

#define TFIN_MAX_TERMS 16

int tfin_version(void);
const char* tfin_last_error(void);

/* Create / destroy a handle bound to CUDA device `device`. */
int tfin_create(int device, tfin_handle_t* out);
int tfin_destroy(tfin_handle_t h);

/*
 * Shared-pattern affine operator  A(theta) = vals[0] + sum_{q>=1} theta_q vals[q]  and right-hand side.
 * Replaces the UFL forms + dolfin.assemble of AffineROMFin.__init__ / forward
 * (rom/averaged_affine_ROM.py:156-163, 237-258) and of Fin.__init__ (fom/forward_solve.py:160-163):
 * vals[0] = Bi*M_Gamma, vals[q] = stiffness over sub-domain q, rhs = int_root v.
 *   row_ptr[n+1], col_idx[nnz] : CSR pattern (must contain the diagonal), vals[n_terms][nnz], rhs[n].
 * prune_zeros != 0 drops off-diagonal entries that are exactly 0.0 in every term.
 */
int tfin_set_operator(tfin_handle_t h, int32_t n, int32_t nnz, const int32_t* row_ptr,
                      const int32_t* col_idx, int32_t n_terms, const double* vals, const double* rhs,
                      int32_t prune_zeros);

/*
 * Observation operator B_obs as CSR (n_obs rows over n dofs): Fin.qoi_operator / AffineROMFin.qoi
 * (fom/forward_solve.py:408-412, 488-511; rom/averaged_affine_ROM.py:312-321).
 */
int tfin_set_observation(tfin_handle_t h, int32_t n_obs, const int32_t* ptr, const int32_t* idx,
                         const double* val);

/*
 * Sub-fin averaging operator (n_terms-1 rows over n dofs) mapping a nodal field to theta:
 * subfin_avg_op (fom/forward_solve.py:466-480, rom/averaged_affine_ROM.py:404-418).
 * Needed for TFIN_IN_NODAL inputs of tfin_fom_affine / tfin_rom.
 */
int tfin_set_averaging(tfin_handle_t h, int32_t n_rows, const int32_t* ptr, const int32_t* idx,
                       const double* val);

/*
 * Mesh cells + element stiffness matrices for the nodal-conductivity operator
 * A(k) = sum_e mean(k on e) K_e + vals[0]  (Fin._F, fom/forward_solve.py:160-161).
 *   cells[n_cells][3], Ke[n_cells][3][3].   Requires tfin_set_operator first (pattern, vals[0], rhs).
 */
int tfin_set_cells(tfin_handle_t h, int32_t n_cells, const int32_t* cells, const double* Ke,
                   int32_t prune_zeros);

/*
 * Offline tensors of the LSPG reduced model (rom/averaged_affine_ROM.py:292-304, 212) for a basis phi
 * (n x n_r), with Psi_t = vals[t] phi:
 *   S[n_pairs][T]   n_pairs = n_terms(n_terms+1)/2 pairs (p<=q, p-major), T = n_r(n_r+1)/2 lower triangle
 *                   row-major packed (i>=j -> i(i+1)/2+j);  S_pq = Psi_p^T Psi_q (+ transpose if p<q)
 *   G[n_terms][n_r] G_t = Psi_t^T rhs
 *   obs_phi[n_obs][n_r] = B_obs phi
 */
int tfin_set_rom(tfin_handle_t h, int32_t n_r, int32_t n_terms, int32_t n_obs, const double* S,
                 const double* G, const double* obs_phi);

/*
 * Batched affine full-order solve A(theta_s) w_s = rhs by Jacobi-PCG, fused with qoi_s = B_obs w_s.
 * = AffineROMFin.forward + qoi (rom/averaged_affine_ROM.py:237-258, 312-321) and forward_nine_param
 * (rom/generate_reduced_basis_nine_param.py:178-183) for N samples.
 *   in: (N, n_terms-1) if in_kind == TFIN_IN_PARAMS, (N, n) if TFIN_IN_NODAL
 *   tol: stop when sqrt(r.z / r0.z0) <= tol;  maxit: iteration cap
 *   w_out (N, n) | NULL, qoi_out (N, n_obs) | NULL, iters_out (N) | NULL, status_out (N) | NULL,
 *   relres_out (N) | NULL: true relative residual ||b - A x|| / ||b|| recomputed after the last iteration
 */
int tfin_fom_affine(tfin_handle_t h, const double* in, int64_t N, int32_t in_kind, int32_t mem,
                    double tol, int32_t maxit, double* w_out, double* qoi_out, int32_t* iters_out,
                    int32_t* status_out, double* relres_out, void* stream);

/*
 * Batched nodal-conductivity full-order solve (per-sample in-kernel assembly) fused with B_obs:
 * = Fin.forward + qoi_operator (fom/forward_solve.py:270-291, 408-412) for N samples; k: (N, n).
 */
int tfin_fom_nodal(tfin_handle_t h, const double* k, int64_t N, int32_t mem, double tol, int32_t maxit,
                   double* w_out, double* qoi_out, int32_t* iters_out, int32_t* status_out,
                   double* relres_out, void* stream);

/*
 * Batched LSPG reduced solve: A_r = sum_{p<=q} th_p th_q S_pq, B_r = sum_t th_t G_t (th_0 = 1),
 * Cholesky solve, qoi = obs_phi w_r.
 * = AffineROMFin.forward_reduced / forward_nine_param_reduced + qoi_reduced
 * (rom/averaged_affine_ROM.py:260-310, 323-333) for N samples.
 *   wr_out (N, n_r) | NULL, qoi_out (N, n_obs) | NULL, status_out (N) | NULL
 */
int tfin_rom(tfin_handle_t h, const double* in, int64_t N, int32_t in_kind, int32_t mem, double* wr_out,
             double* qoi_out, int32_t* status_out, void* stream);

/*
 * Reduced basis for the nodal-conductivity LSPG model (Fin.reduced_forward / r_fwd_no_full,
 * fom/forward_solve.py:421-464): phi (n, n_r) row-major and the rows projected onto the reduced solution,
 * out_phi (n_out, n_r), e.g. C_r = C phi (:442) and B_obs phi (:415-419).  Requires tfin_set_operator.
 */
int tfin_set_basis(tfin_handle_t h, int32_t n, int32_t n_r, const double* phi, int32_t n_out, const double* out_phi);

/*
 * Batched nodal-conductivity LSPG reduced solve = Fin.r_fwd_no_full(k, phi) (fom/forward_solve.py:454-464) for N fields:
 *   psi = A(k) phi (A assembled per sample in-kernel), A_r = psi^T A phi, B_r = psi^T b, x_r = A_r^{-1} B_r,
 *   y = out_phi x_r.   Requires tfin_set_cells and tfin_set_basis.
 *   Ar_out (N, n_r, n_r) | NULL, Br_out (N, n_r) | NULL, xr_out (N, n_r) | NULL, y_out (N, n_out) | NULL,
 *   status_out (N) | NULL
 */
int tfin_rom_nodal(tfin_handle_t h, const double* k, int64_t N, int32_t mem, double* Ar_out, double* Br_out,
                   double* xr_out, double* y_out, int32_t* status_out, void* stream);

/*
 * Offline Gram blocks for the reduced gradient: gram[t][q-1] = Psi_t^T Psi_q (n_r x n_r, row-major) for
 * t = 0..n_terms-1, q = 1..n_terms-1 with Psi_t = vals[t] phi, i.e. psi^T (dA_dsigmak_phi[q-1]) =
 * sum_t theta_t gram[t][q-1]  (rom/averaged_affine_ROM.py:215-220, 343-348).  Requires tfin_set_rom first.
 */
int tfin_set_rom_gradient(tfin_handle_t h, int32_t n_r, int32_t n_terms, const double* gram);

/*
 * Batched reduced-model gradient of J_s = 0.5 ||data_s - (B_obs phi) w_r,s||^2, exactly the reference's formula
 * = AffineROMFin.grad_reduced(k) (rom/averaged_affine_ROM.py:335-356) for N samples:
 *   w_r = A_r^{-1} B_r;  v_r = A_r^{-T} (B_obs phi)^T (data - (B_obs phi) w_r);
 *   g_q = (psi v_r)^T (K_q phi) w_r  (q = 1..n_terms-1);   dJ_dk = g^T dsigma_dk.
 *   in: (N, n_terms-1) conductivities (TFIN_IN_PARAMS) or (N, n) nodal fields averaged first (TFIN_IN_NODAL)
 *   data: (data_rows, n_obs) with data_rows == 1 or N
 *   grad_kind: TFIN_IN_PARAMS -> grad_out (N, n_terms-1) = g;  TFIN_IN_NODAL -> grad_out (N, n) = dJ_dk
 *              (needs tfin_set_averaging: dsigma_dk = the sub-fin averaging operator, :210)
 *   cost_out (N) | NULL, qoi_out (N, n_obs) | NULL, wr_out (N, n_r) | NULL, status_out (N) | NULL
 */
int tfin_rom_gradient(tfin_handle_t h, const double* in, int64_t N, int32_t in_kind, int32_t mem,
                      const double* data, int64_t data_rows, int32_t grad_kind, double* grad_out, double* cost_out,
                      double* qoi_out, double* wr_out, int32_t* status_out, void* stream);

/*
 * Batched adjoint gradient of J_s = 0.5 ||B_obs w_s - data_s||^2 with respect to the nodal conductivity:
 * = Fin.gradient(k, data) (fom/forward_solve.py:293-322) for N samples.  The forward solve, the adjoint solve
 * A v = -B_obs^T (B_obs w - data) (the reference's dense np.linalg.solve, :310) and the gradient form
 * assemble(k_hat * inner(grad w, grad v) * dx) (:313-314).  Default (fom_solver 0 / 2, fronts of <= 32 nodes): ONE sparse
 * direct factorisation serves both solves -- backward substitution (w, observables), forward substitution with the adjoint
 * right-hand side, backward substitution (v), gradient-form kernel; tol / maxit are ignored and iters_out is 0.  Otherwise
 * (fom_solver = 1, wider fronts): forward PCG, adjoint PCG and the gradient form in one kernel on the same on-chip operator.
 *   data: (data_rows, n_obs) with data_rows == 1 (one observation vector for all samples) or N
 *   grad_out (N, n), cost_out (N) | NULL = J_s, qoi_out (N, n_obs) | NULL, iters_out (forward solve) | NULL,
 *   status_out | NULL (worst of the forward and adjoint solves)
 */
int tfin_fom_nodal_gradient(tfin_handle_t h, const double* k, int64_t N, int32_t mem, double tol, int32_t maxit,
                            const double* data, int64_t data_rows, double* grad_out, double* cost_out,
                            double* qoi_out, int32_t* iters_out, int32_t* status_out, void* stream);

/*
 * Batched Jacobian of the observables with respect to the nodal conductivity:
 * = Fin.sensitivity(k) (fom/forward_solve.py:324-342): n_obs adjoint solves A v_o = -B_obs[o,:]^T per sample (with the direct
 * solver: n_obs forward + backward substitution passes on the one factor).
 *   jac_out (N, n_obs, n)
 */
int tfin_fom_nodal_sensitivity(tfin_handle_t h, const double* k, int64_t N, int32_t mem, double tol, int32_t maxit,
                               double* jac_out, double* qoi_out, int32_t* iters_out, int32_t* status_out,
                               void* stream);

/* covariance kernels of tfin_field_set_cov (bayesian_inference/gaussian_field.py:16-28) */
#define TFIN_KERN_SQ_EXP 0 /* exp(-d^2 / (2 l^2)) + 1e-5 I */
#define TFIN_KERN_M52 1    /* Matern 5/2: (1 + t + t^2/3) exp(-t), t = sqrt(5) d / l */
#define TFIN_KERN_M32 2    /* Matern 3/2: (1 + t) exp(-t),          t = sqrt(3) d / l */

/*
 * Gaussian-random-field prior over the dofs = make_cov_chol(V, kern_type, length)
 * (bayesian_inference/gaussian_field.py:9-31): pairwise distances of coords (n_pts, 2), covariance kernel and
 * Cholesky factor, all on the device.  chol_out (n_pts, n_pts) | NULL receives the UPPER factor (cov = chol^T chol)
 * like scipy.linalg.cholesky; the handle keeps the factor for tfin_field_sample.
 */
int tfin_field_set_cov(tfin_handle_t h, int32_t n_pts, const double* coords, int32_t kern_type, double length,
                       double* chol_out);

/* Use a caller-supplied upper factor chol (n_pts, n_pts), e.g. of a stored prior covariance. */
int tfin_field_set_chol(tfin_handle_t h, int32_t n_pts, const double* chol);

/*
 * Batched conductivity draws k_s = exp(0.5 * chol^T z_s) (deep_learning/generate_fin_dataset.py:87-88).
 *   z (N, n_pts) standard normals, or NULL: drawn on the device, one independent Philox4x32-10 stream per row:
 *   entries (2p, 2p+1) of row r = Box-Muller of counter (p, first_row + r, subsequence) under key `seed`, so a sample
 *   depends only on (seed, its global index, subsequence) -- not on chunking or on the rank that draws it.
 *   k_out (N, n_pts);  z_out (N, n_pts) | NULL
 */
int tfin_field_sample(tfin_handle_t h, const double* z, uint64_t seed, uint32_t subsequence, int64_t first_row,
                      int64_t N, int32_t mem, double* k_out, double* z_out, void* stream);

/*
 * Many-chain preconditioned Crank-Nicolson Metropolis on the Gaussian-field prior, all chains advanced together on the
 * device: the batched replacement of one likelihood evaluation per sampler proposal
 * (bayesian_inference/pymc_func_bayes_inverse.py:68-90, 201; inference.py:42-52).
 *   state z_c (n normals), k_c = exp(0.5 chol^T z_c), misfit Phi = 0.5 ||qoi(k_c) - data||^2 / sigma^2,
 *   proposal z' = sqrt(1 - beta^2) z + beta xi, accepted when log u < Phi(z) - Phi(z').
 *   model: 0 = nodal full-order solve (tfin_fom_nodal), 1 = sub-fin-averaged LSPG ROM (tfin_rom, nodal input).
 *   Chain c is global chain first_chain + c: all its random numbers are the Philox streams (seed, first_chain + c,
 *   step), so results do not depend on how chains are sharded over GPUs; steps are numbered first_step + 1 ...
 *   first_step + n_steps (pass the previous total to continue a run).
 *   data (n_obs) HOST pointer.  z_state (C, n) in/out (ignored on input if init_from_prior != 0: draw 0 of the prior).
 *   Outputs (each | NULL): misfit_out (C) of the final state, accepted_out (C) accepted proposals, qoi_out (C, n_obs)
 *   final observables, qoi_sum_out / qoi_sq_out (C, n_obs) sums over the steps of the chain's current observables,
 *   k_sum_out (C, n) sum over the steps of the current conductivity field.
 */
int tfin_pcn_chains(tfin_handle_t h, int32_t model, int64_t C, int64_t first_chain, int32_t n_steps, int32_t first_step,
                    double beta, const double* data, double sigma, uint64_t seed, double tol, int32_t maxit, int32_t mem,
                    double* z_state, int32_t init_from_prior, double* misfit_out, int64_t* accepted_out, double* qoi_out,
                    double* qoi_sum_out, double* qoi_sq_out, double* k_sum_out, void* stream);

/* theta = averaging(k) for N nodal fields (subfin_avg_op batched); out (N, n_rows). */
int tfin_subfin_avg(tfin_handle_t h, const double* k, int64_t N, int32_t mem, double* theta_out,
                    void* stream);

/*
 * Diagnostics of the sparse-direct solver (host only, no device needed).  The direct solve replaces dolfin's default
 * sparse LU (fom/forward_solve.py:286, rom/averaged_affine_ROM.py:256): the symbolic phase (ordering, elimination tree,
 * fill, front storage) is shared by all samples and compiled into a per-pivot program that the kernels interpret.
 * tfin_frontal_analyze runs that phase for an affine operator given like tfin_set_operator (+ optional B_obs CSR) and
 * returns an opaque program; tfin_frontal_array copies one of its arrays ("perm", "piv_slot", "col_ptr", "col_slot",
 * "rhs", "asm_ptr", "asm_addr", "asm_eptr", "ent_term", "ent_coef", "obs_ptr", "obs_row", "obs_val") into dst if
 * dst_bytes suffices and returns its size in bytes, or the value of a scalar ("n", "nslots", "cmax", "nnzL",
 * "pair_updates"); -1 for an unknown name.  tests/frontal_emulator.py interprets the program on the CPU.
 */
int tfin_frontal_analyze(int32_t n, int32_t nnz, const int32_t* row_ptr, const int32_t* col_idx, int32_t n_terms,
                         const double* vals, const double* rhs, int32_t n_obs, const int32_t* obs_ptr,
                         const int32_t* obs_idx, const double* obs_val, void** out);
/* lookahead = 1: as tfin_frontal_analyze (front slots of column j + 1 allocated one step early -- the sample-per-CTA kernel
 * assembles it while step j updates); 0: the pivot's slot is recycled first (the sample-per-thread kernel's program). */
int tfin_frontal_analyze_ex(int32_t n, int32_t nnz, const int32_t* row_ptr, const int32_t* col_idx, int32_t n_terms,
                            const double* vals, const double* rhs, int32_t n_obs, const int32_t* obs_ptr,
                            const int32_t* obs_idx, const double* obs_val, int32_t lookahead, void** out);
int64_t tfin_frontal_array(void* prog, const char* name, void* dst, int64_t dst_bytes);
void tfin_frontal_free(void* prog);

/*
 * Micro-benchmark: aggregate shared-memory read bandwidth of the device in GB/s (conflict-free 8-byte loads, the access width of
 * the solver kernels, on every SM).  The on-chip kernels (PCG K1/K2, frontal D1/D2) keep their working set in shared memory, so this -- not HBM -- is
 * the roofline denominator bench.py reports them against.
 */
int tfin_smem_bandwidth(tfin_handle_t h, double* gbs_out);

/* Number of CUDA kernels this handle has launched since creation (bench.py's gpu_launches). */
int64_t tfin_kernel_launches(tfin_handle_t h);

/* Integer properties: "n", "n_obs", "n_terms", "n_r", "ell_width", "ell_width_nodal", "sm_count",
 * "pcg_threads", "pcg_rows_per_thread", "pcg_ctas_per_sm", "pcg_smem_bytes"; returns -1 if unknown. */
int64_t tfin_get_int(tfin_handle_t h, const char* key);

/* Knobs (before the next solve call): "pcg_rows_per_thread" (0 = auto), "rom_chunk", "pcg_path", "stream_tile" ...;
 * "fom_solver": 0 (default) = sparse-direct frontal Cholesky wherever the active front fits in shared memory, PCG
 * otherwise; 1 = always PCG (tol / maxit apply); 2 = direct required (fails if unavailable).  With the direct solver
 * tol / maxit are ignored, iters_out is 0 and relres_out is the consistency |b.w - y.y| / y.y of the two substitutions
 * (0 in the observables-only mode of the wide-front kernel, which has no backward substitution);
 * "frontal_kernel" (0 auto, 1 = sample per thread, 2 = sample per CTA), "frontal_threads", "frontal_mode",
 * "frontal_split", "frontal_lanes" (samples per warp of the sample-per-thread kernel: 0 = auto, the (warps per SM,
 * samples per warp) pair with the most samples in flight that shared memory allows; or 4..32; results do not depend on it);
 * "nodal_coef_mode": 0 = conductivity k (fom/forward_solve.py:160), 1 = conductivity exp(k) integrated with the
 * degree-3 rule of dolfin's form compiler (fom/forward_solve_exp.py:160) in tfin_fom_nodal / tfin_rom_nodal /
 * tfin_pcn_chains(model 0), and the gradient form k_hat exp(k) grad z . grad v with its degree-4 rule (:299, :328) in
 * tfin_fom_nodal_gradient / _sensitivity;
 * "pcg_precision": 64 (default) or 32 = opt-in single-precision on-chip PCG for tfin_fom_affine, n <= 2048 (operator
 * formed in fp64 and rounded, fp32 CG vectors, fp64 dot products / solution / observables).  Measured error floor of
 * the observables 3.6e-5 relative, 1.4x faster per iteration; for 1e-5 use the fp64 kernel with tol = 1e-9. */
int tfin_set_int(tfin_handle_t h, const char* key, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* TFIN_H */
