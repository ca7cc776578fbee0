// K3: batched LSPG reduced-order solve (rom/averaged_affine_ROM.py:278-310, 323-333).
//
// Offline (host, once per basis):  Psi_t = V_t phi,  S_pq = sym(Psi_p^T Psi_q),  G_t = Psi_t^T b.
// Online, per sample with th_0 = 1:
//     A_r = sum_{p<=q} th_p th_q S_pq          (n_r x n_r, SPD)       \  one GEMM  C = Coef * S_aug
//     B_r = sum_t th_t G_t   ( = pairs (0,t) )                         /  (N x P2) . (P2 x Taug)
//     L L^T = A_r,  w_r = A_r^{-1} B_r,  qoi = (B_obs phi) w_r            one warp per sample
//
// Internal packed layout ("augmented lower, column-major"): column j holds rows i = j..n_r of the
// (n_r+1) x n_r matrix [A_r; B_r^T]:  off(j) = j (n_r+1) - j (j-1)/2,  Taug = T + n_r.  The extra row makes
// the forward substitution L y = B_r part of the factorisation (y is the last row of the factor).
#pragma once

#include "common.cuh"
#include "pcg_small.cuh"  // CsrRows

namespace tfin {

__host__ __device__ inline int rom_col_off(int j, int nr) { return j * (nr + 1) - (j * (j - 1)) / 2; }
__host__ __device__ inline int rom_taug(int nr) { return nr * (nr + 1) / 2 + nr; }

constexpr int ROM_BM = 64;    // samples per CTA
constexpr int ROM_BN = 128;   // packed entries per tile of the sweep
constexpr int ROM_LDA = ROM_BM + 4, ROM_LDB = ROM_BN + 4;  // row strides = 4 (mod 16): conflict-free fragment loads
constexpr int ROM_MAXP2 = TFIN_MAX_TERMS * (TFIN_MAX_TERMS + 1) / 2;

__host__ __device__ inline int rom_combine_k4(int n_terms) { return (n_terms * (n_terms + 1) / 2 + 3) & ~3; }
__host__ __device__ inline size_t rom_combine_smem(int n_terms) {
    return (size_t)rom_combine_k4(n_terms) * (ROM_LDA + 2 * ROM_LDB) * sizeof(double);
}

// ------------------------------------------------------------------------------------------- R1
// C[s][t] = sum_pq coef[s][pq] * S[pq][t],  coef[s][(p,q)] = th_p th_q -- the one real dense contraction of the path, on
// the FP64 tensor cores (mma.sync m8n8k4, the only FP64 tensor instruction of sm_100a).  One CTA owns 64 samples and
// sweeps ALL entry tiles (128 entries each): the coefficient tile is built once, the S tiles stream from L2 through a
// double buffer filled by cp.async while the previous tile is being multiplied, results are stored as they finish.
// Warp (wm, wn) of the 2 x 4 warp grid owns 32 samples x 32 entries = 4 x 4 DMMA blocks: per k-step of 4 it loads 4 + 4
// fragments for 16 MMAs (128 FMAs per lane), against 5 shared loads per 32 FMAs in the plain FMA version.
// Shared memory (P2 = 55 -> K = 56): coef [K][68] 30 KB + 2 x S tile [K][132] 118 KB.
__global__ void __launch_bounds__(256, 1) rom_combine_kernel(const double* __restrict__ theta,  // (N, n_terms-1)
                                                             long long s_begin, long long s_end, int n_terms,
                                                             const double* __restrict__ S,  // [P2][ldS], ldS % 128 == 0
                                                             int ldS, int Taug, double* __restrict__ C /* [Nchunk][Taug] */) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int P2 = n_terms * (n_terms + 1) / 2, K4 = rom_combine_k4(n_terms);
    double* s_coef = reinterpret_cast<double*>(smem);          // [K4][LDA]
    double* s_S = s_coef + (size_t)K4 * ROM_LDA;                // [2][K4][LDB]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;                    // 2 x 4 warps
    const int n_tiles = ldS / ROM_BN;
    const bool even = (Taug & 1) == 0;  // rows of C are then 16-byte aligned
    auto issue = [&](int tile) {
        if (tile < n_tiles) {
            double* dst = s_S + (size_t)(tile & 1) * K4 * ROM_LDB;
            const double* src = S + (size_t)tile * ROM_BN;
            for (int e = tid; e < P2 * (ROM_BN / 2); e += 256) {
                const int pq = e / (ROM_BN / 2), c2 = e - pq * (ROM_BN / 2);
                cp_async16(dst + pq * ROM_LDB + 2 * c2, src + (size_t)pq * ldS + 2 * c2);
            }
        }
        cp_async_commit();
    };
    for (int e = tid; e < (K4 - P2) * ROM_LDB * 2; e += 256) {  // zero the K padding rows of both S buffers, once
        const int bsel = e / ((K4 - P2) * ROM_LDB), r = e - bsel * (K4 - P2) * ROM_LDB;
        s_S[(size_t)bsel * K4 * ROM_LDB + (size_t)P2 * ROM_LDB + r] = 0.0;
    }
    for (long long s0 = s_begin + (long long)blockIdx.x * ROM_BM; s0 < s_end; s0 += (long long)gridDim.x * ROM_BM) {
        __syncthreads();  // previous sample tile is done with both buffers and s_coef
        issue(0);
        for (int sl = tid; sl < ROM_BM; sl += 256) {  // th_p th_q of one sample, straight from global memory
            const long long s = s0 + sl;
            double th[TFIN_MAX_TERMS];
            th[0] = s < s_end ? 1.0 : 0.0;
#pragma unroll
            for (int tt = 1; tt < TFIN_MAX_TERMS; ++tt)
                th[tt] = (tt < n_terms && s < s_end) ? theta[s * (n_terms - 1) + tt - 1] : 0.0;
            int pq = 0;
#pragma unroll
            for (int p = 0; p < TFIN_MAX_TERMS; ++p)
#pragma unroll
                for (int q = p; q < TFIN_MAX_TERMS; ++q)
                    if (p < n_terms && q < n_terms) s_coef[(pq++) * ROM_LDA + sl] = th[p] * th[q];
            for (int k = P2; k < K4; ++k) s_coef[k * ROM_LDA + sl] = 0.0;
        }
        for (int tile = 0; tile < n_tiles; ++tile) {
            cp_async_wait<0>();
            __syncthreads();       // tile landed (and s_coef written); everyone finished the other buffer
            issue(tile + 1);
            const double* sB = s_S + (size_t)(tile & 1) * K4 * ROM_LDB + 32 * wn + g;
            const double* sA = s_coef + 32 * wm + g;
            double acc[4][4][2];
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
#pragma unroll 2
            for (int k0 = 0; k0 < K4; k0 += 4) {
                double a[4], b[4];
#pragma unroll
                for (int mb = 0; mb < 4; ++mb) a[mb] = sA[(k0 + t) * ROM_LDA + 8 * mb];
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) b[nb] = sB[(k0 + t) * ROM_LDB + 8 * nb];
#pragma unroll
                for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                    for (int nb = 0; nb < 4; ++nb) dmma_884(acc[mb][nb][0], acc[mb][nb][1], a[mb], b[nb]);
            }
            const int t0 = tile * ROM_BN + 32 * wn + 2 * t;
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) {
                const long long s = s0 + 32 * wm + 8 * mb + g;
                if (s >= s_end) continue;
                double* row = C + (size_t)(s - s_begin) * Taug;
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) {
                    const int tt = t0 + 8 * nb;
                    if (even && tt + 1 < Taug) {
                        *reinterpret_cast<double2*>(row + tt) = make_double2(acc[mb][nb][0], acc[mb][nb][1]);
                    } else {
                        if (tt < Taug) row[tt] = acc[mb][nb][0];
                        if (tt + 1 < Taug) row[tt + 1] = acc[mb][nb][1];
                    }
                }
            }
        }
        cp_async_wait<0>();
    }
}

// ------------------------------------------------------------------------------------------- R2
// One warp per sample: panel-blocked left-looking Cholesky of the augmented packed matrix in shared memory, back
// substitution, observation projection.
// Adjoint outputs of the reduced gradient (averaged_affine_ROM.py:335-346): v_r = A_r^{-T} (B_obs phi)^T (data - qoi).
struct RomAdj {
    const double* data;   // (1 | N, n_obs)
    long long data_stride;  // 0 = one observation vector for every sample
    double* vr_out;       // (N, n_r)
    double* cost_out;     // (N) | nullptr: J = 0.5 ||data - qoi||^2
};

// Back substitution L^T w = y on the packed factor.  y is lane-distributed in registers (lane l holds entries
// l + 32 m); step j broadcasts w_j by shuffle and every lane updates its entries i < j with L[j][i] (column i of
// the packed factor, offset j - i), loaded ahead of the dependent chain.
template <int MAXM>
__device__ __forceinline__ void rom_back_subst(const double* __restrict__ A, const double* __restrict__ dinv, int nr,
                                               int lane, double (&yv)[MAXM]) {
    // the slab index of the pivot is a compile-time constant inside the unrolled mb loop, so yv stays in registers and
    // the triangle tests reduce to one lane compare in the pivot's own slab (rows of lower slabs are always above it)
    int ci[MAXM];  // rom_col_off(i) - i of this lane's rows: + j addresses L[j][i]
#pragma unroll
    for (int m = 0; m < MAXM; ++m) {
        const int i = min(lane + 32 * m, nr - 1);
        ci[m] = rom_col_off(i, nr) - i;
    }
#pragma unroll
    for (int mb = MAXM - 1; mb >= 0; --mb) {
        for (int j = min(nr, 32 * mb + 32) - 1; j >= 32 * mb; --j) {
            const int jl = j & 31;
            double lji[MAXM];
#pragma unroll
            for (int m = 0; m <= mb; ++m) lji[m] = (m < mb || lane < jl) ? A[ci[m] + j] : 0.0;
            const double wj = __shfl_sync(0xffffffffu, yv[mb], jl) * dinv[j];
#pragma unroll
            for (int m = 0; m < mb; ++m) yv[m] = fma(-lji[m], wj, yv[m]);
            yv[mb] = (lane == jl) ? wj : fma(-lji[mb], wj, yv[mb]);
        }
    }
}

// Forward substitution L z = y (column sweep: column j of the packed factor is contiguous, so the lanes read
// consecutive words); same register distribution as rom_back_subst.
template <int MAXM>
__device__ __forceinline__ void rom_fwd_subst(const double* __restrict__ A, const double* __restrict__ dinv, int nr,
                                              int lane, double (&yv)[MAXM]) {
#pragma unroll
    for (int mb = 0; mb < MAXM; ++mb) {
        for (int j = 32 * mb; j < min(nr, 32 * mb + 32); ++j) {
            const int jl = j & 31;
            const double* colp = A + (rom_col_off(j, nr) - j) + lane;  // + 32 m addresses L[lane + 32 m][j]
            double lij[MAXM];
#pragma unroll
            for (int m = mb; m < MAXM; ++m)
                lij[m] = ((m > mb || lane > jl) && lane + 32 * m < nr) ? colp[32 * m] : 0.0;
            const double zj = __shfl_sync(0xffffffffu, yv[mb], jl) * dinv[j];
            yv[mb] = (lane == jl) ? zj : fma(-lij[mb], zj, yv[mb]);
#pragma unroll
            for (int m = mb + 1; m < MAXM; ++m) yv[m] = fma(-lij[m], zj, yv[m]);
        }
    }
}

// Left-looking update of an 8-column panel on the FP64 tensor cores, in place in the packed factor:
//     A[r][j0 + c] -= sum_{k < j0} L[r][k] L[j0 + c][k],      r = j0 .. n_r (augmented row included), c = 0 .. 7
// i.e. (rows x j0) . (j0 x 8) as DMMA m8n8k4 blocks: lane (g, t) supplies L[j0 + 8 mb + g][k0 + t] as the A fragment of
// row block mb and L[j0 + g][k0 + t] as the B fragment (the same word as its A fragment of block 0), so one k-step of
// 4 costs 4 G shared loads and 4 G MMAs for 32 G rows x 8 columns x 4 k = 1024 G FMAs (the FMA sweep needed 38 G + 40
// instructions for the same work).  Column k of the packed layout starts at k n_r - k (k - 1) / 2 - k relative to its
// row index, advanced incrementally.  Loads are NOT bounds-checked (rows past the matrix read at most 31 doubles
// beyond a column, inside this warp's region); the read-modify-write at the end is masked to the stored triangle.
template <int G /* live 32-row groups */>
__device__ __forceinline__ void rom_chol_sweep_mma(double* __restrict__ A, int j0, int nr, int lane) {
    const int g = lane >> 2, t = lane & 3;
    double acc[4 * G][2];
#pragma unroll
    for (int mb = 0; mb < 4 * G; ++mb) acc[mb][0] = acc[mb][1] = 0.0;
    int k = t, base = t * nr - (t * (t - 1)) / 2;  // rom_col_off(k) - k: + row addresses L[row][k]
#pragma unroll 2
    for (int k0 = 0; k0 < j0; k0 += 4) {
        const double* colp = A + base + j0 + g;
        double a[4 * G];
#pragma unroll
        for (int mb = 0; mb < 4 * G; ++mb) a[mb] = colp[8 * mb];
#pragma unroll
        for (int mb = 0; mb < 4 * G; ++mb) dmma_884(acc[mb][0], acc[mb][1], a[mb], a[0]);
        base += 4 * nr - 4 * k - 6;
        k += 4;
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int col = j0 + 2 * t + e;
        if (col < nr) {
            const int oc = rom_col_off(col, nr) - col;
#pragma unroll
            for (int mb = 0; mb < 4 * G; ++mb) {
                const int row = j0 + 8 * mb + g;
                if (row <= nr && row >= col) A[oc + row] -= acc[mb][e];
            }
        }
    }
}

// One panel (NB columns starting at j0) of the left-looking Cholesky of the augmented packed matrix, for a warp.  Lane l
// holds rows j0 + l + 32 m (m < M live slabs) of the panel columns in registers.  With SWEEP the contribution of the
// finished columns k < j0 is applied here with FMAs (M + NB shared loads for NB M FMAs per k); without it the caller
// has already applied it (rom_chol_sweep_mma).  Then the panel is factored in registers with shuffles and written
// back once.
template <int M, int NB, bool SWEEP>
__device__ __forceinline__ void rom_chol_panel(double* __restrict__ A, double* __restrict__ dinv, int j0, int nr,
                                               int lane, int& status) {
    const int nrow = nr + 1, ncol = min(NB, nr - j0);
    double c[NB][M];
    bool row_ok[M];  // row j0 + lane + 32 m exists; "on or below the diagonal of column cc" is lane >= cc in slab 0 only
#pragma unroll
    for (int m = 0; m < M; ++m) row_ok[m] = j0 + lane + 32 * m < nrow;
    int oc[NB];      // + lane + 32 m addresses row j0 + lane + 32 m of column j0 + cc
#pragma unroll
    for (int cc = 0; cc < NB; ++cc) oc[cc] = rom_col_off(min(j0 + cc, nr - 1), nr) - cc;
#pragma unroll
    for (int cc = 0; cc < NB; ++cc)
#pragma unroll
        for (int m = 0; m < M; ++m)
            c[cc][m] = (cc < ncol && row_ok[m] && (m > 0 || lane >= cc)) ? A[oc[cc] + lane + 32 * m] : 0.0;
    if (SWEEP) {
        // colp walks row (j0 + lane) of column k: consecutive columns of the packed layout are nr - k entries apart.
        // Unchecked loads, see rom_chol_sweep_mma.
        const double* colp = A + j0 + lane;
        int stride = nr;
#pragma unroll 4
        for (int k = 0; k < j0; ++k) {
            double lk[M], lj[NB];
#pragma unroll
            for (int m = 0; m < M; ++m) lk[m] = colp[32 * m];
            const double* rowp = colp - lane;  // row j0 of column k (warp-uniform address: broadcast)
#pragma unroll
            for (int cc = 0; cc < NB; ++cc) lj[cc] = rowp[cc];
#pragma unroll
            for (int m = 0; m < M; ++m)
#pragma unroll
                for (int cc = 0; cc < NB; ++cc) c[cc][m] = fma(-lk[m], lj[cc], c[cc][m]);
            colp += stride;
            --stride;
        }
    }
#pragma unroll
    for (int cc = 0; cc < NB; ++cc) {
        if (cc < ncol) {
            // updates from the already finished columns of this panel
#pragma unroll
            for (int c2 = 0; c2 < NB; ++c2) {
                if (c2 < cc) {
                    const double ljc = __shfl_sync(0xffffffffu, c[c2][0], cc);  // L[j0+cc][j0+c2]
#pragma unroll
                    for (int m = 0; m < M; ++m) c[cc][m] = fma(-c[c2][m], ljc, c[cc][m]);
                }
            }
            const double d = __shfl_sync(0xffffffffu, c[cc][0], cc);
            if (!(d > 0.0)) status = TFIN_STATUS_BREAKDOWN;
            // sqrt(d) and 1/sqrt(d) from ONE reciprocal square root (the serial chain of the factorisation runs through
            // here 81 times): ljj = d r corrected by its exact residual, then one Newton step on 1 / ljj
            const double inv0 = rsqrt(d);
            double ljj = d * inv0;
            ljj = fma(fma(-ljj, ljj, d), 0.5 * inv0, ljj);
            const double invl = inv0 * (2.0 - ljj * inv0);
#pragma unroll
            for (int m = 0; m < M; ++m) {
                c[cc][m] = (m == 0 && lane == cc) ? ljj : c[cc][m] * invl;
                if (row_ok[m] && (m > 0 || lane >= cc)) A[oc[cc] + lane + 32 * m] = c[cc][m];
            }
            if (lane == 0) dinv[j0 + cc] = invl;
        }
    }
}

template <int MAXM /* ceil((n_r+1)/32) */, bool ADJ = false>
__global__ void __launch_bounds__(256, 1) rom_chol_kernel(const double* __restrict__ C, long long s_begin,
                                                       long long s_end, int nr, int n_obs,
                                                       const double* __restrict__ obs_phi,  // [n_obs][nr]
                                                       double* __restrict__ wr_out, double* __restrict__ qoi_out,
                                                       int* __restrict__ status_out, RomAdj adj) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int Taug = rom_taug(nr);
    const int wpb = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per_warp = ((Taug + 2 * nr + 2) + 1) & ~1;           // A | dinv[nr] | w[nr+1]
    double* A = reinterpret_cast<double*>(smem) + (size_t)warp * per_warp;
    double* dinv = A + Taug;
    double* wv = dinv + nr;
    const int nrow = nr + 1;
    // The observation rows (B_obs phi) of this lane's entries, read once per warp instead of once per sample: the
    // projection then works on the register-distributed solution and never touches memory.
    constexpr int NOBS_REG = 9;
    const bool obs_reg = n_obs <= NOBS_REG;
    double op[NOBS_REG][MAXM];
#pragma unroll
    for (int o = 0; o < NOBS_REG; ++o)
#pragma unroll
        for (int m = 0; m < MAXM; ++m)
            op[o][m] = (obs_reg && o < n_obs && lane + 32 * m < nr) ? obs_phi[o * nr + lane + 32 * m] : 0.0;
    // The packed matrix arrives in four cp.async groups cut at columns 8 / 24 / 48: the left-looking factorisation of
    // panel j0 needs columns < j0 + 8 only, i.e. a prefix of the packed array, so it starts after the first ~18 % have
    // landed and the rest of the load hides behind the first panels.
    const bool staged = (Taug & 1) == 0;   // rows of C are 16-byte aligned when Taug is even (per_warp is)
    int cut[5];                            // group g = double2 elements [cut[g], cut[g + 1])
    cut[0] = 0;
    cut[1] = rom_col_off(min(8, nr), nr) >> 1;
    cut[2] = rom_col_off(min(24, nr), nr) >> 1;
    cut[3] = rom_col_off(min(48, nr), nr) >> 1;
    cut[4] = Taug >> 1;
    const unsigned A_s = smem_u32(A);

    for (long long s = s_begin + (long long)blockIdx.x * wpb + warp; s < s_end;
         s += (long long)gridDim.x * wpb) {
        const double* src = C + (size_t)(s - s_begin) * Taug;
        if (staged) {
            const double2* src2 = reinterpret_cast<const double2*>(src);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                for (int e = cut[g] + lane; e < cut[g + 1]; e += 32)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(A_s + 16u * (unsigned)e), "l"(src2 + e) : "memory");
                cp_async_commit();
            }
        } else {
            for (int e = lane; e < Taug; e += 32) A[e] = ldg_stream(src + e);
            __syncwarp();
        }
        int status = TFIN_STATUS_CONVERGED;
        // Panel-blocked left-looking Cholesky, 8 columns at a time: the update from the finished columns runs on the
        // tensor cores (DMMA), the panel itself is factored in registers.  Both are instantiated for the number of
        // 32-row slabs that still hold rows (warp-uniform), so empty slabs cost no issue slots at all.
        for (int j0 = 0; j0 < nr; j0 += 8) {
            if (staged) {   // wait for the groups that hold columns < j0 + 8 (an element that straddles a cut counts as late)
                const int need = (rom_col_off(min(j0 + 8, nr), nr) + 1) >> 1;
                if (need > cut[3]) cp_async_wait<0>();
                else if (need > cut[2]) cp_async_wait<1>();
                else if (need > cut[1]) cp_async_wait<2>();
                else cp_async_wait<3>();
                __syncwarp();
            }
            const int mact = min(MAXM, (nrow - j0 + 31) >> 5);
            if (MAXM >= 4 && mact == 4) {
                if (j0) rom_chol_sweep_mma<(MAXM >= 4 ? 4 : 1)>(A, j0, nr, lane);
                __syncwarp();
                rom_chol_panel<(MAXM >= 4 ? 4 : 1), 8, false>(A, dinv, j0, nr, lane, status);
            } else if (MAXM >= 3 && mact == 3) {
                if (j0) rom_chol_sweep_mma<(MAXM >= 3 ? 3 : 1)>(A, j0, nr, lane);
                __syncwarp();
                rom_chol_panel<(MAXM >= 3 ? 3 : 1), 8, false>(A, dinv, j0, nr, lane, status);
            } else if (MAXM >= 2 && mact == 2) {
                if (j0) rom_chol_sweep_mma<(MAXM >= 2 ? 2 : 1)>(A, j0, nr, lane);
                __syncwarp();
                rom_chol_panel<(MAXM >= 2 ? 2 : 1), 8, false>(A, dinv, j0, nr, lane, status);
            } else {
                if (j0) rom_chol_sweep_mma<1>(A, j0, nr, lane);
                __syncwarp();
                rom_chol_panel<1, 8, false>(A, dinv, j0, nr, lane, status);
            }
            __syncwarp();
        }
        // y = last row of the factor (the forward substitution rode along with the factorisation)
        double yv[MAXM];
#pragma unroll
        for (int m = 0; m < MAXM; ++m) {
            const int i = lane + 32 * m;
            yv[m] = i < nr ? A[rom_col_off(i, nr) + (nr - i)] : 0.0;
        }
        rom_back_subst<MAXM>(A, dinv, nr, lane, yv);
#pragma unroll
        for (int m = 0; m < MAXM; ++m) {
            const int i = lane + 32 * m;
            if (i < nr) wv[i] = yv[m];
        }
        __syncwarp();
        if (status_out && lane == 0) status_out[s] = status;
        if (wr_out)
            for (int j = lane; j < nr; j += 32) wr_out[s * nr + j] = wv[j];
        if (qoi_out || ADJ) {
            double rv[MAXM], cost = 0.0;
#pragma unroll
            for (int m = 0; m < MAXM; ++m) rv[m] = 0.0;
            if (obs_reg) {
#pragma unroll
                for (int o = 0; o < NOBS_REG; ++o) {
                    if (o >= n_obs) break;
                    double acc = 0.0;
#pragma unroll
                    for (int m = 0; m < MAXM; ++m)   // yv: the solution, lane-distributed (entries past n_r hold junk)
                        acc = fma(op[o][m], lane + 32 * m < nr ? yv[m] : 0.0, acc);
                    acc = warp_sum(acc);
                    if (qoi_out && lane == 0) qoi_out[s * n_obs + o] = acc;
                    if (ADJ) {  // reduced adjoint right-hand side (B_obs phi)^T (data - qoi), :338-339
                        const double res = adj.data[s * adj.data_stride + o] - acc;
                        cost = fma(res, res, cost);
#pragma unroll
                        for (int m = 0; m < MAXM; ++m) rv[m] = fma(op[o][m], res, rv[m]);
                    }
                }
            } else {
                for (int o = 0; o < n_obs; ++o) {
                    double acc = 0.0;
                    for (int j = lane; j < nr; j += 32) acc = fma(obs_phi[o * nr + j], wv[j], acc);
                    acc = warp_sum(acc);
                    if (qoi_out && lane == 0) qoi_out[s * n_obs + o] = acc;
                    if (ADJ) {  // reduced adjoint right-hand side (B_obs phi)^T (data - qoi), :338-339
                        const double res = adj.data[s * adj.data_stride + o] - acc;
                        cost = fma(res, res, cost);
#pragma unroll
                        for (int m = 0; m < MAXM; ++m) {
                            const int j = lane + 32 * m;
                            if (j < nr) rv[m] = fma(obs_phi[o * nr + j], res, rv[m]);
                        }
                    }
                }
            }
            if (ADJ) {  // v_r = A_r^{-T} rhs = L^{-T} L^{-1} rhs (A_r symmetric), :345
                rom_fwd_subst<MAXM>(A, dinv, nr, lane, rv);
                rom_back_subst<MAXM>(A, dinv, nr, lane, rv);
#pragma unroll
                for (int m = 0; m < MAXM; ++m) {
                    const int j = lane + 32 * m;
                    if (j < nr) adj.vr_out[s * nr + j] = rv[m];
                }
                if (adj.cost_out && lane == 0) adj.cost_out[s] = 0.5 * cost;
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------- R3
// Reduced gradient contraction (averaged_affine_ROM.py:347-351 rewritten with offline Gram blocks):
//     g_q = (psi v_r)^T (K_q phi) w_r = sum_t th_t  v_r^T N_tq w_r,     N_tq = Psi_t^T Psi_q   (n_r x n_r)
// as a batched GEMM  c[s][o] = sum_{i,j} (v_i w_j) NG[(i,j)][o],  o = (t, q), followed by the th_t fold -- on the FP64
// tensor cores (DMMA m8n8k4).  The A operand is never stored: lane (g, t) of an 8 x 4 fragment forms
// v[s_g][i] * w[s_g][j0 + t] with one multiply from the transposed, padded copies of w and v in shared memory.  The j
// range is padded to a multiple of 4 (zero rows of NG) so that a k-step of 4 never straddles two i.
// CTA = 64 samples x 96 outputs at a time; warp (wm, wn) of the 2 x 4 grid owns 32 samples x 24 outputs = 4 x 3 DMMA
// blocks.  NG streams from L2 through a 3-stage cp.async ring of 32-row chunks (row stride 100: conflict-free
// fragment loads).
constexpr int RG_BM = 64, RG_OB = 96, RG_KC = 32, RG_STAGES = 3, RG_LDB = RG_OB + 4, RG_LDA = RG_BM + 4;

__host__ __device__ inline int rom_grad_jpad(int nr) { return (nr + 3) & ~3; }
__host__ __device__ inline size_t rom_grad_smem(int nr, int n_par) {
    const size_t ring = (size_t)RG_STAGES * RG_KC * RG_LDB, cbuf = (size_t)RG_BM * RG_OB;
    return ((size_t)(rom_grad_jpad(nr) + nr) * RG_LDA + (ring > cbuf ? ring : cbuf) + (size_t)RG_BM * n_par) * sizeof(double);
}

__global__ void __launch_bounds__(256, 1) rom_grad_kernel(const double* __restrict__ theta,  // (N, n_terms-1)
                                                          const double* __restrict__ wr,     // (N, nr)
                                                          const double* __restrict__ vr,     // (N, nr)
                                                          long long N, int nr, int n_terms,
                                                          const double* __restrict__ NG,  // [n_ob][nr*jpad][RG_OB]
                                                          int n_ob, double* __restrict__ g_out /* (N, n_terms-1) */) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int n_par = n_terms - 1, O = n_terms * n_par, jp = rom_grad_jpad(nr), K = nr * jp;
    double* s_w = reinterpret_cast<double*>(smem);     // [jp][LDA]  (rows >= nr are zero)
    double* s_v = s_w + (size_t)jp * RG_LDA;            // [nr][LDA]
    double* s_ring = s_v + (size_t)nr * RG_LDA;         // [STAGES][KC][LDB]; reused as c[BM][OB] in the epilogue
    const size_t ring_sz = (size_t)RG_STAGES * RG_KC * RG_LDB, cbuf_sz = (size_t)RG_BM * RG_OB;
    double* s_g = s_ring + (ring_sz > cbuf_sz ? ring_sz : cbuf_sz);  // [BM][n_par]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    const int n_chunks = (K + RG_KC - 1) / RG_KC;

    for (long long s0 = (long long)blockIdx.x * RG_BM; s0 < N; s0 += (long long)gridDim.x * RG_BM) {
        __syncthreads();  // previous tile's epilogue is done with s_ring / s_g
        for (int e = tid; e < RG_BM * jp; e += 256) {
            const int sl = e / jp, j = e - sl * jp;
            const long long s = s0 + sl;
            const bool ok = s < N && j < nr;
            s_w[j * RG_LDA + sl] = ok ? wr[s * nr + j] : 0.0;
            if (j < nr) s_v[j * RG_LDA + sl] = ok ? vr[s * nr + j] : 0.0;
        }
        for (int e = tid; e < RG_BM * n_par; e += 256) s_g[e] = 0.0;
        for (int ob = 0; ob < n_ob; ++ob) {
            const double* src = NG + (size_t)ob * K * RG_OB;
            auto issue = [&](int chunk) {
                if (chunk < n_chunks) {
                    const int rows = min(RG_KC, K - chunk * RG_KC);
                    const double* gsrc = src + (size_t)chunk * RG_KC * RG_OB;
                    double* d = s_ring + (size_t)(chunk % RG_STAGES) * RG_KC * RG_LDB;
                    for (int e = tid; e < rows * (RG_OB / 2); e += 256) {
                        const int r = e / (RG_OB / 2), c2 = e - r * (RG_OB / 2);
                        cp_async16(d + r * RG_LDB + 2 * c2, gsrc + (size_t)r * RG_OB + 2 * c2);
                    }
                }
                cp_async_commit();
            };
            __syncthreads();  // s_w / s_v visible; ring free
            for (int c = 0; c < RG_STAGES - 1; ++c) issue(c);
            double acc[4][3][2];
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb = 0; nb < 3; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
            const double* sA = s_w + 32 * wm + g;
            const double* sV = s_v + 32 * wm + g;
            int i = 0, j0 = 0;
            double vi[4];
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) vi[mb] = sV[8 * mb];   // v[.][i = 0]; s_v is visible after the barrier above
            for (int chunk = 0; chunk < n_chunks; ++chunk) {
                cp_async_wait<RG_STAGES - 2>();
                __syncthreads();             // chunk landed for everyone; the stage refilled below was consumed
                issue(chunk + RG_STAGES - 1);
                const double* sB = s_ring + (size_t)(chunk % RG_STAGES) * RG_KC * RG_LDB + 24 * wn + g;
                const int rows = min(RG_KC, K - chunk * RG_KC);   // a multiple of 4
#pragma unroll 2
                for (int r = 0; r < rows; r += 4) {
                    double a[4], b[3];
#pragma unroll
                    for (int mb = 0; mb < 4; ++mb) a[mb] = vi[mb] * sA[(j0 + t) * RG_LDA + 8 * mb];
#pragma unroll
                    for (int nb = 0; nb < 3; ++nb) b[nb] = sB[(r + t) * RG_LDB + 8 * nb];
#pragma unroll
                    for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                        for (int nb = 0; nb < 3; ++nb) dmma_884(acc[mb][nb][0], acc[mb][nb][1], a[mb], b[nb]);
                    j0 += 4;
                    if (j0 == jp) {  // next row i of the outer product
                        j0 = 0;
                        ++i;
                        if (i < nr) {
#pragma unroll
                            for (int mb = 0; mb < 4; ++mb) vi[mb] = sV[i * RG_LDA + 8 * mb];
                        }
                    }
                }
            }
            cp_async_wait<0>();
            __syncthreads();  // everyone is done with the ring: reuse it for c[BM][OB]
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb = 0; nb < 3; ++nb) {
                    double* dst = s_ring + (32 * wm + 8 * mb + g) * RG_OB + 24 * wn + 8 * nb + 2 * t;
                    dst[0] = acc[mb][nb][0];
                    dst[1] = acc[mb][nb][1];
                }
            __syncthreads();
            for (int e = tid; e < RG_BM * n_par; e += 256) {
                const int sl = e / n_par, q = e - sl * n_par;
                const long long s = s0 + sl;
                if (s >= N) continue;
                double gq = 0.0;
                // outputs of this block: o = ob*OB + ol = t * n_par + q
                for (int tt = 0; tt < n_terms; ++tt) {
                    const int ol = tt * n_par + q - ob * RG_OB;
                    if (ol < 0 || ol >= RG_OB || tt * n_par + q >= O) continue;
                    const double th = tt == 0 ? 1.0 : theta[s * n_par + tt - 1];
                    gq = fma(th, s_ring[sl * RG_OB + ol], gq);
                }
                s_g[e] += gq;
            }
        }
        __syncthreads();
        for (int e = tid; e < RG_BM * n_par; e += 256) {
            const long long s = s0 + e / n_par;
            if (s < N) g_out[s0 * n_par + e] = s_g[e];
        }
    }
}

// dJ/dk = g^T dsigma_dk (averaged_affine_ROM.py:350-351): out[s][i] = sum_q g[s][q] Avg[q][i], through the transposed
// averaging operator (CSR over the n dofs).  One thread per (sample, dof), coalesced stores.
__global__ void __launch_bounds__(256) rom_grad_lift_kernel(CsrRows avgT, const double* __restrict__ g, long long N,
                                                            int n_par, double* __restrict__ out) {
    const long long total = N * avgT.rows;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long s = e / avgT.rows;
        const int i = (int)(e - s * avgT.rows);
        double acc = 0.0;
        for (int j = avgT.ptr[i]; j < avgT.ptr[i + 1]; ++j) acc = fma(avgT.val[j], g[s * n_par + avgT.idx[j]], acc);
        out[e] = acc;
    }
}

}  // namespace tfin
