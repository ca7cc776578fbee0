// F: Gaussian-random-field conductivity sampler on the device.
//   covariance over the dof coordinates + Cholesky factor  = make_cov_chol (bayesian_inference/gaussian_field.py:9-31)
//   k_s = exp(0.5 * chol^T z_s), z_s ~ N(0, I_n)             = deep_learning/generate_fin_dataset.py:87-88
// The factor is kept as the LOWER triangle L (cov = L L^T, row-major), i.e. chol = L^T, so that
// (chol^T z)_j = sum_{i <= j} L[j][i] z_i is an "NT" product with both operands contiguous along the contraction.
#pragma once

#include "common.cuh"

namespace tfin {

// ------------------------------------------------------------------------------------------- F1 covariance
__global__ void __launch_bounds__(256) field_cov_kernel(const double* __restrict__ xy, int n, int kern, double length,
                                                        double* __restrict__ cov /* [n][n] */) {
    const long long total = (long long)n * n;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e / n), j = (int)(e - (long long)i * n);
        const double dx = xy[2 * i] - xy[2 * j], dy = xy[2 * i + 1] - xy[2 * j + 1];
        const double d = sqrt(dx * dx + dy * dy);
        double c;
        if (kern == TFIN_KERN_SQ_EXP) {
            c = exp(-(1.0 / (2.0 * length * length)) * (d * d)) + (i == j ? 1e-5 : 0.0);
        } else if (kern == TFIN_KERN_M52) {
            const double t = sqrt(5.0) * d / length;
            c = (1.0 + t + t * t / 3.0) * exp(-t);
        } else {
            const double t = sqrt(3.0) * d / length;
            c = (1.0 + t) * exp(-t);
        }
        cov[e] = c;
    }
}

// ------------------------------------------------------------------------------------------- F2 blocked Cholesky
// Right-looking, panel width 32, in place on the lower triangle of a row-major n x n matrix (one-time set-up work).
constexpr int FC_NB = 32;

// Diagonal block: one warp, lane = row, unblocked left-looking in registers/shared memory.
__global__ void __launch_bounds__(32) field_potf2_kernel(double* __restrict__ A, int n, int p0, int* __restrict__ info) {
    __shared__ double L[FC_NB][FC_NB + 1];
    const int lane = threadIdx.x, nb = min(FC_NB, n - p0);
    for (int j = 0; j < nb; ++j) L[lane][j] = (lane < nb && j <= lane) ? A[(size_t)(p0 + lane) * n + p0 + j] : 0.0;
    __syncwarp();
    for (int j = 0; j < nb; ++j) {
        double s = L[lane][j];
        for (int k = 0; k < j; ++k) s = fma(-L[lane][k], L[j][k], s);
        const double d = __shfl_sync(0xffffffffu, s, j);
        if (!(d > 0.0)) {
            if (lane == 0) atomicCAS(info, 0, p0 + j + 1);
            return;
        }
        const double ljj = sqrt(d);
        __syncwarp();
        if (lane >= j) L[lane][j] = (lane == j) ? ljj : s / ljj;
        __syncwarp();
    }
    for (int j = 0; j < nb; ++j)
        if (lane < nb) A[(size_t)(p0 + lane) * n + p0 + j] = (j <= lane) ? L[lane][j] : 0.0;  // zero the upper part
}

// Panel below the diagonal block: X L_pp^T = A_panel, one thread per row (forward substitution along the row).
__global__ void __launch_bounds__(128) field_trsm_kernel(double* __restrict__ A, int n, int p0) {
    __shared__ double L[FC_NB][FC_NB + 1];
    const int nb = min(FC_NB, n - p0);
    for (int e = threadIdx.x; e < FC_NB * FC_NB; e += blockDim.x) {
        const int r = e / FC_NB, c = e - r * FC_NB;
        L[r][c] = (r < nb && c <= r) ? A[(size_t)(p0 + r) * n + p0 + c] : 0.0;
    }
    __syncthreads();
    const int i = p0 + nb + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x[FC_NB];
#pragma unroll
    for (int j = 0; j < FC_NB; ++j) x[j] = j < nb ? A[(size_t)i * n + p0 + j] : 0.0;
#pragma unroll
    for (int j = 0; j < FC_NB; ++j) {
        if (j < nb) {
            double s = x[j];
#pragma unroll
            for (int k = 0; k < j; ++k) s = fma(-x[k], L[j][k], s);
            x[j] = s / L[j][j];
        }
    }
#pragma unroll
    for (int j = 0; j < FC_NB; ++j)
        if (j < nb) A[(size_t)i * n + p0 + j] = x[j];
}

// Trailing update A[i][j] -= sum_k P[i][k] P[j][k] for i >= j > panel; 32 x 32 tiles, lower tiles only.
__global__ void __launch_bounds__(256) field_syrk_kernel(double* __restrict__ A, int n, int p0) {
    __shared__ double Pi[FC_NB][FC_NB + 1], Pj[FC_NB][FC_NB + 1];
    const int t0 = p0 + FC_NB;
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj > bi) return;
    const int i0 = t0 + bi * FC_NB, j0 = t0 + bj * FC_NB;
    for (int e = threadIdx.x; e < FC_NB * FC_NB; e += 256) {
        const int r = e / FC_NB, c = e - r * FC_NB;
        Pi[r][c] = (i0 + r < n) ? A[(size_t)(i0 + r) * n + p0 + c] : 0.0;
        Pj[r][c] = (j0 + r < n) ? A[(size_t)(j0 + r) * n + p0 + c] : 0.0;
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // column tx, rows ty + 8 m
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int r = ty + 8 * m, i = i0 + r, j = j0 + tx;
        if (i >= n || j >= n || j > i) continue;
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < FC_NB; ++k) s = fma(Pi[r][k], Pj[tx][k], s);
        A[(size_t)i * n + j] -= s;
    }
}

// ------------------------------------------------------------------------------------------- F3 normals
// Philox4x32-10 counter RNG + Box-Muller, one independent stream per ROW (sample / chain) of the (N, n) array:
// the pair of entries (2p, 2p+1) of global row g draws counter (p, g_lo, g_hi, sub) under key (seed_lo, seed_hi), so a
// row's normals depend only on (seed, g, sub) -- not on the batch, chunk or rank that computes it.  `sub` numbers
// independent draws of the same row (e.g. the MCMC step).  The four 32-bit outputs give two 53-bit uniforms.
__host__ __device__ inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        c[1] = (uint32_t)p1;
        c[3] = (uint32_t)p0;
        c[0] = n0;
        c[2] = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

constexpr uint32_t PHILOX_PAIR_UNIFORM = 0xFFFFFFFFu;  // pair index reserved for the per-row accept/reject uniform

// two uniforms of (row g, pair p, draw sub): u1 in (0, 1], u2 in [0, 1)
__device__ __forceinline__ void philox_uniform2(unsigned long long seed, unsigned long long g, uint32_t p, uint32_t sub,
                                                double& u1, double& u2) {
    uint32_t c[4] = {p, (uint32_t)g, (uint32_t)(g >> 32), sub};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double two53 = 1.0 / 9007199254740992.0;
    u1 = ((double)((((uint64_t)c[0] << 32) | c[1]) >> 11) + 1.0) * two53;
    u2 = ((double)((((uint64_t)c[2] << 32) | c[3]) >> 11)) * two53;
}

__device__ __forceinline__ void philox_normal2(unsigned long long seed, unsigned long long g, uint32_t p, uint32_t sub,
                                               double& z0, double& z1) {
    double u1, u2, sn, cs;
    philox_uniform2(seed, g, p, sub, u1, u2);
    const double r = sqrt(-2.0 * log(u1));
    sincospi(2.0 * u2, &sn, &cs);
    z0 = r * cs;
    z1 = r * sn;
}

__global__ void __launch_bounds__(256) field_normal_kernel(unsigned long long seed, uint32_t sub, long long row0,
                                                           long long N, int n, double* __restrict__ z) {
    const int ppr = (n + 1) / 2;  // pairs per row
    const long long total = N * ppr;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / ppr;
        const int p = (int)(e - r * ppr);
        double z0, z1;
        philox_normal2(seed, (unsigned long long)(row0 + r), (uint32_t)p, sub, z0, z1);
        z[r * n + 2 * p] = z0;
        if (2 * p + 1 < n) z[r * n + 2 * p + 1] = z1;
    }
}

// ------------------------------------------------------------------------------------------- F4 sampler
// K[s][j] = exp(0.5 * sum_{i <= j} Z[s][i] L[j][i]):  fp64 "NT" GEMM restricted to the lower triangle, exp fused in
// the epilogue, on the FP64 tensor cores (DMMA m8n8k4).  128 x 64 tile, BK = 8, 8 warps x (32 x 32): 4 + 4 shared loads
// feed 16 MMAs per k-step of 4; operands are stored k-major in shared memory and the next k-slab is prefetched into
// registers while the current one is consumed.
constexpr int FS_BM = 128, FS_BN = 64, FS_BK = 8;

__global__ void __launch_bounds__(256) field_sample_kernel(const double* __restrict__ Z, long long N, int n,
                                                           const double* __restrict__ L, double* __restrict__ K) {
    __shared__ __align__(16) double As[2][FS_BK][FS_BM + 4], Bs[2][FS_BK][FS_BN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long s0 = (long long)blockIdx.y * FS_BM;
    const int j0 = blockIdx.x * FS_BN;
    const int kmax = min(n, j0 + FS_BN);  // L[j][i] = 0 for i > j
    const int nslab = (kmax + FS_BK - 1) / FS_BK;
    // loader mapping: A rows ar = tid / 2 (0..127), k quad (tid % 2) * 4;  B rows br = tid / 4 (0..63), k pair (tid % 4) * 2
    const int ar = tid >> 1, ak = (tid & 1) * 4, br = tid >> 2, bk = (tid & 3) * 2;
    double ra[4], rb[2];
    auto fetch = [&](int slab) {
        const int k0 = slab * FS_BK;
        const long long s = s0 + ar;
        const int j = j0 + br;
#pragma unroll
        for (int q = 0; q < 4; ++q) ra[q] = (s < N && k0 + ak + q < kmax) ? Z[s * n + k0 + ak + q] : 0.0;
#pragma unroll
        for (int q = 0; q < 2; ++q) rb[q] = (j < n && k0 + bk + q <= j) ? L[(size_t)j * n + k0 + bk + q] : 0.0;
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int q = 0; q < 4; ++q) As[buf][ak + q][ar] = ra[q];
#pragma unroll
        for (int q = 0; q < 2; ++q) Bs[buf][bk + q][br] = rb[q];
    };
    // warp (wm, wn) of the 4 x 2 warp grid owns 32 samples x 32 columns = 4 x 4 DMMA blocks; the row strides of As / Bs
    // (132, 68) are 4 (mod 16) doubles, so the one-element-per-lane fragment loads are bank-conflict free
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3, wm = warp >> 1, wn = warp & 1;
    double acc[4][4][2];
#pragma unroll
    for (int mb = 0; mb < 4; ++mb)
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
    fetch(0);
    stash(0);
    __syncthreads();
    for (int slab = 0; slab < nslab; ++slab) {
        const int buf = slab & 1;
        if (slab + 1 < nslab) fetch(slab + 1);
#pragma unroll
        for (int k0 = 0; k0 < FS_BK; k0 += 4) {
            double a[4], b[4];
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) a[mb] = As[buf][k0 + t][32 * wm + 8 * mb + g];
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) b[nb] = Bs[buf][k0 + t][32 * wn + 8 * nb + g];
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) dmma_884(acc[mb][nb][0], acc[mb][nb][1], a[mb], b[nb]);
        }
        if (slab + 1 < nslab) stash(buf ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) {
        const long long s = s0 + 32 * wm + 8 * mb + g;
        if (s >= N) continue;
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
            const int j = j0 + 32 * wn + 8 * nb + 2 * t;
            if (j < n) K[s * n + j] = exp(0.5 * acc[mb][nb][0]);
            if (j + 1 < n) K[s * n + j + 1] = exp(0.5 * acc[mb][nb][1]);
        }
    }
}

}  // namespace tfin
