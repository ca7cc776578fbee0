// K4: batched Jacobi-PCG for meshes too large for one SM (refined meshes, n ~ 1e5): a genuinely HBM-streaming
// kernel.  Included by tfin_api.cu only.
//
//   * a TILE of S samples is solved by ONE persistent CTA (1024 threads); tiles are pulled from a global counter,
//     so there is no grid-wide synchronisation at all: every reduction is CTA-local
//   * the CG vectors x, r, p, q of a tile live in HBM in an interleaved layout  v[row][S]  (sample fastest): a lane
//     owns 2 adjacent samples of one row and moves them with 16-byte loads/stores, a warp covers 64/S rows, so
//     every access of a warp is made of contiguous 16*S/2-byte segments (fully coalesced, sector aligned)
//   * the operator is applied MATRIX-FREE: the shared term-tagged CSR  (col, term, coef)  is read once per row for
//     all S samples (warp-uniform loads that hit L1/L2) and combined with the per-sample conductivities
//         (A(theta) p)_i = sum_e theta[term_e] * coef_e * p[col_e]          (averaged_affine_ROM.py:156-162)
//     so no per-sample matrix values exist anywhere; 1/diag is recomputed from the term list as well
//   * standard 3-pass PCG per iteration, algorithmic HBM traffic per sample:
//         P1  q = A p, delta = p.q                     read p, write q            16 n bytes
//         P2  x += a p, r -= a q, gamma' = r.D^-1 r    read x,p,q,r write x,r     48 n bytes
//         P3  p = D^-1 r + b p                         read r,p write p           24 n bytes   => 88 n bytes/iter
//   * converged samples of a tile are frozen (alpha = beta = 0) until the whole tile is done
#pragma once

#include "pcg_small.cuh"

namespace tfin {

struct StreamOp {
    int n, n_terms;
    const int* row_ptr;    // [n+1]
    const int2* ent;       // [nnz_t]  (col, term)
    const double* coef;    // [nnz_t]
    const int* dptr;       // [n+1]    diagonal (term, coef) list
    const int* dterm;      // [dnnz]
    const double* dcoef;   // [dnnz]
    const double* rhs;     // [n]
};

template <int S>
__global__ void __launch_bounds__(1024, 1) pcg_stream_kernel(StreamOp op, CsrRows obs, PcgIO io,
                                                             double* __restrict__ work) {
    static_assert(S == 8 || S == 16 || S == 32, "tile width");
    constexpr int LPR = S / 2;     // lanes per row (each lane owns 2 samples)
    constexpr int RPW = 32 / LPR;  // rows per warp
    __shared__ double s_theta[TFIN_MAX_TERMS * S];
    __shared__ double s_red[2][32][S];  // per-warp partial sums
    __shared__ double s_tot[2][S];
    __shared__ double s_alpha[S], s_beta[S], s_gamma[S], s_thresh[S];
    __shared__ int s_active[S], s_iters[S], s_status[S], s_nactive;
    __shared__ long long s_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int n = op.n;
    const int sp = lane % LPR, rw = lane / LPR;  // sample pair, row within the warp's row group
    const size_t vec = (size_t)n * S;
    double2* x2 = reinterpret_cast<double2*>(work + (size_t)blockIdx.x * 4 * vec);
    double2* r2 = x2 + vec / 2;
    double2* p2 = r2 + vec / 2;
    double2* q2 = p2 + vec / 2;
    const double2* th2 = reinterpret_cast<const double2*>(s_theta);
    const long long n_tiles = (io.N + S - 1) / S;
    const int n_groups = (n + RPW - 1) / RPW;

    // lane-local partials -> per-sample CTA totals in s_tot[which][S].  Lanes with equal sp (same samples) are
    // first folded across the RPW rows of the warp, then across warps in fixed order.
    auto reduce2 = [&](double2 a, double2 b) {
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) {
            a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
            a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
            b.x += __shfl_xor_sync(0xffffffffu, b.x, o);
            b.y += __shfl_xor_sync(0xffffffffu, b.y, o);
        }
        if (lane < LPR) {
            s_red[0][warp][2 * sp] = a.x;
            s_red[0][warp][2 * sp + 1] = a.y;
            s_red[1][warp][2 * sp] = b.x;
            s_red[1][warp][2 * sp + 1] = b.y;
        }
        __syncthreads();
        if (tid < 2 * S) {
            const int which = tid / S, s = tid % S;
            double t = 0.0;
            for (int w = 0; w < nwarps; ++w) t += s_red[which][w][s];
            s_tot[which][s] = t;
        }
        __syncthreads();
    };
    // 1/diag of row i for this lane's two samples, recomputed from the shared term list
    auto dinv_of = [&](int i) -> double2 {
        double2 d = make_double2(0.0, 0.0);
        for (int e = op.dptr[i]; e < op.dptr[i + 1]; ++e) {
            const double c = op.dcoef[e];
            const double2 th = th2[op.dterm[e] * LPR + sp];
            d.x = fma(th.x, c, d.x);
            d.y = fma(th.y, c, d.y);
        }
        return make_double2(1.0 / d.x, 1.0 / d.y);
    };
    // (A v)_i for this lane's two samples
    auto apply_row = [&](int i, bool valid, const double2* __restrict__ v) -> double2 {
        double2 acc = make_double2(0.0, 0.0);
        const int e0 = valid ? op.row_ptr[i] : 0, e1 = valid ? op.row_ptr[i + 1] : 0;
        const int cnt = __reduce_max_sync(0xffffffffu, e1 - e0);
        for (int k = 0; k < cnt; ++k) {
            if (e0 + k < e1) {
                const int2 ct = op.ent[e0 + k];
                const double c = op.coef[e0 + k];
                const double2 th = th2[ct.y * LPR + sp];
                const double2 pv = v[(size_t)ct.x * LPR + sp];
                acc.x = fma(th.x * c, pv.x, acc.x);
                acc.y = fma(th.y * c, pv.y, acc.y);
            }
        }
        return acc;
    };

    for (;;) {
        __syncthreads();
        if (tid == 0) s_tile = (long long)atomicAdd(io.counter, 1ULL);
        __syncthreads();
        const long long tile = s_tile;
        if (tile >= n_tiles) break;
        const long long s_base = tile * S;
        // ---- per-sample conductivities (theta_0 = 1); dummy samples of a partial tile get theta = 1, inactive
        for (int e = tid; e < op.n_terms * S; e += blockDim.x) {
            const int t = e / S, s = e % S;
            const long long smp = s_base + s;
            s_theta[e] = (t == 0 || smp >= io.N) ? 1.0 : io.in[smp * io.in_stride + t - 1];
        }
        if (tid < S) {
            s_active[tid] = (s_base + tid < io.N) ? 1 : 0;
            s_iters[tid] = 0;
            s_status[tid] = TFIN_STATUS_MAXIT;
        }
        __syncthreads();

        // ---- init: x = 0, r = b, p = D^-1 r, gamma = r.D^-1 r
        double2 g_acc = make_double2(0.0, 0.0), zero2 = make_double2(0.0, 0.0);
        for (int g = warp; g < n_groups; g += nwarps) {
            const int i = g * RPW + rw;
            if (i < n) {
                const double b = op.rhs[i];
                const double2 di = dinv_of(i);
                const size_t o = (size_t)i * LPR + sp;
                x2[o] = zero2;
                r2[o] = make_double2(b, b);
                p2[o] = make_double2(di.x * b, di.y * b);
                g_acc.x = fma(b * di.x, b, g_acc.x);
                g_acc.y = fma(b * di.y, b, g_acc.y);
            }
        }
        reduce2(g_acc, zero2);
        if (tid < S) {
            const double g0 = s_tot[0][tid];
            s_gamma[tid] = g0;
            s_thresh[tid] = io.tol2 * g0;
            if (!(g0 > 0.0)) {  // b == 0 or NaN (non-positive diagonal)
                s_status[tid] = (g0 == 0.0) ? TFIN_STATUS_CONVERGED : TFIN_STATUS_BREAKDOWN;
                s_active[tid] = 0;
            }
        }
        __syncthreads();

        for (int it = 1; it <= io.maxit; ++it) {
            // ---------------- P1: q = A p, delta = p.q
            double2 d_acc = zero2;
            for (int g = warp; g < n_groups; g += nwarps) {
                const int i = g * RPW + rw;
                const bool valid = i < n;
                const double2 qv = apply_row(i, valid, p2);
                if (valid) {
                    const size_t o = (size_t)i * LPR + sp;
                    const double2 pv = p2[o];
                    q2[o] = qv;
                    d_acc.x = fma(pv.x, qv.x, d_acc.x);
                    d_acc.y = fma(pv.y, qv.y, d_acc.y);
                }
            }
            reduce2(d_acc, zero2);
            if (tid < S) {
                const double dl = s_tot[0][tid];
                double a = 0.0;
                if (s_active[tid]) {
                    if (dl > 0.0) a = s_gamma[tid] / dl;
                    else {
                        s_status[tid] = TFIN_STATUS_BREAKDOWN;
                        s_active[tid] = 0;
                        s_iters[tid] = it;
                    }
                }
                s_alpha[tid] = a;
            }
            __syncthreads();
            // ---------------- P2: x += a p, r -= a q, gamma' = r.D^-1 r
            const double2 al = make_double2(s_alpha[2 * sp], s_alpha[2 * sp + 1]);
            g_acc = zero2;
            for (int g = warp; g < n_groups; g += nwarps) {
                const int i = g * RPW + rw;
                if (i < n) {
                    const size_t o = (size_t)i * LPR + sp;
                    double2 xv = x2[o], rv = r2[o];
                    const double2 pv = p2[o], qv = q2[o];
                    const double2 di = dinv_of(i);
                    xv.x = fma(al.x, pv.x, xv.x);
                    xv.y = fma(al.y, pv.y, xv.y);
                    rv.x = fma(-al.x, qv.x, rv.x);
                    rv.y = fma(-al.y, qv.y, rv.y);
                    x2[o] = xv;
                    r2[o] = rv;
                    g_acc.x = fma(rv.x * di.x, rv.x, g_acc.x);
                    g_acc.y = fma(rv.y * di.y, rv.y, g_acc.y);
                }
            }
            reduce2(g_acc, zero2);
            if (tid < S) {
                const double gn = s_tot[0][tid];
                double b = 0.0;
                if (s_active[tid]) {
                    if (gn <= s_thresh[tid]) {
                        s_status[tid] = TFIN_STATUS_CONVERGED;
                        s_active[tid] = 0;
                        s_iters[tid] = it;
                    } else if (!(gn == gn)) {
                        s_status[tid] = TFIN_STATUS_BREAKDOWN;
                        s_active[tid] = 0;
                        s_iters[tid] = it;
                    } else {
                        b = gn / s_gamma[tid];
                        s_gamma[tid] = gn;
                    }
                }
                s_beta[tid] = b;
            }
            __syncthreads();
            if (tid == 0) {
                int na = 0;
                for (int s = 0; s < S; ++s) na += s_active[s];
                s_nactive = na;
            }
            __syncthreads();
            if (s_nactive == 0) break;
            // ---------------- P3: p = D^-1 r + b p     (frozen samples keep b = 0: harmless)
            const double2 be = make_double2(s_beta[2 * sp], s_beta[2 * sp + 1]);
            for (int g = warp; g < n_groups; g += nwarps) {
                const int i = g * RPW + rw;
                if (i < n) {
                    const size_t o = (size_t)i * LPR + sp;
                    const double2 rv = r2[o], pv = p2[o];
                    const double2 di = dinv_of(i);
                    p2[o] = make_double2(fma(be.x, pv.x, di.x * rv.x), fma(be.y, pv.y, di.y * rv.y));
                }
            }
            __syncthreads();
        }
        if (tid < S && s_active[tid]) s_iters[tid] = io.maxit;  // hit the cap
        __syncthreads();

        // ---------------- epilogue: true residual ||b - A x|| / ||b||, observables, optional w
        if (io.relres_out || io.status_out) {
            double2 rr = zero2, bb = zero2;
            for (int g = warp; g < n_groups; g += nwarps) {
                const int i = g * RPW + rw;
                const bool valid = i < n;
                const double2 ax = apply_row(i, valid, x2);
                if (valid) {
                    const double b = op.rhs[i];
                    rr.x = fma(b - ax.x, b - ax.x, rr.x);
                    rr.y = fma(b - ax.y, b - ax.y, rr.y);
                    bb.x = fma(b, b, bb.x);
                    bb.y = bb.x;
                }
            }
            reduce2(rr, bb);
            if (tid < S && s_base + tid < io.N) {
                const double bbt = s_tot[1][tid];
                const double relres = bbt > 0.0 ? sqrt(s_tot[0][tid] / bbt) : sqrt(s_tot[0][tid]);
                if (!(relres == relres)) s_status[tid] = TFIN_STATUS_BREAKDOWN;
                if (io.relres_out) io.relres_out[s_base + tid] = relres;
            }
            __syncthreads();
        }
        if (tid < S && s_base + tid < io.N) {
            if (io.iters_out) io.iters_out[s_base + tid] = s_iters[tid];
            if (io.status_out) io.status_out[s_base + tid] = s_status[tid];
        }
        if (io.qoi_out) {
            for (int o = 0; o < obs.rows; ++o) {
                double2 acc = zero2;
                const int j0 = obs.ptr[o], j1 = obs.ptr[o + 1];
                for (int j = j0 + warp * RPW + rw; j < j1; j += nwarps * RPW) {
                    const double w = obs.val[j];
                    const double2 xv = x2[(size_t)obs.idx[j] * LPR + sp];
                    acc.x = fma(w, xv.x, acc.x);
                    acc.y = fma(w, xv.y, acc.y);
                }
                reduce2(acc, zero2);
                if (tid < S && s_base + tid < io.N) io.qoi_out[(s_base + tid) * obs.rows + o] = s_tot[0][tid];
                __syncthreads();
            }
        }
        if (io.w_out) {
            const double* xs = reinterpret_cast<const double*>(x2);
            for (int s = 0; s < S && s_base + s < io.N; ++s)
                for (int i = tid; i < n; i += blockDim.x) io.w_out[(s_base + s) * (long long)n + i] = xs[(size_t)i * S + s];
        }
    }
}

}  // namespace tfin
