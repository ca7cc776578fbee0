// K1 / K2: batched Jacobi-PCG for meshes whose CG state fits one SM (n <= 8191 dofs).
//
// One persistent CTA per resident slot; each CTA pulls samples from a global counter and solves
//     A(sample) w = b,   qoi = B_obs w
// entirely on chip:
//   * the per-sample operator is formed in-kernel, either from the affine terms
//       vals = sum_t theta_t V_t                       (K1, AffineROMFin._F, averaged_affine_ROM.py:156-162)
//     or by element assembly from the nodal field, staged in shared memory by a TMA bulk copy,
//       vals = sum_e mean(k|e) K_e + Bi M              (K2, Fin._F, forward_solve.py:160-161)
//   * Jacobi preconditioning is applied as a symmetric diagonal scaling  A~ = D^-1/2 A D^-1/2  done once per
//     sample, so the iteration is plain CG on a unit-diagonal matrix (identical iterates to Jacobi-PCG in
//     exact arithmetic, no 1/diag multiply and no z vector in the loop)
//   * thread t owns rows t, t+T, ..., t+(R-1)T: x, r, p, q live in registers; the packed ELL column offsets live
//     in registers; the first WR off-diagonal values per row live in registers, the rest in shared memory; only
//     the residual r is published to shared memory for the SpMV gather
//   * Chronopoulos-Gear CG: ONE fused block reduction (r.r, r.Ar) and two barriers per iteration
//   * epilogue: true residual, B_obs projection (warp per observation row), optional w write-back
//   * ADJ variants (nodal operator only) keep the assembled operator on chip and run the ADJOINT solves of
//     Fin.gradient / Fin.sensitivity (forward_solve.py:293-342) right after the forward solve:
//       A v = -B_obs^T (B_obs w - data)   (gradient)      or      A v_o = -B_obs[o,:]^T, o = 1..n_obs   (sensitivity)
//     followed by the gradient form  g_i = sum_{e ni i} (1/3) w_e^T K_e v_e  (= assemble(k_hat grad w . grad v dx))
//
// Results do not depend on which CTA solves a sample (fixed reduction order) => bit-reproducible indexing.
#pragma once

#include "common.cuh"

namespace tfin {

struct PcgOp {
    int n, ld, W, n_terms, n_cells;
    const uint16_t* col;  // [W][ld]   off-diagonal columns (padding: col = row, val = 0)
    const double* rhs;    // [ld]
    // affine (K1)
    const double* val;    // [n_terms][W][ld]
    const double* diag;   // [n_terms][ld]
    // nodal (K2)
    const int* cell;      // [2][W][ld]   the (<=2) cells sharing edge (row, col); n_cells = none
    const double* coef;   // [2][W][ld]   K_e[a][b] of that cell for this entry
    const double* cst;    // [W][ld]      constant (Robin) part
    const int* dptr;      // [ld+1]       diagonal: variable-length cell list
    const int* dcell;     // [dnnz]
    const double* dcoef;  // [dnnz]
    const double* dcst;   // [ld]
    const int* cells;     // [n_cells][3]
    int coef_mode;        // cell coefficient of the nodal operator: 0 = mean(k) (exact for P1 k), 1 = quadrature of exp(k)
};

// Cell coefficient (1/|e|) int_e c(k) dx of the nodal-conductivity stiffness for vertex values (ka, kb, kc):
//   mode 0: c = k      -> mean of the vertex values (Fin._F, fom/forward_solve.py:160-161)
//   mode 1: c = exp(k) -> the degree-3 rule dolfin's form compiler picks for exp(P1) * grad(P1).grad(P1)
//           (fom/forward_solve_exp.py:160-161; UFL estimates degree(exp(k)) = 1 + 2, FIAT's degree-3 triangle scheme is
//           the 6-point Strang-Fix rule with equal weights: barycentric permutations of (a, b, c) below)
__device__ __forceinline__ double cell_coefficient(int mode, double ka, double kb, double kc) {
    if (mode == 0) return ((ka + kb) + kc) / 3.0;
    const double a = 0.659027622374092, b = 0.231933368553031, c = 0.109039009072877;
    double s = exp(fma(a, ka, fma(b, kb, c * kc)));
    s += exp(fma(a, ka, fma(c, kb, b * kc)));
    s += exp(fma(b, ka, fma(a, kb, c * kc)));
    s += exp(fma(b, ka, fma(c, kb, a * kc)));
    s += exp(fma(c, ka, fma(a, kb, b * kc)));
    s += exp(fma(c, ka, fma(b, kb, a * kc)));
    return s * (1.0 / 6.0);
}

// Weight (1/|e|) int_e lambda_a exp(k) dx of vertex a (values ka; the other two vertices kb, kc) in the gradient form
// assemble(k_hat * exp(k) * inner(grad z, grad v) * dx) of fom/forward_solve_exp.py:299, 328: the integrand has estimated
// degree 1 + 3, for which FIAT's triangle scheme is the 6-point degree-4 Strang-Fix rule (two orbits of three points).
// For the plain parametrisation the weight is 1/3.
__device__ __forceinline__ double vertex_weight_exp(double ka, double kb, double kc) {
    const double a1 = 0.816847572980459, b1 = 0.091576213509771, w1 = 0.109951743655322;
    const double a2 = 0.108103018168070, b2 = 0.445948490915965, w2 = 0.223381589678011;
    double s = w1 * a1 * exp(fma(a1, ka, b1 * (kb + kc)));
    s = fma(w1 * b1, exp(fma(a1, kb, b1 * (ka + kc))) + exp(fma(a1, kc, b1 * (ka + kb))), s);
    s = fma(w2 * a2, exp(fma(a2, ka, b2 * (kb + kc))), s);
    s = fma(w2 * b2, exp(fma(a2, kb, b2 * (ka + kc))) + exp(fma(a2, kc, b2 * (ka + kb))), s);
    return s;
}

struct CsrRows {
    int rows;
    const int* ptr;
    const int* idx;
    const double* val;
};

// adjoint outputs of the ADJ kernel variants (mode 0 = none)
struct PcgAdj {
    int mode;                // 1: gradient of 0.5 ||B_obs w - data||^2;  2: sensitivity d(B_obs w)/dk (n_obs solves)
    const double* data;      // observations, row s at data + s * data_stride (data_stride = 0: one shared vector)
    long long data_stride;
    CsrRows obsT;            // B_obs^T: n rows over n_obs columns
    const double* Ke;        // [n_cells][3][3]
    double* grad_out;        // mode 1: (N, n);  mode 2: (N, n_obs, n)
    double* cost_out;        // mode 1: (N) | NULL, 0.5 ||B_obs w - data||^2
};

struct PcgIO {
    const double* in;  // (N, in_stride)
    long long N;
    int in_stride;
    double tol2;  // tol^2 on r.z / r0.z0
    int maxit;
    double* w_out;
    double* qoi_out;
    int* iters_out;
    int* status_out;
    double* relres_out;
    unsigned long long* counter;
};

// Shared-memory carve-up shared by host (size) and device (pointers).
struct PcgSmem {
    size_t r_off, part_off, misc_off, adj_off, dsi_off, kbar_off, kbuf_off, val_off, total;
    __host__ __device__ static PcgSmem make(int w_smem, int np, int n_cells /*0 = affine*/, int n) {
        PcgSmem s;
        size_t o = 0;
        s.r_off = o;    o += (size_t)np * sizeof(double);
        s.part_off = o; o += 64 * sizeof(double);
        s.misc_off = o; o += 32 * sizeof(double);      // theta[16] | next sample | mbarrier
        s.adj_off = o;  o += 64 * sizeof(double);      // adjoint source coefficients (n_obs <= 64)
        s.dsi_off = o;  o += (size_t)np * sizeof(double);
        s.kbar_off = o; o += n_cells ? (((size_t)n_cells + 2 + 1) & ~size_t(1)) * sizeof(double) : 0;
        s.kbuf_off = o; o += n_cells ? (((size_t)n + 4 + 1) & ~size_t(1)) * sizeof(double) : 0;
        s.val_off = o;  o += (size_t)w_smem * np * sizeof(double);
        s.total = (o + 15) & ~size_t(15);
        return s;
    }
};

// Block-wide sums of two values with one barrier.  Stage 1 folds both values through ONE 5-step butterfly
// (lower half-warp accumulates a, upper half b); stage 2 re-reduces the per-warp partials lane-parallel.
// Every thread returns bit-identical totals.  s_part: 64 doubles.
__device__ __forceinline__ void block_sum2(double& a, double& b, double* s_part, int lane, int warp,
                                           int nwarps) {
    const bool hi = lane >= 16;
    const double send = hi ? a : b;
    double keep = hi ? b : a;
    keep += __shfl_xor_sync(0xffffffffu, send, 16);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) keep += __shfl_xor_sync(0xffffffffu, keep, o);
    double* base = s_part + (hi ? 32 : 0);
    if ((lane & 15) == 0) base[warp] = keep;
    __syncthreads();
    const int l = lane & 15;
    double v = (l < nwarps ? base[l] : 0.0) + (l + 16 < nwarps ? base[l + 16] : 0.0);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    a = __shfl_sync(0xffffffffu, v, 0);
    b = __shfl_sync(0xffffffffu, v, 16);
}

// R rows per thread, WT padded ELL width (even), WR of the WT off-diagonal values per row in registers.
template <int R, int WT, int WR, bool NODAL, int MAXT, int MINB, bool ADJ = false>
__global__ void __launch_bounds__(MAXT, MINB) pcg_kernel(PcgOp op, CsrRows obs, PcgIO io, PcgAdj adj) {
    static_assert(!ADJ || NODAL, "adjoint variants exist for the nodal operator only");
    static_assert(WT % 2 == 0 && WR <= WT, "bad ELL template widths");
    constexpr int WS = WT - WR;  // slots whose values live in shared memory
    extern __shared__ __align__(16) unsigned char smem[];
    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const int np = R * T, W = op.W, n = op.n, ld = op.ld, nc = NODAL ? op.n_cells : 0;
    const PcgSmem L = PcgSmem::make(WS, np, nc, n);
    double* s_r = reinterpret_cast<double*>(smem + L.r_off);
    double* s_part = reinterpret_cast<double*>(smem + L.part_off);
    double* s_theta = reinterpret_cast<double*>(smem + L.misc_off);
    long long* s_next = reinterpret_cast<long long*>(smem + L.misc_off) + 16;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + L.misc_off) + 20;
    double* s_adj = reinterpret_cast<double*>(smem + L.adj_off);
    double* s_dsi = reinterpret_cast<double*>(smem + L.dsi_off);
    double* s_kbar = reinterpret_cast<double*>(smem + L.kbar_off);
    double* s_kbuf = reinterpret_cast<double*>(smem + L.kbuf_off);
    double* s_val = reinterpret_cast<double*>(smem + L.val_off);
    const unsigned char* s_rb = reinterpret_cast<const unsigned char*>(s_r);

    // ---- once per CTA: packed byte offsets (8 * column) of this thread's ELL slots, two per register
    uint32_t pk[R][WT / 2];
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const int i = tid + k * T;
#pragma unroll
        for (int w = 0; w < WT; w += 2) {
            const uint32_t c0 = (i < n && w < W) ? op.col[(size_t)w * ld + i] : (uint32_t)i;
            const uint32_t c1 = (i < n && w + 1 < W) ? op.col[(size_t)(w + 1) * ld + i] : (uint32_t)i;
            pk[k][w / 2] = (c0 << 3) | (c1 << 19);
        }
    }
    auto gather = [&](int k, int w) -> double {
        const uint32_t off = (w & 1) ? (pk[k][w >> 1] >> 16) : (pk[k][w >> 1] & 0xffffu);
        return *reinterpret_cast<const double*>(s_rb + off);
    };
    if (NODAL && tid == 0) {
        mbar_init(s_bar, 1);
        mbar_fence_init();
        s_kbar[nc] = 0.0;  // sentinel "no cell"
    }
    uint32_t parity = 0;

    for (;;) {
        __syncthreads();
        if (tid == 0) *s_next = (long long)atomicAdd(io.counter, 1ULL);
        __syncthreads();
        const long long sample = *s_next;
        if (sample >= io.N) break;

        // ================= per-sample operator =================
        double dsi[R];      // 1/sqrt(diag)
        double av[R][WR > 0 ? WR : 1];
        if (!NODAL) {
            if (tid < op.n_terms) s_theta[tid] = tid == 0 ? 1.0 : io.in[sample * io.in_stride + tid - 1];
            __syncthreads();
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const int i = tid + k * T;
                double d = 1.0;
                if (i < n) {
                    d = 0.0;
                    for (int t = 0; t < op.n_terms; ++t) d = fma(s_theta[t], op.diag[t * ld + i], d);
                }
                dsi[k] = 1.0 / sqrt(d);
                s_r[i] = dsi[k];
            }
        } else {
            // stage the nodal field k (n doubles) into shared memory with ONE TMA bulk copy.  cp.async.bulk needs
            // 16-byte aligned addresses/sizes; rows of odd n start 8-byte aligned, so the row is placed in s_kbuf
            // with the same 16-byte phase and the <=1 head / tail doubles are loaded with ordinary loads.
            const double* krow = io.in + sample * (long long)io.in_stride;
            const int off = (int)((reinterpret_cast<uintptr_t>(krow) >> 3) & 1);
            const int head = off;              // doubles before the first 16B boundary
            const int body = (n - head) & ~1;  // doubles moved by the bulk copy
            if (tid == 0) {
                if (body > 0) {
                    mbar_expect_tx(s_bar, (uint32_t)body * 8u);
                    tma_bulk_g2s(s_kbuf + off + head, krow + head, (uint32_t)body * 8u, s_bar);
                }
                if (head) s_kbuf[off] = krow[0];
                if (head + body < n) s_kbuf[off + n - 1] = krow[n - 1];
            }
            if (body > 0) mbar_wait(s_bar, parity);
            parity ^= (body > 0);
            __syncthreads();
            const double* kk = s_kbuf + off;
            // cell means:  int k grad w.grad v over a cell = mean(k at its vertices) * K_e
            for (int e = tid; e < nc; e += T) {
                const int a = op.cells[3 * e], b = op.cells[3 * e + 1], c = op.cells[3 * e + 2];
                s_kbar[e] = cell_coefficient(op.coef_mode, kk[a], kk[b], kk[c]);
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const int i = tid + k * T;
                double d = 1.0;
                if (i < n) {
                    d = op.dcst[i];
                    for (int j = op.dptr[i]; j < op.dptr[i + 1]; ++j)
                        d = fma(op.dcoef[j], s_kbar[op.dcell[j]], d);
                }
                dsi[k] = 1.0 / sqrt(d);
                s_r[i] = dsi[k];
            }
        }
        __syncthreads();  // 1/sqrt(diag) of every row visible
#pragma unroll
        for (int w = 0; w < WT; ++w) {
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const int i = tid + k * T;
                double v = 0.0;
                if (i < n && w < W) {
                    const size_t o = (size_t)w * ld + i;
                    if (!NODAL) {
                        for (int t = 0; t < op.n_terms; ++t)
                            v = fma(s_theta[t], op.val[(size_t)t * W * ld + o], v);
                    } else {
                        const size_t plane = (size_t)W * ld;
                        v = op.cst[o];
                        v = fma(op.coef[o], s_kbar[op.cell[o]], v);
                        v = fma(op.coef[plane + o], s_kbar[op.cell[plane + o]], v);
                    }
                    v *= dsi[k] * gather(k, w);
                }
                if (w < WR) av[k][w < WR ? w : 0] = v;
                else s_val[(w - WR) * np + i] = v;
            }
        }
#pragma unroll
        for (int k = 0; k < R; ++k) s_dsi[tid + k * T] = dsi[k];
        __syncthreads();  // all gathers of dsi done before s_r is reused for r

        // ================= CG on the scaled system =================
        auto spmv = [&](const double (&rv)[R], double (&sv)[R]) {
#pragma unroll
            for (int k = 0; k < R; ++k) sv[k] = rv[k];  // unit diagonal
#pragma unroll
            for (int w = 0; w < WT; ++w) {
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    const double a = (w < WR) ? av[k][w < WR ? w : 0] : s_val[(w - WR) * np + tid + k * T];
                    sv[k] = fma(a, gather(k, w), sv[k]);
                }
            }
        };
        // CG on the scaled system for the scaled right-hand side rin (registers); returns the scaled solution in x
        auto cg_solve = [&](const double (&rin)[R], double (&x)[R], int& it, int& status) {
            double r[R], p[R], q[R], s[R];
#pragma unroll
            for (int k = 0; k < R; ++k) {
                x[k] = 0.0;
                r[k] = rin[k];
                s_r[tid + k * T] = r[k];
            }
            __syncthreads();
            spmv(r, s);
            double gam = 0.0, del = 0.0;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                gam = fma(r[k], r[k], gam);
                del = fma(r[k], s[k], del);
            }
            block_sum2(gam, del, s_part, lane, warp, nwarps);
            status = TFIN_STATUS_MAXIT;
            it = 0;
            if (!(gam > 0.0) || !(del > 0.0)) {  // b == 0 (x = 0 is exact) or not SPD / NaN
                status = (gam == 0.0) ? TFIN_STATUS_CONVERGED : TFIN_STATUS_BREAKDOWN;
                return;
            }
            const double thresh = io.tol2 * gam;
            double alpha = gam / del, denom = del;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                p[k] = r[k];
                q[k] = s[k];
            }
            while (it < io.maxit) {
                ++it;
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    x[k] = fma(alpha, p[k], x[k]);
                    r[k] = fma(-alpha, q[k], r[k]);
                    s_r[tid + k * T] = r[k];
                }
                __syncthreads();  // r visible; everyone is done with s_part of the previous iteration
                spmv(r, s);
                double gn = 0.0, dl = 0.0;
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    gn = fma(r[k], r[k], gn);
                    dl = fma(r[k], s[k], dl);
                }
                block_sum2(gn, dl, s_part, lane, warp, nwarps);  // barrier inside: all gathers of r done
                if (gn <= thresh) {
                    status = TFIN_STATUS_CONVERGED;
                    break;
                }
                const double beta = gn / gam;
                denom = dl - beta * beta * denom;  // = dl - beta * gn / alpha_old
                if (!(denom > 0.0) || !(gn == gn)) {
                    status = TFIN_STATUS_BREAKDOWN;
                    break;
                }
                alpha = gn / denom;
                gam = gn;
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    p[k] = fma(beta, p[k], r[k]);
                    q[k] = fma(beta, q[k], s[k]);
                }
            }
        };
        double x[R], s[R];
        int status, it;
        {
            double b0[R];
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const int i = tid + k * T;
                b0[k] = (i < n ? op.rhs[i] : 0.0) * dsi[k];
            }
            cg_solve(b0, x, it, status);
        }

        // ================= epilogue =================
        __syncthreads();
        if (io.relres_out || io.status_out) {  // true residual of the scaled system: ||b~ - A~ x~|| / ||b~||
#pragma unroll
            for (int k = 0; k < R; ++k) s_r[tid + k * T] = x[k];
            __syncthreads();
            spmv(x, s);
            double rr = 0.0, bb = 0.0;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const int i = tid + k * T;
                const double bt = (i < n ? op.rhs[i] : 0.0) * s_dsi[i];
                const double t = bt - s[k];
                rr = fma(t, t, rr);
                bb = fma(bt, bt, bb);
            }
            block_sum2(rr, bb, s_part, lane, warp, nwarps);
            const double relres = bb > 0.0 ? sqrt(rr / bb) : sqrt(rr);
            if (!(relres == relres)) status = TFIN_STATUS_BREAKDOWN;
            if (tid == 0 && io.relres_out) io.relres_out[sample] = relres;
            __syncthreads();
        }
        if (!ADJ && tid == 0) {
            if (io.iters_out) io.iters_out[sample] = it;
            if (io.status_out) io.status_out[sample] = status;
        }
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int i = tid + k * T;
            x[k] *= s_dsi[i];  // w = D^-1/2 x~
            s_r[i] = x[k];
        }
        __syncthreads();
        if (io.qoi_out) {
            for (int o = warp; o < obs.rows; o += nwarps) {
                double acc = 0.0;
                for (int j = obs.ptr[o] + lane; j < obs.ptr[o + 1]; j += 32)
                    acc = fma(obs.val[j], s_r[obs.idx[j]], acc);
                acc = warp_sum(acc);
                if (lane == 0) io.qoi_out[sample * obs.rows + o] = acc;
            }
        }
        if (io.w_out) {
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const int i = tid + k * T;
                if (i < n) io.w_out[sample * (long long)n + i] = x[k];
            }
        }
        if (ADJ) {
            // ================= adjoint solves + gradient form (operator still on chip) =================
            double* s_w = s_kbuf;  // the nodal field is no longer needed: keep the state w here
            const int n_adj = adj.mode == 2 ? obs.rows : 1;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const int i = tid + k * T;
                if (i < n) s_w[i] = x[k];
            }
            if (adj.mode == 1) {  // source coefficients c_o = (B_obs w)_o - data_o
                const double* d = adj.data + sample * adj.data_stride;
                for (int o = warp; o < obs.rows; o += nwarps) {
                    double acc = 0.0;
                    for (int j = obs.ptr[o] + lane; j < obs.ptr[o + 1]; j += 32)
                        acc = fma(obs.val[j], s_r[obs.idx[j]], acc);
                    acc = warp_sum(acc);
                    if (lane == 0) s_adj[o] = acc - d[o];
                }
            }
            __syncthreads();
            if (adj.mode == 1 && adj.cost_out && tid == 0) {
                double c = 0.0;
                for (int o = 0; o < obs.rows; ++o) c = fma(s_adj[o], s_adj[o], c);
                adj.cost_out[sample] = 0.5 * c;
            }
            for (int a = 0; a < n_adj; ++a) {
                double rhs[R], v[R];
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    const int i = tid + k * T;
                    double rr = 0.0;
                    if (i < n)
                        for (int j = adj.obsT.ptr[i]; j < adj.obsT.ptr[i + 1]; ++j) {
                            const int o = adj.obsT.idx[j];
                            const double c = adj.mode == 1 ? s_adj[o] : (o == a ? 1.0 : 0.0);
                            rr = fma(-adj.obsT.val[j], c, rr);
                        }
                    rhs[k] = rr * s_dsi[i];
                }
                int it_a, st_a;
                cg_solve(rhs, v, it_a, st_a);
                if (st_a > status) status = st_a;
                __syncthreads();
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    const int i = tid + k * T;
                    s_r[i] = v[k] * s_dsi[i];  // adjoint state
                }
                __syncthreads();
                const bool expm = op.coef_mode == 1;
                for (int e = tid; e < nc; e += T) {  // (1/3) w_e^T K_e v_e per cell (exp(k): vertex weights below)
                    const int ca = op.cells[3 * e], cb = op.cells[3 * e + 1], cc = op.cells[3 * e + 2];
                    const double* K = adj.Ke + 9 * (size_t)e;
                    const double va = s_r[ca], vb = s_r[cb], vc = s_r[cc];
                    const double t0 = fma(K[0], va, fma(K[1], vb, K[2] * vc));
                    const double t1 = fma(K[3], va, fma(K[4], vb, K[5] * vc));
                    const double t2 = fma(K[6], va, fma(K[7], vb, K[8] * vc));
                    s_kbar[e] = fma(s_w[ca], t0, fma(s_w[cb], t1, s_w[cc] * t2)) * (expm ? 1.0 : 1.0 / 3.0);
                }
                __syncthreads();
                double* g = adj.grad_out + (sample * n_adj + a) * (long long)n;
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    const int i = tid + k * T;
                    if (i < n) {
                        double acc = 0.0;
                        if (!expm) {
                            for (int j = op.dptr[i]; j < op.dptr[i + 1]; ++j) acc += s_kbar[op.dcell[j]];
                        } else {  // k_hat-weighted quadrature of exp(k): the field is re-read from global memory
                            const double* krow = io.in + sample * (long long)io.in_stride;
                            for (int j = op.dptr[i]; j < op.dptr[i + 1]; ++j) {
                                const int e = op.dcell[j];
                                const int c0 = op.cells[3 * e], c1 = op.cells[3 * e + 1], c2 = op.cells[3 * e + 2];
                                const int ob = c0 == i ? c1 : c0, oc = c2 == i ? c1 : c2;  // the two other vertices
                                acc = fma(s_kbar[e], vertex_weight_exp(krow[i], krow[ob], krow[oc]), acc);
                            }
                        }
                        g[i] = acc;
                    }
                }
                __syncthreads();
            }
            if (tid == 0) {
                if (io.iters_out) io.iters_out[sample] = it;
                if (io.status_out) io.status_out[sample] = status;
            }
        }
    }
}

}  // namespace tfin
