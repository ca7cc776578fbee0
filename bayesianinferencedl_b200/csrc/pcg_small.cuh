// K1 / K2: batched Jacobi-PCG for meshes whose CG state fits one SM (n <~ 4k dofs).
//
// One persistent CTA per resident slot; each CTA pulls samples from a global counter and solves
//     A(sample) w = b,   qoi = B_obs w
// entirely on chip:
//   * the shared ELL sparsity pattern (uint16 columns) is loaded to shared memory ONCE per CTA
//   * per sample the ELL values are formed in shared memory, either from the affine terms
//       vals = sum_t theta_t V_t                       (K1, AffineROMFin._F, averaged_affine_ROM.py:156-162)
//     or by in-kernel element assembly from the nodal field staged with a TMA bulk copy
//       vals = sum_e mean(k|e) K_e + Bi M              (K2, Fin._F, forward_solve.py:160-161)
//   * thread t owns rows t, t+T, ..., t+(R-1)T: x, r, p, q, 1/diag live in registers, only the
//     preconditioned residual z is published to shared memory for the SpMV gather
//   * Chronopoulos-Gear PCG: ONE fused block reduction (r.z, z.Az) and two barriers per iteration
//   * epilogue: true residual, B_obs projection (warp per observation row), optional w write-back
//
// Results are independent of which CTA solves a sample (fixed reduction order), so sample indexing is
// bit-reproducible for a given launch geometry.
#pragma once

#include "common.cuh"

namespace tfin {

struct EllAffine {
    int n, ld, W, n_terms;
    const uint16_t* col;  // [W][ld]   off-diagonal columns (padding: col = row, val = 0)
    const double* val;    // [n_terms][W][ld]
    const double* diag;   // [n_terms][ld]
    const double* rhs;    // [ld]
};

struct EllNodal {
    int n, ld, W, n_cells;
    const uint16_t* col;  // [W][ld]
    const int* cell;      // [2][W][ld]   the (<=2) cells sharing edge (row, col); n_cells = none
    const double* coef;   // [2][W][ld]   K_e[a][b] of that cell for this entry
    const double* cst;    // [W][ld]      constant (Robin) part
    const int* dptr;      // [ld+1]       diagonal: variable-length cell list
    const int* dcell;     // [dnnz]
    const double* dcoef;  // [dnnz]
    const double* dcst;   // [ld]
    const int* cells;     // [n_cells][3]
    const double* rhs;    // [ld]
};

struct CsrRows {
    int rows;
    const int* ptr;
    const int* idx;
    const double* val;
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
    size_t val_off, z_off, part_off, misc_off, col_off, kbar_off, kbuf_off, total;
    __host__ __device__ static PcgSmem make(int W, int np, int n_cells /*0 = affine*/, int n) {
        PcgSmem s;
        size_t o = 0;
        s.val_off = o;  o += (size_t)W * np * sizeof(double);
        s.z_off = o;    o += (size_t)np * sizeof(double);
        s.part_off = o; o += 4 * 32 * sizeof(double);  // two regions of 32 double2 partials
        s.misc_off = o; o += 32 * sizeof(double);      // theta[16] | next sample | mbarrier
        s.kbar_off = o; o += n_cells ? (((size_t)n_cells + 2 + 1) & ~size_t(1)) * sizeof(double) : 0;
        s.kbuf_off = o; o += n_cells ? (((size_t)n + 4 + 1) & ~size_t(1)) * sizeof(double) : 0;
        s.col_off = o;  o += (size_t)W * np * sizeof(uint16_t);
        s.total = (o + 15) & ~size_t(15);
        return s;
    }
};

__device__ __forceinline__ void block_sum2(double& a, double& b, double* s_part, int lane, int warp,
                                           int nwarps) {
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) reinterpret_cast<double2*>(s_part)[warp] = make_double2(a, b);
    __syncthreads();
    double sa = 0.0, sb = 0.0;
    for (int w = 0; w < nwarps; ++w) {
        const double2 v = reinterpret_cast<const double2*>(s_part)[w];
        sa += v.x;
        sb += v.y;
    }
    a = sa;
    b = sb;
}

// The CG core.  On entry: s_val (this thread's rows) and dg[] (diagonal) are set, bvec[] = rhs rows.
// On exit x[] holds the solution rows of this thread; returns iterations, sets status.
template <int R>
__device__ __forceinline__ int cg_core(const double* __restrict__ s_val, const uint16_t* __restrict__ s_col,
                                       double* __restrict__ s_z, double* __restrict__ s_part, int W,
                                       int np, int T, int tid, const double (&dg)[R],
                                       const double (&bvec)[R], double tol2, int maxit, double (&x)[R],
                                       int& status) {
    const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    double dinv[R], r[R], p[R], q[R], z[R], s[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
        dinv[k] = 1.0 / dg[k];
        x[k] = 0.0;
        r[k] = bvec[k];
        z[k] = dinv[k] * r[k];
        s_z[tid + k * T] = z[k];
    }
    __syncthreads();

    auto spmv = [&](double& lg, double& ld) {
#pragma unroll
        for (int k = 0; k < R; ++k) s[k] = dg[k] * z[k];
        for (int w = 0; w < W; ++w) {
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const int o = w * np + tid + k * T;
                s[k] = fma(s_val[o], s_z[s_col[o]], s[k]);
            }
        }
        lg = 0.0;
        ld = 0.0;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            lg = fma(r[k], z[k], lg);
            ld = fma(z[k], s[k], ld);
        }
    };

    double gam, del;
    spmv(gam, del);
    block_sum2(gam, del, s_part, lane, warp, nwarps);
    status = TFIN_STATUS_MAXIT;
    if (!(gam > 0.0) || !(del > 0.0)) {  // b == 0 (x = 0 is exact) or not SPD
        status = (gam == 0.0) ? TFIN_STATUS_CONVERGED : TFIN_STATUS_BREAKDOWN;
        return 0;
    }
    const double thresh = tol2 * gam;
    double alpha = gam / del, denom = del;
#pragma unroll
    for (int k = 0; k < R; ++k) {
        p[k] = z[k];
        q[k] = s[k];
    }
    int it = 0;
    while (it < maxit) {
        ++it;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            x[k] = fma(alpha, p[k], x[k]);
            r[k] = fma(-alpha, q[k], r[k]);
            z[k] = dinv[k] * r[k];
            s_z[tid + k * T] = z[k];
        }
        __syncthreads();  // z visible; everyone is done reading s_part of the previous iteration
        double gn, dl;
        spmv(gn, dl);
        block_sum2(gn, dl, s_part, lane, warp, nwarps);  // barrier inside: all gathers of z done
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
            p[k] = fma(beta, p[k], z[k]);
            q[k] = fma(beta, q[k], s[k]);
        }
    }
    return it;
}

// Epilogue shared by K1/K2: publish x, true residual, observation projection, outputs.
template <int R>
__device__ __forceinline__ void pcg_epilogue(const double* s_val, const uint16_t* s_col, double* s_z,
                                             double* s_part, int W, int np, int T, int tid, int n,
                                             const double (&dg)[R], const double (&bvec)[R],
                                             const double (&x)[R], const CsrRows& obs, const PcgIO& io,
                                             long long sample, int iters, int status) {
    const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    __syncthreads();  // all SpMV gathers of the last iteration are complete
#pragma unroll
    for (int k = 0; k < R; ++k) s_z[tid + k * T] = x[k];
    __syncthreads();
    if (io.relres_out || io.status_out) {
        double rr = 0.0, bb = 0.0;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            double ax = dg[k] * x[k];
            for (int w = 0; w < W; ++w) {
                const int o = w * np + tid + k * T;
                ax = fma(s_val[o], s_z[s_col[o]], ax);
            }
            const double t = bvec[k] - ax;
            rr = fma(t, t, rr);
            bb = fma(bvec[k], bvec[k], bb);
        }
        block_sum2(rr, bb, s_part + 64, lane, warp, nwarps);
        const double relres = bb > 0.0 ? sqrt(rr / bb) : sqrt(rr);
        if (!(relres == relres)) status = TFIN_STATUS_BREAKDOWN;
        if (tid == 0 && io.relres_out) io.relres_out[sample] = relres;
    }
    if (tid == 0) {
        if (io.iters_out) io.iters_out[sample] = iters;
        if (io.status_out) io.status_out[sample] = status;
    }
    if (io.qoi_out) {
        for (int o = warp; o < obs.rows; o += nwarps) {
            double acc = 0.0;
            for (int j = obs.ptr[o] + lane; j < obs.ptr[o + 1]; j += 32)
                acc = fma(obs.val[j], s_z[obs.idx[j]], acc);
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
}

// ------------------------------------------------------------------------------------------- K1
template <int R, int MAXT>
__global__ void __launch_bounds__(MAXT) pcg_affine_kernel(EllAffine op, CsrRows obs, PcgIO io) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int T = blockDim.x, tid = threadIdx.x;
    const int np = R * T, W = op.W, n = op.n, ld = op.ld;
    const PcgSmem L = PcgSmem::make(W, np, 0, n);
    double* s_val = reinterpret_cast<double*>(smem + L.val_off);
    double* s_z = reinterpret_cast<double*>(smem + L.z_off);
    double* s_part = reinterpret_cast<double*>(smem + L.part_off);
    double* s_theta = reinterpret_cast<double*>(smem + L.misc_off);
    long long* s_next = reinterpret_cast<long long*>(smem + L.misc_off) + 16;
    uint16_t* s_col = reinterpret_cast<uint16_t*>(smem + L.col_off);

    double bvec[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const int i = tid + k * T;
        bvec[k] = i < n ? op.rhs[i] : 0.0;
        for (int w = 0; w < W; ++w) s_col[w * np + i] = i < n ? op.col[w * ld + i] : (uint16_t)i;
    }

    for (;;) {
        __syncthreads();
        if (tid == 0) *s_next = (long long)atomicAdd(io.counter, 1ULL);
        __syncthreads();
        const long long sample = *s_next;
        if (sample >= io.N) break;
        if (tid < op.n_terms)
            s_theta[tid] = tid == 0 ? 1.0 : io.in[sample * io.in_stride + tid - 1];
        __syncthreads();

        // ---- per-sample operator: vals = sum_t theta_t V_t (thread-private rows, no barrier needed)
        double dg[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int i = tid + k * T;
            double d = 0.0;
            if (i < n)
                for (int t = 0; t < op.n_terms; ++t) d = fma(s_theta[t], op.diag[t * ld + i], d);
            dg[k] = i < n ? d : 1.0;
        }
        for (int w = 0; w < W; ++w) {
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const int i = tid + k * T;
                double v = 0.0;
                if (i < n)
                    for (int t = 0; t < op.n_terms; ++t)
                        v = fma(s_theta[t], op.val[((size_t)t * W + w) * ld + i], v);
                s_val[w * np + i] = v;
            }
        }

        double x[R];
        int status;
        const int iters = cg_core<R>(s_val, s_col, s_z, s_part, W, np, T, tid, dg, bvec, io.tol2,
                                     io.maxit, x, status);
        pcg_epilogue<R>(s_val, s_col, s_z, s_part, W, np, T, tid, n, dg, bvec, x, obs, io, sample, iters,
                        status);
    }
}

// ------------------------------------------------------------------------------------------- K2
template <int R, int MAXT>
__global__ void __launch_bounds__(MAXT) pcg_nodal_kernel(EllNodal op, CsrRows obs, PcgIO io) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int T = blockDim.x, tid = threadIdx.x;
    const int np = R * T, W = op.W, n = op.n, ld = op.ld, nc = op.n_cells;
    const PcgSmem L = PcgSmem::make(W, np, nc, n);
    double* s_val = reinterpret_cast<double*>(smem + L.val_off);
    double* s_z = reinterpret_cast<double*>(smem + L.z_off);
    double* s_part = reinterpret_cast<double*>(smem + L.part_off);
    long long* s_next = reinterpret_cast<long long*>(smem + L.misc_off) + 16;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + L.misc_off) + 20;
    double* s_kbar = reinterpret_cast<double*>(smem + L.kbar_off);
    double* s_kbuf = reinterpret_cast<double*>(smem + L.kbuf_off);
    uint16_t* s_col = reinterpret_cast<uint16_t*>(smem + L.col_off);

    double bvec[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const int i = tid + k * T;
        bvec[k] = i < n ? op.rhs[i] : 0.0;
        for (int w = 0; w < W; ++w) s_col[w * np + i] = i < n ? op.col[w * ld + i] : (uint16_t)i;
    }
    if (tid == 0) {
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

        // ---- stage the nodal field k (n doubles) into shared memory with ONE TMA bulk copy.
        // cp.async.bulk needs 16-byte aligned addresses/sizes; rows of odd n start 8-byte aligned, so the
        // row is placed in s_kbuf with the same 16-byte phase and the <=1 head / tail doubles are
        // loaded with ordinary loads.
        const double* krow = io.in + sample * (long long)io.in_stride;
        const int off = (int)((reinterpret_cast<uintptr_t>(krow) >> 3) & 1);
        const int head = off;                      // doubles before the first 16B boundary
        const int body = (n - head) & ~1;          // doubles moved by the bulk copy
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

        // ---- cell means:  int k grad w.grad v over a cell = mean(k at its vertices) * K_e
        for (int e = tid; e < nc; e += T) {
            const int a = op.cells[3 * e], b = op.cells[3 * e + 1], c = op.cells[3 * e + 2];
            s_kbar[e] = ((kk[a] + kk[b]) + kk[c]) / 3.0;
        }
        __syncthreads();

        // ---- in-kernel assembly of this thread's rows
        double dg[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int i = tid + k * T;
            double d = 1.0;
            if (i < n) {
                d = op.dcst[i];
                for (int j = op.dptr[i]; j < op.dptr[i + 1]; ++j) d = fma(op.dcoef[j], s_kbar[op.dcell[j]], d);
            }
            dg[k] = d;
        }
        const size_t plane = (size_t)W * ld;
        for (int w = 0; w < W; ++w) {
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const int i = tid + k * T;
                double v = 0.0;
                if (i < n) {
                    const size_t o = (size_t)w * ld + i;
                    v = op.cst[o];
                    v = fma(op.coef[o], s_kbar[op.cell[o]], v);
                    v = fma(op.coef[plane + o], s_kbar[op.cell[plane + o]], v);
                }
                s_val[w * np + i] = v;
            }
        }

        double x[R];
        int status;
        const int iters = cg_core<R>(s_val, s_col, s_z, s_part, W, np, T, tid, dg, bvec, io.tol2,
                                     io.maxit, x, status);
        pcg_epilogue<R>(s_val, s_col, s_z, s_part, W, np, T, tid, n, dg, bvec, x, obs, io, sample, iters,
                        status);
    }
}

// ------------------------------------------------------------------------------------------- K0
// theta = Avg k for a batch of nodal fields: one warp per (sample, row); a streaming, HBM-bound kernel.
__global__ void __launch_bounds__(256) csr_project_kernel(CsrRows op, const double* __restrict__ in,
                                                          long long N, int n, double* __restrict__ out) {
    const long long gw = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long total = N * op.rows;
    const long long stride = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long t = gw; t < total; t += stride) {
        const long long s = t / op.rows;
        const int o = (int)(t - s * op.rows);
        const double* row = in + s * (long long)n;
        double acc = 0.0;
        for (int j = op.ptr[o] + lane; j < op.ptr[o + 1]; j += 32) acc = fma(op.val[j], row[op.idx[j]], acc);
        acc = warp_sum(acc);
        if (lane == 0) out[t] = acc;
    }
}

}  // namespace tfin
