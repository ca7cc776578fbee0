// K1f: OPTIONAL single-precision variant of the on-chip affine PCG (north_star: "1e-5 in the optional fp32 path").
//
// Same organisation as K1 (pcg_small.cuh): one sample per CTA, persistent CTAs fed by a global counter, symmetric
// Jacobi scaling done once per sample, Chronopoulos-Gear CG with one block reduction per iteration.  What changes:
//   * the scaled off-diagonal values, the published residual and the CG vectors r, p, q are fp32: the shared-memory
//     traffic that bounds K1 (78 % of the SMEM pipe) halves
//   * the operator is still FORMED in fp64 (theta combination, diagonal, scaling) and only then rounded, the solution
//     x is accumulated in fp64, and every dot product is reduced in fp64 -- so the error floor is the fp32 rounding of
//     the matrix entries and of the residual recurrence (about cond(A~) * 6e-8 ~ 1e-5 .. 1e-6 on this problem)
//   * the true residual returned in relres_out is recomputed with the fp32 operator but fp64 accumulation
// Observables (B_obs w) are computed in fp64 from the fp64 solution vector.
//
// MEASURED (tools/probe_fp32.py, 1e5 five-parameter samples, n = 1597): the error floor of the observables against the
// fp64 path is 3.6e-5 (max; reached from tol = 1e-8 on) -- the fp32 rounding of the matrix entries times cond(A~) --
// which MISSES the 1e-5 target of the north star; an iteration is 1.4x faster than fp64 (934 k solves/s at tol = 1e-8;
// the kernel turns issue-bound: conversions, shuffles and address arithmetic).  The fp64 kernel with tol = 1e-9 gives
// 2e-6 at 666 k solves/s and is the recommended reduced-accuracy setting; this kernel stays an opt-in experiment.
#pragma once

#include "pcg_small.cuh"

namespace tfin {

struct PcgF32Smem {
    size_t r_off, val_off, w_off, part_off, misc_off, total;
    __host__ __device__ static PcgF32Smem make(int WT, int np) {
        PcgF32Smem s;
        size_t o = 0;
        s.w_off = o;    o += (size_t)np * sizeof(double);        // 1/sqrt(diag), later the fp64 solution
        s.part_off = o; o += 64 * sizeof(double);
        s.misc_off = o; o += 32 * sizeof(double);                // theta[16] | next sample
        s.r_off = o;    o += (size_t)np * sizeof(float);
        s.val_off = o;  o += (size_t)WT * np * sizeof(float);
        s.total = (o + 15) & ~size_t(15);
        return s;
    }
};

template <int R, int WT, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) pcg_f32_kernel(PcgOp op, CsrRows obs, PcgIO io) {
    static_assert(WT % 2 == 0, "ELL width must be even");
    extern __shared__ __align__(16) unsigned char smem[];
    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const int np = R * T, W = op.W, n = op.n, ld = op.ld;
    const PcgF32Smem L = PcgF32Smem::make(WT, np);
    double* s_w = reinterpret_cast<double*>(smem + L.w_off);
    double* s_part = reinterpret_cast<double*>(smem + L.part_off);
    double* s_theta = reinterpret_cast<double*>(smem + L.misc_off);
    long long* s_next = reinterpret_cast<long long*>(smem + L.misc_off) + 16;
    float* s_r = reinterpret_cast<float*>(smem + L.r_off);
    float* s_val = reinterpret_cast<float*>(smem + L.val_off);
    const unsigned char* s_rb = reinterpret_cast<const unsigned char*>(s_r);

    // packed byte offsets (4 * column) of this thread's ELL slots, two per register
    uint32_t pk[R][WT / 2];
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const int i = tid + k * T;
#pragma unroll
        for (int w = 0; w < WT; w += 2) {
            const uint32_t c0 = (i < n && w < W) ? op.col[(size_t)w * ld + i] : (uint32_t)i;
            const uint32_t c1 = (i < n && w + 1 < W) ? op.col[(size_t)(w + 1) * ld + i] : (uint32_t)i;
            pk[k][w / 2] = (c0 << 2) | (c1 << 18);
        }
    }
    auto gather = [&](int k, int w) -> float {
        const uint32_t off = (w & 1) ? (pk[k][w >> 1] >> 16) : (pk[k][w >> 1] & 0xffffu);
        return *reinterpret_cast<const float*>(s_rb + off);
    };
    auto spmv = [&](const float (&rv)[R], float (&sv)[R]) {
#pragma unroll
        for (int k = 0; k < R; ++k) sv[k] = rv[k];  // unit diagonal
#pragma unroll
        for (int w = 0; w < WT; ++w)
#pragma unroll
            for (int k = 0; k < R; ++k) sv[k] = fmaf(s_val[w * np + tid + k * T], gather(k, w), sv[k]);
    };

    for (;;) {
        __syncthreads();
        if (tid == 0) *s_next = (long long)atomicAdd(io.counter, 1ULL);
        __syncthreads();
        const long long sample = *s_next;
        if (sample >= io.N) break;

        // ---- per-sample operator, formed in fp64 and rounded once
        if (tid < op.n_terms) s_theta[tid] = tid == 0 ? 1.0 : io.in[sample * io.in_stride + tid - 1];
        __syncthreads();
        double dsi[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int i = tid + k * T;
            double d = 1.0;
            if (i < n) {
                d = 0.0;
                for (int t = 0; t < op.n_terms; ++t) d = fma(s_theta[t], op.diag[t * ld + i], d);
            }
            dsi[k] = 1.0 / sqrt(d);
            s_w[i] = dsi[k];
        }
        __syncthreads();
#pragma unroll
        for (int w = 0; w < WT; ++w)
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const int i = tid + k * T;
                double v = 0.0;
                if (i < n && w < W) {
                    const size_t o = (size_t)w * ld + i;
                    for (int t = 0; t < op.n_terms; ++t) v = fma(s_theta[t], op.val[(size_t)t * W * ld + o], v);
                    v *= dsi[k] * s_w[op.col[o]];
                }
                s_val[w * np + i] = (float)v;
            }

        // ---- CG on the scaled system (fp32 vectors, fp64 scalars and solution)
        float r[R], p[R], q[R], s[R];
        double x[R];
        double bnorm2 = 0.0;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int i = tid + k * T;
            const double bt = (i < n ? op.rhs[i] : 0.0) * dsi[k];
            x[k] = 0.0;
            r[k] = (float)bt;
            s_r[i] = r[k];
        }
        __syncthreads();  // s_val and r visible; all gathers of dsi done
        spmv(r, s);
        double gam = 0.0, del = 0.0;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            gam = fma((double)r[k], (double)r[k], gam);
            del = fma((double)r[k], (double)s[k], del);
        }
        block_sum2(gam, del, s_part, lane, warp, nwarps);
        bnorm2 = gam;
        int status = TFIN_STATUS_MAXIT, it = 0;
        if (!(gam > 0.0) || !(del > 0.0)) {
            status = (gam == 0.0) ? TFIN_STATUS_CONVERGED : TFIN_STATUS_BREAKDOWN;
        } else {
            const double thresh = io.tol2 * gam;
            double alpha = gam / del, denom = del;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                p[k] = r[k];
                q[k] = s[k];
            }
            while (it < io.maxit) {
                ++it;
                const float af = (float)alpha;
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    x[k] = fma(alpha, (double)p[k], x[k]);
                    r[k] = fmaf(-af, q[k], r[k]);
                    s_r[tid + k * T] = r[k];
                }
                __syncthreads();
                spmv(r, s);
                double gn = 0.0, dl = 0.0;
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    gn = fma((double)r[k], (double)r[k], gn);
                    dl = fma((double)r[k], (double)s[k], dl);
                }
                block_sum2(gn, dl, s_part, lane, warp, nwarps);
                if (gn <= thresh) {
                    status = TFIN_STATUS_CONVERGED;
                    break;
                }
                const double beta = gn / gam;
                denom = dl - beta * beta * denom;
                if (!(denom > 0.0) || !(gn == gn)) {
                    status = TFIN_STATUS_BREAKDOWN;
                    break;
                }
                alpha = gn / denom;
                gam = gn;
                const float bf = (float)beta;
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    p[k] = fmaf(bf, p[k], r[k]);
                    q[k] = fmaf(bf, q[k], s[k]);
                }
            }
        }

        // ---- epilogue
        __syncthreads();
        if (io.relres_out) {  // ||b~ - A~ x~|| / ||b~|| with the fp32 operator, fp64 accumulation
#pragma unroll
            for (int k = 0; k < R; ++k) s_r[tid + k * T] = (float)x[k];
            __syncthreads();
            double rr = 0.0, dummy = 0.0;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const int i = tid + k * T;
                double ax = x[k];
#pragma unroll
                for (int w = 0; w < WT; ++w) ax = fma((double)s_val[w * np + i], (double)gather(k, w), ax);
                const double t = (i < n ? op.rhs[i] : 0.0) * s_w[i] - ax;
                rr = fma(t, t, rr);
            }
            block_sum2(rr, dummy, s_part, lane, warp, nwarps);
            const double relres = bnorm2 > 0.0 ? sqrt(rr / bnorm2) : sqrt(rr);
            if (!(relres == relres)) status = TFIN_STATUS_BREAKDOWN;
            if (tid == 0) io.relres_out[sample] = relres;
            __syncthreads();
        }
        if (tid == 0) {
            if (io.iters_out) io.iters_out[sample] = it;
            if (io.status_out) io.status_out[sample] = status;
        }
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int i = tid + k * T;
            x[k] *= s_w[i];  // w = D^-1/2 x~ (s_w still holds 1/sqrt(diag) of this thread's rows)
            s_w[i] = x[k];
        }
        __syncthreads();
        if (io.qoi_out) {
            for (int o = warp; o < obs.rows; o += nwarps) {
                double acc = 0.0;
                for (int j = obs.ptr[o] + lane; j < obs.ptr[o + 1]; j += 32) acc = fma(obs.val[j], s_w[obs.idx[j]], acc);
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
}

struct PcgF32Variant {
    int R, WT, maxT;
    const void* func;
};

inline const PcgF32Variant* pcg_f32_variants(int* count) {
    // register budgets keep r, p, q, s (fp32), x (fp64) and the packed offsets on chip: n <= 2048 (the reference mesh has
    // 1446 dofs); larger meshes use the fp64 kernels
    static const PcgF32Variant tab[] = {
        {4, 4, 416, (const void*)&pcg_f32_kernel<4, 4, 416, 2>},
        {5, 4, 320, (const void*)&pcg_f32_kernel<5, 4, 320, 2>},
        {8, 4, 256, (const void*)&pcg_f32_kernel<8, 4, 256, 2>},
        {6, 8, 288, (const void*)&pcg_f32_kernel<6, 8, 288, 2>},
        {8, 8, 256, (const void*)&pcg_f32_kernel<8, 8, 256, 2>},
        {8, 12, 256, (const void*)&pcg_f32_kernel<8, 12, 256, 2>},
    };
    *count = (int)(sizeof(tab) / sizeof(tab[0]));
    return tab;
}

}  // namespace tfin
