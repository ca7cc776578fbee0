// D1 / D2: batched sparse-direct (frontal Cholesky) solve of  A(sample) w = b,  qoi = B_obs w.
//
// The reference solves every sample with a sparse direct factorisation (dolfin `solve`, fom/forward_solve.py:286,
// rom/averaged_affine_ROM.py:256).  Here the symbolic work is shared by the whole batch (frontal_host.h compiles it into
// two sequential instruction streams) and only the numeric factorisation runs per sample, entirely on chip:
//
//   D1 frontal_lane_kernel   ONE SAMPLE PER THREAD (a warp = 32 samples in lock step).  The active front of every sample
//      is a packed triangle in shared memory laid out [entry][lane], so every access of the warp is one conflict-free
//      256-byte row; all indices (slots, addresses, program counters) are warp-uniform.  The pivot column lives in
//      registers (fully unrolled, uniform early exits), the factor columns stream to a per-warp workspace in HBM
//      ([entry][lane], coalesced) and come back -- prefetched one pivot ahead -- for the backward substitution, which
//      also accumulates the observables.  For fronts of up to 32 nodes (meshes up to ~3 k dofs).
//   D2 frontal_cta_kernel    ONE SAMPLE PER CTA for wide fronts (refined meshes, front 30-200 nodes): the front lives in
//      the CTA's shared memory, the rank-1 update of a pivot is spread over the threads, two barriers per pivot.
//      The observables need no backward substitution: qoi_o = (L^-1 B_obs[o])^T (L^-1 b), so the n_obs observation rows
//      ride along as extra right-hand sides of the forward elimination and the factor is never stored (QOI mode).  When
//      the full solution is requested the columns stream to HBM and a backward substitution follows (SOLVE mode).
//
// Both read the program through a shared-memory RING that cp.async (LDGSTS) keeps filled one ring ahead of the reader:
// with one or two resident warps per scheduler a dependent global load per pivot would cost a full L2 round trip each.
//
// Per-sample numeric values: A = sum over the assembly list of coef * cvec[term], cvec = [1, theta_1..theta_Q] for the
// affine operator (AffineROMFin._F, averaged_affine_ROM.py:156-162) or [1, cell coefficients] for the nodal operator
// (Fin._F, forward_solve.py:160-161).
#pragma once

#include "common.cuh"

namespace tfin {

struct FrontalDev {
    int n, nslots, cmax, ncv;      // ncv = length of the coefficient vector (incl. the leading 1)
    int ntri;                      // nslots (nslots + 1) / 2
    int ring_bytes;                // power of two >= 2 * largest record + 512
    long long nnzL;
    const unsigned char* fwd;      // forward stream (frontal_host.h: frontal_pack_streams)
    const unsigned char* bwd;      // backward stream
};

struct FrontalIO {
    const double* in;        // (N, in_stride): parameters theta (affine operator)
    long long N;
    int in_stride;
    int n_obs;
    double* w_out;           // (N, n) | null
    double* qoi_out;         // (N, n_obs) | null
    int* iters_out;          // (N) | null   (0: direct solve)
    int* status_out;         // (N) | null
    double* relres_out;      // (N) | null   |b.w - y.y| / y.y : consistency of the two substitutions
    unsigned long long* counter;
    double* work;            // factor workspace
    const double* cv_global; // nodal operator: coefficient vectors [group][ncv][32] (D1) or [sample][ncv] (D2); null = affine
};

__device__ __forceinline__ unsigned tri_u(unsigned s) { return s * (s + 1u) / 2u; }

// ---- instruction-stream ring.  All fields are uniform across the threads that share the ring.
struct StreamRing {
    unsigned char* buf;         // shared memory, mask + 1 bytes
    const unsigned char* src;   // global stream, zero padded by one ring + 512 bytes
    unsigned mask, fetched, rd;
    __device__ __forceinline__ void reset(const unsigned char* s) {
        src = s;
        fetched = 0;
        rd = 0;
    }
    // prefetch whole 512-byte chunks while they fit ahead of the reader; `issue` selects the lanes that copy (one warp)
    __device__ __forceinline__ void fill(int lane, bool issue) {
        while (fetched + 512u - rd <= mask + 1u) {
            if (issue) cp_async16(buf + ((fetched + lane * 16u) & mask), src + fetched + lane * 16u);
            fetched += 512u;
        }
        if (issue) cp_async_commit();
    }
    __device__ __forceinline__ unsigned u32(unsigned off) const {
        return *reinterpret_cast<const unsigned*>(buf + ((rd + off) & mask));
    }
    __device__ __forceinline__ unsigned u16(unsigned off) const {
        return *reinterpret_cast<const unsigned short*>(buf + ((rd + off) & mask));
    }
    __device__ __forceinline__ double f64(unsigned off) const {
        return *reinterpret_cast<const double*>(buf + ((rd + off) & mask));
    }
};

// ------------------------------------------------------------------------------------------------ D1
// shared memory per warp:  F[ntri][32] | yv[nslots][32] | qacc[n_obs][32] | cvec[ncv][32] (affine only) | ring
__host__ __device__ inline size_t frontal_lane_smem(int ntri, int nslots, int n_obs, int ncv_smem, int ring_bytes) {
    return (size_t)(ntri + nslots + n_obs + ncv_smem) * 32 * sizeof(double) + (size_t)ring_bytes;
}

template <int CM>
__global__ void __launch_bounds__(32) frontal_lane_kernel(FrontalDev P, FrontalIO io) {
    extern __shared__ __align__(16) double fsm[];
    const int lane = threadIdx.x;
    double* F = fsm + lane;                            // F[e * 32]
    double* yv = F + (size_t)P.ntri * 32;
    double* qacc = yv + (size_t)P.nslots * 32;
    double* cvs = qacc + (size_t)io.n_obs * 32;        // affine only
    StreamRing ring;
    ring.buf = reinterpret_cast<unsigned char*>(fsm + (size_t)(P.ntri + P.nslots + io.n_obs + (io.cv_global ? 0 : P.ncv)) * 32);
    ring.mask = (unsigned)P.ring_bytes - 1u;
    const unsigned full = 0xffffffffu;
    const long long n_groups = (io.N + 31) / 32;
    const int n = P.n;
    const size_t wstride = ((size_t)P.nnzL + 2 * (size_t)n) * 32;   // per-CTA workspace (doubles)
    double* Lw = io.work + (size_t)blockIdx.x * wstride + lane;
    double* RY = Lw + (size_t)P.nnzL * 32;

    for (;;) {
        long long g = 0;
        if (lane == 0) g = (long long)atomicAdd(io.counter, 1ULL);
        g = __shfl_sync(full, g, 0);
        if (g >= n_groups) break;
        const long long s = g * 32 + lane;
        const bool valid = s < io.N;
        const long long sc = valid ? s : io.N - 1;
        ring.reset(P.fwd);
        ring.fill(lane, true);
        const double* cv;      // coefficient vector of this lane: cv[t * 32]
        if (io.cv_global) {
            cv = io.cv_global + (size_t)g * P.ncv * 32 + lane;
        } else {
            cvs[0] = 1.0;
            for (int t = 1; t < P.ncv; ++t) cvs[t * 32] = io.in[sc * io.in_stride + (t - 1)];
            cv = cvs;
        }
        for (int e = 0; e < P.ntri; ++e) F[e * 32] = 0.0;
        for (int e = 0; e < P.nslots; ++e) yv[e * 32] = 0.0;
        for (int o = 0; o < io.n_obs; ++o) qacc[o * 32] = 0.0;

        bool bad = false;
        double yy = 0.0;
        size_t cp = 0;
        for (int j = -1; j < n; ++j) {
            cp_async_wait<0>();
            __syncwarp();
            ring.fill(lane, true);
            const unsigned c = ring.u32(0), p = ring.u32(4), npos = ring.u32(8), nent = ring.u32(12);
            const unsigned reclen = ring.u32(20);
            unsigned sl[CM];   // slots of the column (uniform values)
            double l[CM];      // scaled pivot column
            if (j >= 0) {
#pragma unroll
                for (int a = 0; a < CM; ++a) {
                    if (a >= (int)c) break;
                    sl[a] = ring.u16(32 + 2 * a);
                }
                const unsigned pd = tri_u(p) + p;
                const double dd = F[pd * 32];
                F[pd * 32] = 0.0;
                // gather the pivot column (independent loads first, then the zeroing stores)
#pragma unroll
                for (int a = 0; a < CM; ++a) {
                    if (a >= (int)c) break;
                    const unsigned ad = sl[a] > p ? tri_u(sl[a]) + p : tri_u(p) + sl[a];
                    l[a] = F[ad * 32];
                }
#pragma unroll
                for (int a = 0; a < CM; ++a) {
                    if (a >= (int)c) break;
                    const unsigned ad = sl[a] > p ? tri_u(sl[a]) + p : tri_u(p) + sl[a];
                    F[ad * 32] = 0.0;
                }
                bad |= !(dd > 0.0);
                const double rinv = rsqrt(dd);
                const double yp = (yv[p * 32] + ring.f64(24)) * rinv;
                yv[p * 32] = 0.0;
                yy = fma(yp, yp, yy);
                double* Lj = Lw + cp * 32;
                RY[(size_t)(2 * j) * 32] = rinv;
                RY[(size_t)(2 * j + 1) * 32] = yp;
#pragma unroll
                for (int a = 0; a < CM; ++a) {
                    if (a >= (int)c) break;
                    l[a] *= rinv;
                    Lj[a * 32] = l[a];
                    yv[sl[a] * 32] = fma(-l[a], yp, yv[sl[a] * 32]);
                }
                cp += c;
            }
            {   // assembly of column j + 1
                const unsigned off_pos = 32 + ((2 * c + 7) & ~7u), off_coef = off_pos + 8 * npos, off_term = off_coef + 8 * nent;
                unsigned e = 0;
                for (unsigned q = 0; q < npos; ++q) {
                    const unsigned ad = ring.u32(off_pos + 8 * q), cnt = ring.u32(off_pos + 8 * q + 4);
                    double sum = 0.0;
                    for (unsigned k = 0; k < cnt; ++k, ++e)
                        sum = fma(ring.f64(off_coef + 8 * e), cv[(size_t)ring.u32(off_term + 4 * e) * 32], sum);
                    F[ad * 32] += sum;
                }
            }
            if (j >= 0) {
                // rank-1 update of the front: F[tri(s_a) + s_b] -= l_a l_b, b <= a (slots ascend with the position);
                // a whole row is loaded before it is stored so that the loads overlap
#pragma unroll
                for (int a = 0; a < CM; ++a) {
                    if (a >= (int)c) break;
                    double* Fr = F + (size_t)tri_u(sl[a]) * 32;
                    double f[CM];
#pragma unroll
                    for (int b = 0; b <= a; ++b) f[b] = Fr[sl[b] * 32];
#pragma unroll
                    for (int b = 0; b <= a; ++b) Fr[sl[b] * 32] = fma(-l[a], l[b], f[b]);
                }
            }
            ring.rd += reclen;
        }
        // backward substitution L^T w = y (yv doubles as the slot-indexed solution), observables on the fly; the factor
        // column and (1/L_jj, y_j) of the next pivot are prefetched into registers while the current one is reduced
        ring.reset(P.bwd);
        ring.fill(lane, true);
        cp_async_wait<0>();
        __syncwarp();
        double bw = 0.0;
        double lA[CM], lB[CM], rA, yA, rB = 0.0, yB = 0.0;
        {
            const unsigned c0 = ring.u32(0);
            cp -= c0;
#pragma unroll
            for (int a = 0; a < CM; ++a) {
                if (a >= (int)c0) break;
                lA[a] = Lw[(cp + a) * 32];
            }
            rA = RY[(size_t)(2 * (n - 1)) * 32];
            yA = RY[(size_t)(2 * (n - 1) + 1) * 32];
        }
        auto bstep = [&](int j, double (&lc)[CM], double (&ln)[CM], double rc, double yc, double& rn, double& yn) {
            ring.fill(lane, true);
            const unsigned c = ring.u32(0), p = ring.u32(4), nobs = ring.u32(8), dof = ring.u32(12);
            const unsigned reclen = ring.u32(16), cnext = ring.u32(20);
            if (j > 0) {
                cp -= cnext;
#pragma unroll
                for (int a = 0; a < CM; ++a) {
                    if (a >= (int)cnext) break;
                    ln[a] = Lw[(cp + a) * 32];
                }
                rn = RY[(size_t)(2 * (j - 1)) * 32];
                yn = RY[(size_t)(2 * (j - 1) + 1) * 32];
            }
            double acc = yc;
#pragma unroll
            for (int a = 0; a < CM; ++a) {
                if (a >= (int)c) break;
                acc = fma(-lc[a], yv[ring.u16(32 + 2 * a) * 32], acc);
            }
            const double wj = acc * rc;
            yv[p * 32] = wj;
            bw = fma(ring.f64(24), wj, bw);
            if (io.w_out && valid) io.w_out[(size_t)s * n + dof] = wj;
            const unsigned off_val = 32 + ((2 * c + 7) & ~7u), off_row = off_val + 8 * nobs;
            for (unsigned o = 0; o < nobs; ++o) {
                const unsigned row = ring.u32(off_row + 4 * o);
                qacc[row * 32] = fma(ring.f64(off_val + 8 * o), wj, qacc[row * 32]);
            }
            ring.rd += reclen;
            cp_async_wait<0>();
            __syncwarp();
        };
        for (int j = n - 1; j >= 0; j -= 2) {
            bstep(j, lA, lB, rA, yA, rB, yB);
            if (j >= 1) bstep(j - 1, lB, lA, rB, yB, rA, yA);
        }
        if (valid) {
            if (io.qoi_out)
                for (int o = 0; o < io.n_obs; ++o) io.qoi_out[(size_t)s * io.n_obs + o] = qacc[o * 32];
            const bool nan = !(bw == bw);
            if (io.status_out) io.status_out[s] = (bad || nan) ? TFIN_STATUS_BREAKDOWN : TFIN_STATUS_CONVERGED;
            if (io.iters_out) io.iters_out[s] = 0;
            if (io.relres_out) io.relres_out[s] = fabs(bw - yy) / yy;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------ D2
#define FRONTAL_MODE_QOI 0    // observables only: n_obs extra right-hand sides, no factor storage, no backward pass
#define FRONTAL_MODE_SOLVE 1  // full solution: factor columns stream to HBM, backward substitution

struct FrontalCtaSmem {
    size_t F, yv, lcol, ypiv, red, qacc, cvec, cslot, ring, total;   // byte offsets
    static FrontalCtaSmem make(int ntri, int nslots, int cmax, int R, int n_obs, int ncv_smem, int ring_bytes) {
        FrontalCtaSmem L;
        size_t o = 0;
        auto take = [&](size_t bytes) {
            const size_t at = o;
            o += (bytes + 15) & ~(size_t)15;
            return at;
        };
        L.F = take((size_t)ntri * 8);
        L.yv = take((size_t)nslots * R * 8);
        L.lcol = take((size_t)cmax * 8);
        L.ypiv = take((size_t)R * 8);
        L.red = take(2 * 32 * 8);
        L.qacc = take((size_t)n_obs * 8);
        L.cvec = take((size_t)ncv_smem * 8);
        L.cslot = take((size_t)cmax * 4);
        L.ring = take((size_t)ring_bytes);
        L.total = o;
        return L;
    }
};

template <int MODE>
__global__ void frontal_cta_kernel(FrontalDev P, FrontalIO io, FrontalCtaSmem L) {
    extern __shared__ __align__(16) unsigned char fsm_raw[];
    double* F = reinterpret_cast<double*>(fsm_raw + L.F);
    double* yv = reinterpret_cast<double*>(fsm_raw + L.yv);
    double* lcol = reinterpret_cast<double*>(fsm_raw + L.lcol);
    double* ypiv = reinterpret_cast<double*>(fsm_raw + L.ypiv);
    double* red = reinterpret_cast<double*>(fsm_raw + L.red);
    double* qacc = reinterpret_cast<double*>(fsm_raw + L.qacc);
    double* cvs = reinterpret_cast<double*>(fsm_raw + L.cvec);
    unsigned* cslot = reinterpret_cast<unsigned*>(fsm_raw + L.cslot);
    __shared__ long long s_sample;
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = NT >> 5;
    const int R = MODE == FRONTAL_MODE_QOI ? 1 + io.n_obs : 1;
    const unsigned rmagic = (unsigned)((0x100000000ULL + R - 1) / R);   // a = umulhi(i, rmagic) == i / R for i < 2^20, R >= 2
    const int n = P.n;
    const int rtid = NT - 1 - tid;   // threads counted from the top take the right-hand sides and the assembly
    const bool loader = warp == 0;   // warp 0 keeps the ring filled
    StreamRing ring;
    ring.buf = fsm_raw + L.ring;
    ring.mask = (unsigned)P.ring_bytes - 1u;
    double* Lw = nullptr;
    double* RY = nullptr;
    if (MODE == FRONTAL_MODE_SOLVE) {
        Lw = io.work + (size_t)blockIdx.x * ((size_t)P.nnzL + 2 * (size_t)n);
        RY = Lw + (size_t)P.nnzL;
    }

    for (;;) {
        __syncthreads();
        if (tid == 0) s_sample = (long long)atomicAdd(io.counter, 1ULL);
        __syncthreads();
        const long long s = s_sample;
        if (s >= io.N) break;
        ring.reset(P.fwd);
        ring.fill(lane, loader);
        const double* cv;
        if (io.cv_global) {
            cv = io.cv_global + (size_t)s * P.ncv;
        } else {
            for (int t = tid; t < P.ncv; t += NT) cvs[t] = t == 0 ? 1.0 : io.in[s * io.in_stride + (t - 1)];
            cv = cvs;
        }
        for (int e = tid; e < P.ntri; e += NT) F[e] = 0.0;
        for (int e = tid; e < P.nslots * R; e += NT) yv[e] = 0.0;
        for (int o = tid; o < io.n_obs; o += NT) qacc[o] = 0.0;
        if (loader) cp_async_wait<0>();
        __syncthreads();

        bool bad = false;
        double myq = 0.0;   // QOI mode: thread rtid = 1 + o accumulates observable o; rtid = 0 accumulates y.y
        size_t cp = 0;
        for (int j = -1; j < n; ++j) {
            // ---- G_j: pivot, scaled column (zeroing the consumed entries), pivot row of the right-hand sides,
            //      and the assembly of column j + 1 (disjoint entries)
            ring.fill(lane, loader);
            const unsigned c = ring.u32(0), p = ring.u32(4), npos = ring.u32(8), nent = ring.u32(12), nobs = ring.u32(16);
            const unsigned reclen = ring.u32(20);
            const unsigned off_pos = 32 + ((2 * c + 7) & ~7u), off_coef = off_pos + 8 * npos, off_term = off_coef + 8 * nent;
            const unsigned off_oval = (off_term + 4 * nent + 7) & ~7u, off_orow = off_oval + 8 * nobs;
            const unsigned pd = tri_u(p) + p;
            double rinv = 0.0;
            if (j >= 0) {
                const double dd = F[pd];
                bad |= !(dd > 0.0);
                rinv = rsqrt(dd);
                for (int a = tid; a < (int)c; a += NT) {
                    const unsigned sa = ring.u16(32 + 2 * a);
                    const unsigned ad = sa > p ? tri_u(sa) + p : tri_u(p) + sa;
                    const double l = F[ad] * rinv;
                    F[ad] = 0.0;
                    lcol[a] = l;
                    cslot[a] = sa;
                    if (MODE == FRONTAL_MODE_SOLVE) Lw[cp + a] = l;
                }
                if (rtid < R) {
                    double v = yv[p * R + rtid];
                    yv[p * R + rtid] = 0.0;
                    if (rtid == 0) {
                        v += ring.f64(24);
                    } else {
                        for (unsigned o = 0; o < nobs; ++o)
                            if ((int)ring.u32(off_orow + 4 * o) == rtid - 1) v += ring.f64(off_oval + 8 * o);
                    }
                    ypiv[rtid] = v * rinv;
                }
            }
            for (unsigned q = (unsigned)rtid; q < npos; q += NT) {   // threads from the top: one assembly position each
                unsigned e = 0;
                for (unsigned q2 = 0; q2 < q; ++q2) e += ring.u32(off_pos + 8 * q2 + 4);
                const unsigned cnt = ring.u32(off_pos + 8 * q + 4);
                double sum = 0.0;
                for (unsigned k = 0; k < cnt; ++k, ++e) sum = fma(ring.f64(off_coef + 8 * e), cv[ring.u32(off_term + 4 * e)], sum);
                F[ring.u32(off_pos + 8 * q)] += sum;
            }
            __syncthreads();
            // ---- U_j: rank-1 update of the front and of the right-hand sides
            if (j >= 0) {
                if (rtid < R) {
                    const double y0 = ypiv[0];
                    myq = fma(ypiv[rtid], y0, myq);
                    if (MODE == FRONTAL_MODE_SOLVE && rtid == 0) {
                        RY[2 * (size_t)j] = rinv;
                        RY[2 * (size_t)j + 1] = y0;
                    }
                }
                if (tid == 0) F[pd] = 0.0;
                for (int i = tid; i < (int)c * R; i += NT) {
                    const int a = R == 1 ? i : (int)__umulhi((unsigned)i, rmagic), r = i - a * R;
                    const unsigned ys = cslot[a] * R + r;
                    yv[ys] = fma(-lcol[a], ypiv[r], yv[ys]);
                }
                // folded triangle: combined row q = (row c-1-q, then row q) has c + 1 elements for every q
                const int half = ((int)c + 1) >> 1;
                for (int q = warp; q < half; q += nw) {
                    const int rowA = (int)c - 1 - q, rowB = q;
                    const double laA = lcol[rowA], laB = lcol[rowB];
                    const unsigned trA = tri_u(cslot[rowA]), trB = tri_u(cslot[rowB]);
                    const int xend = rowA == rowB ? rowA : (int)c;
                    for (int x = lane; x <= xend; x += 32) {
                        const bool first = x <= rowA;
                        const int b = first ? x : x - rowA - 1;
                        const unsigned ad = (first ? trA : trB) + cslot[b];
                        F[ad] = fma(-(first ? laA : laB), lcol[b], F[ad]);
                    }
                }
            }
            if (loader) cp_async_wait<0>();   // the chunks requested at the top of this step have had the whole step to land
            __syncthreads();
            cp += (j >= 0 ? c : 0u);
            ring.rd += reclen;
        }
        if (MODE == FRONTAL_MODE_QOI) {
            if (rtid >= 1 && rtid < R && io.qoi_out) io.qoi_out[(size_t)s * io.n_obs + (rtid - 1)] = myq;
            if (tid == 0) {
                if (io.status_out) io.status_out[s] = bad ? TFIN_STATUS_BREAKDOWN : TFIN_STATUS_CONVERGED;
                if (io.iters_out) io.iters_out[s] = 0;
                if (io.relres_out) io.relres_out[s] = 0.0;
            }
            continue;
        }
        // ---- SOLVE mode: backward substitution L^T w = y; yv (R = 1) doubles as the slot-indexed solution
        if (rtid == 0) ypiv[0] = myq;   // y.y, read by thread 0 after the loop (barriers inside)
        ring.reset(P.bwd);
        ring.fill(lane, loader);
        if (loader) cp_async_wait<0>();
        __syncthreads();
        double bw = 0.0;
        // factor entries of the next pivot are prefetched while the current one is reduced (thread a holds entry a, a + NT, ..)
        constexpr int LPF = 2;   // columns of up to LPF * NT entries are prefetched, longer ones read in place
        double lpre[LPF], rcur, ycur;
        {
            const unsigned c0 = ring.u32(0);
            cp -= c0;
#pragma unroll
            for (int u = 0; u < LPF; ++u) lpre[u] = tid + u * NT < (int)c0 ? Lw[cp + tid + u * NT] : 0.0;
            rcur = RY[2 * (size_t)(n - 1)];
            ycur = RY[2 * (size_t)(n - 1) + 1];
        }
        for (int j = n - 1; j >= 0; --j) {
            ring.fill(lane, loader);
            const unsigned c = ring.u32(0), p = ring.u32(4), nobs = ring.u32(8), dof = ring.u32(12);
            const unsigned reclen = ring.u32(16), cnext = ring.u32(20);
            const size_t cpj = cp;
            double lnext[LPF], rnext = 0.0, ynext = 0.0;
#pragma unroll
            for (int u = 0; u < LPF; ++u) lnext[u] = 0.0;
            if (j > 0) {
                cp -= cnext;
#pragma unroll
                for (int u = 0; u < LPF; ++u) lnext[u] = tid + u * NT < (int)cnext ? Lw[cp + tid + u * NT] : 0.0;
                rnext = RY[2 * (size_t)(j - 1)];
                ynext = RY[2 * (size_t)(j - 1) + 1];
            }
            double part = 0.0;
#pragma unroll
            for (int u = 0; u < LPF; ++u)
                if (tid + u * NT < (int)c) part = fma(lpre[u], yv[ring.u16(32 + 2 * (tid + u * NT))], part);
            for (int a = tid + LPF * NT; a < (int)c; a += NT) part = fma(Lw[cpj + a], yv[ring.u16(32 + 2 * a)], part);
            part = warp_sum(part);
            double* rd = red + (j & 1) * 32;
            if (lane == 0) rd[warp] = part;
            __syncthreads();
            double acc = 0.0;
            for (int w = 0; w < nw; ++w) acc += rd[w];
            const double wj = (ycur - acc) * rcur;
            if (tid == 0) {
                yv[p] = wj;
                bw = fma(ring.f64(24), wj, bw);
                if (io.w_out) io.w_out[(size_t)s * n + dof] = wj;
                const unsigned off_val = 32 + ((2 * c + 7) & ~7u), off_row = off_val + 8 * nobs;
                for (unsigned o = 0; o < nobs; ++o) qacc[ring.u32(off_row + 4 * o)] += ring.f64(off_val + 8 * o) * wj;
            }
#pragma unroll
            for (int u = 0; u < LPF; ++u) lpre[u] = lnext[u];
            rcur = rnext;
            ycur = ynext;
            if (loader) cp_async_wait<0>();
            __syncthreads();
            ring.rd += reclen;
        }
        if (tid == 0) {
            if (io.qoi_out)
                for (int o = 0; o < io.n_obs; ++o) io.qoi_out[(size_t)s * io.n_obs + o] = qacc[o];
            if (io.status_out) io.status_out[s] = (bad || !(bw == bw)) ? TFIN_STATUS_BREAKDOWN : TFIN_STATUS_CONVERGED;
            if (io.iters_out) io.iters_out[s] = 0;
            if (io.relres_out) io.relres_out[s] = fabs(bw - ypiv[0]) / ypiv[0];
        }
    }
}

// ------------------------------------------------------------------------------------------------ nodal coefficients
// Coefficient vectors of the nodal operator, cv[0] = 1, cv[1 + e] = cell coefficient of cell e (mean of k, or the
// quadrature of exp(k): cell_coefficient, pcg_small.cuh).  lane_major != 0: [group][ncv][32] for D1, else [sample][ncv].
// Block (32, 8): a 32-sample x 32-cell tile goes through shared memory so that both the reads of k (along a sample's
// row) and the writes (along the fastest output axis) are coalesced.
__global__ void frontal_cellcoef_kernel(const double* __restrict__ k, long long N, int n, int n_cells,
                                        const int* __restrict__ cells, int coef_mode, int lane_major,
                                        double* __restrict__ cv) {
    __shared__ double tile[32][33];
    const int ncv = n_cells + 1;
    const long long g = blockIdx.y;
    const int c0 = blockIdx.x * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    {   // x = cell, y = sample
        const int c = c0 + tx;
        int v0 = 0, v1 = 0, v2 = 0;
        if (c < n_cells) {
            v0 = cells[3 * c];
            v1 = cells[3 * c + 1];
            v2 = cells[3 * c + 2];
        }
        for (int sy = ty; sy < 32; sy += 8) {
            const long long s = g * 32 + sy;
            double val = 0.0;
            if (c < n_cells && s < N) {
                const double* row = k + (size_t)s * n;
                val = cell_coefficient(coef_mode, row[v0], row[v1], row[v2]);
            }
            tile[sy][tx] = val;
        }
    }
    __syncthreads();
    if (lane_major) {   // x = sample lane, y = cell
        for (int cy = ty; cy < 32; cy += 8) {
            const int c = c0 + cy;
            if (c < n_cells) cv[((size_t)g * ncv + 1 + c) * 32 + tx] = tile[tx][cy];
        }
        if (blockIdx.x == 0 && ty == 0) cv[(size_t)g * ncv * 32 + tx] = 1.0;
    } else {            // x = cell, y = sample
        for (int sy = ty; sy < 32; sy += 8) {
            const long long s = g * 32 + sy;
            const int c = c0 + tx;
            if (s < N && c < n_cells) cv[(size_t)s * ncv + 1 + c] = tile[sy][tx];
            if (s < N && blockIdx.x == 0 && tx == 0) cv[(size_t)s * ncv] = 1.0;
        }
    }
}

}  // namespace tfin
