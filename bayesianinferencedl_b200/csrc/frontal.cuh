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
    int ring_bytes;                // power of two, see frontal_pack_streams
    int lr_rows;                   // D1: rows of the factor-row ring of the backward substitution
    int lanes;                     // D1: samples per warp (<= 32): fewer lanes = narrower rows = more resident warps
    int ring_fwd1;                 // D1 factor kernel: its own (smaller) instruction ring, bytes (power of two)
    long long nnzL;
    const unsigned char* fwd;      // forward stream (frontal_host.h: frontal_pack_streams)
    const unsigned char* bwd;      // backward stream
    const unsigned char* fsub;     // D1: forward-substitution stream (adjoint right-hand sides)
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
    // D1 adjoint pass (PHASE_FSUB): right-hand side -B_obs^T (qoi - data); qoi_out is READ, cost_out = 0.5 |qoi - data|^2
    const double* data;      // (1 | N, n_obs)
    long long data_stride;   // 0: one shared observation vector
    double* cost_out;        // (N) | null
    int unit_row;            // data == null: right-hand side -B_obs[unit_row, :]^T (one row of the sensitivity)
    // D2 adjoint solve (MODE_SOLVE, adj != 0): the matrix is factorised again with the right-hand side
    // -B_obs^T res, res = qoi_in - data (adj = 1) or the unit vector e_unit_row (adj = 2); w_out receives the adjoint state
    int adj;
    const double* qoi_in;    // (N, n_obs), adj = 1
};

__device__ __forceinline__ unsigned tri_u(unsigned s) { return s * (s + 1u) / 2u; }

// ---- instruction-stream ring.  All fields are uniform across the threads that share the ring.
struct StreamRing {
    unsigned char* buf;         // shared memory, mask + 1 bytes
    const unsigned char* src;   // global stream, zero padded by one ring + 512 bytes
    unsigned mask, fetched, rd;
    unsigned buf_s;             // shared-window address of buf, formed once (the conversion reads a special register)
    __device__ __forceinline__ void reset(const unsigned char* s) {
        src = s;
        fetched = 0;
        rd = 0;
        buf_s = smem_u32(buf);
    }
    // prefetch whole 512-byte chunks while they fit ahead of the reader; `issue` selects the lanes that copy (one warp).
    // The caller commits the copy group.
    __device__ __forceinline__ void fill(int lane, bool issue) {
        while (fetched + 512u - rd <= mask + 1u) {
            if (issue)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(buf_s + ((fetched + lane * 16u) & mask)),
                             "l"(src + fetched + lane * 16u) : "memory");
            fetched += 512u;
        }
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
// shared memory per warp:  F[ntri][32] | yv[nslots][32] | qacc[n_obs][32] | cvec[ncv][32] (affine only) |
//                          factor-row ring [lr_rows][32] (backward substitution) | instruction ring
// phase: 0 = factorisation and backward substitution in one kernel, 1 = factorisation + forward elimination only (factor
// blocks of every group stay in HBM), 2 = backward substitution + observables only
#define FRONTAL_PHASE_BOTH 0
#define FRONTAL_PHASE_FACTOR 1
#define FRONTAL_PHASE_BSUB 2
#define FRONTAL_PHASE_FSUB 3   // forward substitution L y = -B_obs^T (qoi - data) with the stored factor (adjoint solve)
__host__ __device__ inline size_t frontal_lane_smem(int ntri, int nslots, int n_obs, int ncv_smem, int lr_rows, int lanes,
                                                    int ring_bytes, int phase) {
    const bool sub = phase == FRONTAL_PHASE_BSUB || phase == FRONTAL_PHASE_FSUB;
    const int rows = (sub ? 0 : ntri) + nslots + (phase == FRONTAL_PHASE_FACTOR ? 0 : n_obs) + (sub ? 0 : ncv_smem) +
                     (phase == FRONTAL_PHASE_FACTOR ? 0 : lr_rows);
    return (((size_t)rows * lanes * sizeof(double) + 15) & ~(size_t)15) + (size_t)ring_bytes;   // ring_bytes: of this phase
}

#define FRONTAL_DMAX 7   // the backward substitution prefetches the factor rows of up to DMAX pivots ahead

__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_dyn(unsigned k) {   // k <= FRONTAL_DMAX, warp-uniform
    switch (k) {
        case 0: cp_async_wait<0>(); break;
        case 1: cp_async_wait<1>(); break;
        case 2: cp_async_wait<2>(); break;
        case 3: cp_async_wait<3>(); break;
        case 4: cp_async_wait<4>(); break;
        case 5: cp_async_wait<5>(); break;
        case 6: cp_async_wait<6>(); break;
        default: cp_async_wait<7>(); break;
    }
}

// Split launch (PHASE 1 then PHASE 2): the factorisation needs the front in shared memory (three warps of 27 samples per SM at n = 1597),
// the substitution only a right-hand side and the factor-block ring, so eight or more of its warps fit -- and both are
// bound by instruction latency, i.e. by resident warps.  The factor blocks of all groups of a chunk then live in HBM
// (workspace row block g), plus two rows per group for y.y and the breakdown flag.
template <int CM, int PHASE>
__global__ void __launch_bounds__(32) frontal_lane_kernel(FrontalDev P, FrontalIO io) {
    static_assert(CM % 4 == 0, "columns are read four entries at a time");
    extern __shared__ __align__(16) double fsm[];
    // A warp carries LPG = P.lanes <= 32 samples; with fewer than 32 the upper lanes shadow the lower ones (same
    // addresses, same values), which costs nothing but lets more warps share the SM's shared memory: the kernel is bound
    // by instruction latency, i.e. by samples in flight = resident warps x LPG, and shared memory bounds that product
    // (n = 1597: 3 warps x 27 samples instead of 2 x 32).
    // Shadow lanes all take the LAST sample's column: same address as lane LPG - 1, i.e. a broadcast -- any other choice
    // (say xlane % LPG) puts a second, different word on banks the real lanes use and costs a third wavefront per access.
    const int LPG = P.lanes, xlane = threadIdx.x;
    const int lane = xlane < LPG ? xlane : LPG - 1;
    const unsigned rb = 8u * (unsigned)LPG;            // bytes of one [row][lane] row
    // byte-addressed views of this lane's column of every [row][lane] array
    char* F = reinterpret_cast<char*>(fsm + lane);
    constexpr bool SUB = PHASE == FRONTAL_PHASE_BSUB || PHASE == FRONTAL_PHASE_FSUB;   // substitution-only launch
    char* yv = F + (size_t)(SUB ? 0 : P.ntri) * rb;
    char* qacc = yv + (size_t)P.nslots * rb;
    char* cvs = qacc + (size_t)(PHASE == FRONTAL_PHASE_FACTOR ? 0 : io.n_obs) * rb;        // affine only
    char* Lring = cvs + (size_t)((io.cv_global || SUB) ? 0 : P.ncv) * rb;
    StreamRing ring;
    {
        const size_t off = (size_t)((Lring - lane * 8) - reinterpret_cast<char*>(fsm)) +
                           (size_t)(PHASE == FRONTAL_PHASE_FACTOR ? 0 : P.lr_rows) * rb;
        ring.buf = reinterpret_cast<unsigned char*>(fsm) + ((off + 15) & ~(size_t)15);
    }
    ring.mask = (unsigned)(PHASE == FRONTAL_PHASE_FACTOR ? P.ring_fwd1 : P.ring_bytes) - 1u;
    const unsigned full = 0xffffffffu;
    const long long n_groups = (io.N + LPG - 1) / LPG;
    const int n = P.n;
    const size_t wrows = (size_t)P.nnzL + 2 * (size_t)n;   // workspace rows: per pivot [1/L_jj, y_j, column]
    // one workspace per CTA (both phases in one kernel) or per group (split launch; + 2 rows: y.y, breakdown flag)
    char* Lw0 = reinterpret_cast<char*>(io.work + lane);
    auto ld = [](const char* base, unsigned off) { return *reinterpret_cast<const double*>(base + off); };
    auto st = [](char* base, unsigned off, double v) { *reinterpret_cast<double*>(base + off) = v; };

    for (;;) {
        long long g = 0;
        if (xlane == 0) g = (long long)atomicAdd(io.counter, 1ULL);
        g = __shfl_sync(full, g, 0);
        if (g >= n_groups) break;
        const long long s = g * LPG + lane;
        const bool valid = s < io.N && xlane < LPG;
        const long long sc = s < io.N ? s : io.N - 1;
        char* Lw = Lw0 + (PHASE == FRONTAL_PHASE_BOTH ? (size_t)blockIdx.x * wrows : (size_t)g * (wrows + 2)) * rb;
        bool bad = false;
        double yy = 0.0;
        if (!SUB) {
        ring.reset(P.fwd);
        ring.fill(xlane, true);
        cp_async_commit();
        const char* cv;      // coefficient vector of this lane, row t at byte offset 256 t
        if (io.cv_global) {
            cv = reinterpret_cast<const char*>(io.cv_global + (size_t)g * P.ncv * LPG + lane);
        } else {
            st(cvs, 0, 1.0);
            for (int t = 1; t < P.ncv; ++t) st(cvs, rb * t, io.in[sc * io.in_stride + (t - 1)]);
            cv = cvs;
        }
        for (int e = 0; e < P.ntri; ++e) st(F, rb * e, 0.0);
        for (int e = 0; e < P.nslots; ++e) st(yv, rb * e, 0.0);
        if (PHASE == FRONTAL_PHASE_BOTH)
            for (int o = 0; o < io.n_obs; ++o) st(qacc, rb * o, 0.0);

        char* Lj = Lw;   // workspace block of the current pivot
        for (int j = -1; j < n; ++j) {
            cp_async_wait<0>();
            __syncwarp();
            ring.fill(xlane, true);
            cp_async_commit();
            const unsigned char* rec = ring.buf + (ring.rd & ring.mask);   // records never straddle the wrap point
            const uint4 h0 = *reinterpret_cast<const uint4*>(rec);        // c | 256 p | 256 (tri(p) + p) | npos
            const uint4 h1 = *reinterpret_cast<const uint4*>(rec + 16);   // nent | bytes | rhs
            const unsigned c = h0.x, npos = h0.w, nent = h1.x, reclen = h1.y;
            const unsigned c4 = (c + 3u) & ~3u;
            unsigned sr[CM], sc_[CM];   // 256 tri(s_a), 256 s_a  (uniform values)
            double l[CM];               // scaled pivot column
            if (j >= 0) {
                const double dd = ld(F, h0.z);
                const double ypre = ld(yv, h0.y);
                st(F, h0.z, 0.0);
                st(yv, h0.y, 0.0);
                // gather the pivot column and the right-hand-side entries it updates (independent loads first)
                double yo[CM];
#pragma unroll
                for (int q = 0; q < CM / 4; ++q) {
                    if (4 * q >= (int)c) break;
                    const uint4 ga = *reinterpret_cast<const uint4*>(rec + 32 + 16 * q);
                    const uint4 r4 = *reinterpret_cast<const uint4*>(rec + 32 + 4 * c4 + 16 * q);
                    const uint4 c4v = *reinterpret_cast<const uint4*>(rec + 32 + 8 * c4 + 16 * q);
                    sr[4 * q] = r4.x; sr[4 * q + 1] = r4.y; sr[4 * q + 2] = r4.z; sr[4 * q + 3] = r4.w;
                    sc_[4 * q] = c4v.x; sc_[4 * q + 1] = c4v.y; sc_[4 * q + 2] = c4v.z; sc_[4 * q + 3] = c4v.w;
                    // entries past c read address 0 / slot 0: harmless loads, never stored back
                    l[4 * q] = ld(F, ga.x); l[4 * q + 1] = ld(F, ga.y); l[4 * q + 2] = ld(F, ga.z); l[4 * q + 3] = ld(F, ga.w);
                    yo[4 * q] = ld(yv, c4v.x); yo[4 * q + 1] = ld(yv, c4v.y); yo[4 * q + 2] = ld(yv, c4v.z); yo[4 * q + 3] = ld(yv, c4v.w);
                    st(F, ga.x, 0.0);
                    if (4 * q + 1 < (int)c) st(F, ga.y, 0.0);
                    if (4 * q + 2 < (int)c) st(F, ga.z, 0.0);
                    if (4 * q + 3 < (int)c) st(F, ga.w, 0.0);
                }
                bad |= !(dd > 0.0);
                const double rinv = rsqrt(dd);
                const double yp = (ypre + __longlong_as_double(((long long)h1.w << 32) | h1.z)) * rinv;
                yy = fma(yp, yp, yy);
                st(Lj, 0, rinv);
                st(Lj, rb, yp);
#pragma unroll
                for (int a = 0; a < CM; ++a) {
                    if (a >= (int)c) break;
                    l[a] *= rinv;
                    st(Lj, rb * (2u + a), l[a]);
                    st(yv, sc_[a], fma(-l[a], yp, yo[a]));
                }
                Lj += (size_t)(c + 2) * rb;
            }
            {   // assembly of column j + 1: positions in chunks of 4 (loads first), the first two entries of a position unrolled
                const unsigned char* pos = rec + 32 + 12 * c4;
                const unsigned char* coef = pos + ((8 * npos + 15) & ~15u);
                const unsigned char* term = coef + ((8 * nent + 15) & ~15u);
                unsigned e0 = 0;
                for (unsigned q0 = 0; q0 < npos; q0 += 4) {
                    unsigned ad[4], cnt[4], eb[4];
                    double sum[4], fv[4];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const uint4 pp = *reinterpret_cast<const uint4*>(pos + 8 * (q0 + 2 * u));   // padded: reads past npos are zeros/other fields, masked below
                        ad[2 * u] = pp.x; cnt[2 * u] = q0 + 2 * u < npos ? pp.y : 0u;
                        ad[2 * u + 1] = pp.z; cnt[2 * u + 1] = q0 + 2 * u + 1 < npos ? pp.w : 0u;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        eb[u] = e0;
                        e0 += cnt[u];
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (q0 + u >= npos) break;
                        fv[u] = ld(F, ad[u]);
                        double s0 = 0.0, s1 = 0.0;
                        if (cnt[u] > 0) s0 = *reinterpret_cast<const double*>(coef + 8 * eb[u]) * ld(cv, *reinterpret_cast<const unsigned*>(term + 4 * eb[u]));
                        if (cnt[u] > 1) s1 = *reinterpret_cast<const double*>(coef + 8 * eb[u] + 8) * ld(cv, *reinterpret_cast<const unsigned*>(term + 4 * eb[u] + 4));
                        sum[u] = s0 + s1;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (q0 + u >= npos) break;
                        for (unsigned k = 2; k < cnt[u]; ++k)
                            sum[u] = fma(*reinterpret_cast<const double*>(coef + 8 * (eb[u] + k)),
                                         ld(cv, *reinterpret_cast<const unsigned*>(term + 4 * (eb[u] + k))), sum[u]);
                        st(F, ad[u], fv[u] + sum[u]);
                    }
                }
            }
            if (j >= 0) {
                // rank-1 update of the front: F[tri(s_a) + s_b] -= l_a l_b, b <= a (slots ascend with the position);
                // a whole row is loaded before it is stored so that the loads overlap
#pragma unroll
                for (int a = 0; a < CM; ++a) {
                    if (a >= (int)c) break;
                    char* Fr = F + sr[a];
                    double f[CM];
#pragma unroll
                    for (int b = 0; b <= a; ++b) f[b] = ld(Fr, sc_[b]);
#pragma unroll
                    for (int b = 0; b <= a; ++b) st(Fr, sc_[b], fma(-l[a], l[b], f[b]));
                }
            }
            ring.rd += reclen;
        }
        if (PHASE == FRONTAL_PHASE_FACTOR) {   // hand y.y and the breakdown flag to the substitution kernel
            st(Lw, (unsigned)(wrows * rb), yy);
            st(Lw, (unsigned)((wrows + 1) * rb), bad ? 1.0 : 0.0);
            cp_async_wait<0>();
            __syncwarp();
            continue;
        }
        } else {   // substitution launch: right-hand side slots start empty, y.y / flag come from the factor kernel
            for (int e = 0; e < P.nslots; ++e) st(yv, rb * e, 0.0);
            yy = ld(Lw, (unsigned)(wrows * rb));
            bad = ld(Lw, (unsigned)((wrows + 1) * rb)) != 0.0;
            if (PHASE == FRONTAL_PHASE_FSUB) {   // residual of the observables, kept where the other pass accumulates them
                if (io.data) {
                    const double* d = io.data + (io.data_stride ? sc * io.data_stride : 0);
                    double cost = 0.0;
                    for (int o = 0; o < io.n_obs; ++o) {
                        const double r = io.qoi_out[(size_t)sc * io.n_obs + o] - d[o];
                        st(qacc, rb * o, r);
                        cost = fma(r, r, cost);
                    }
                    if (io.cost_out && valid) io.cost_out[s] = 0.5 * cost;
                } else {
                    for (int o = 0; o < io.n_obs; ++o) st(qacc, rb * o, o == io.unit_row ? 1.0 : 0.0);
                }
            } else {
                for (int o = 0; o < io.n_obs; ++o) st(qacc, rb * o, 0.0);
            }
        }
        // ---- backward substitution L^T w = y (yv doubles as the slot-indexed solution), observables on the fly.  The
        // factor blocks [1/L_jj, y_j, column] come back from HBM through a block ring in shared memory that cp.async fills
        // up to DMAX pivots ahead; the host simulated that ring and wrote into every record which blocks to request and
        // how many of the youngest copy groups may still be in flight when the record is consumed (kw).
        cp_async_wait<0>();
        __syncwarp();
        ring.reset(PHASE == FRONTAL_PHASE_FSUB ? P.fsub : P.bwd);
        ring.fill(xlane, true);
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        double bw = 0.0;
        const unsigned lring_s = smem_u32(Lring);
        auto request = [&](const unsigned char* rq, unsigned count) {
            for (unsigned i = 0; i < count; ++i) {
                const uint4 r = *reinterpret_cast<const uint4*>(rq + 16 * i);   // rb x ring row | rows | source row
                unsigned dst = lring_s + r.x;
                const char* src = Lw + (size_t)r.z * rb;
                if (xlane < LPG)   // shadow lanes would write lane LPG - 1's word again: serialised as bank conflicts
                    for (unsigned k = 0; k < r.y; ++k, dst += rb, src += rb)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
            }
        };
        for (int j = n; j >= 0; --j) {   // step n - 1 - j of the stream; j == n: prologue record (initial requests only)
            ring.fill(xlane, true);
            const unsigned char* rec = ring.buf + (ring.rd & ring.mask);
            const uint4 h0 = *reinterpret_cast<const uint4*>(rec);        // c | rb p | nobs | dof
            const uint4 h1 = *reinterpret_cast<const uint4*>(rec + 16);   // bytes | npf, kw | rhs
            const unsigned c = h0.x, nobs = h0.z, npf = h1.y & 0xffffu, kw = h1.y >> 16;
            request(rec + 48, npf);
            cp_async_commit();
            if (j < n) {
                cp_async_wait_dyn(kw);
                __syncwarp();
                const char* blk = Lring + *reinterpret_cast<const unsigned*>(rec + 32);
                const unsigned char* cols = rec + 48 + 16 * npf;
                const double rinv = ld(blk, 0);
                if (PHASE == FRONTAL_PHASE_FSUB) {
                    // forward substitution with the adjoint right-hand side: y_j = (pending + rhs_j) / L_jj, then the
                    // column pushes it down; y_j replaces the first solve's y_j in the factor block (HBM)
                    const unsigned char* oval = cols + 4 * ((c + 3u) & ~3u);
                    const unsigned char* orow = oval + ((8 * nobs + 15) & ~15u);
                    double rhs = ld(yv, h0.y);
                    st(yv, h0.y, 0.0);
                    for (unsigned o = 0; o < nobs; ++o)
                        rhs = fma(-*reinterpret_cast<const double*>(oval + 8 * o), ld(qacc, *reinterpret_cast<const unsigned*>(orow + 4 * o)), rhs);
                    const double yj = rhs * rinv;
                    st(Lw, (*reinterpret_cast<const unsigned*>(rec + 36) + 1u) * rb, yj);
#pragma unroll
                    for (int q = 0; q < CM / 4; ++q) {
                        if (4 * q >= (int)c) break;
                        const uint4 cc = *reinterpret_cast<const uint4*>(cols + 16 * q);
                        const double l0 = ld(blk, rb * (2u + 4 * q)), w0 = ld(yv, cc.x);
                        const double l1 = 4 * q + 1 < (int)c ? ld(blk, rb * (3u + 4 * q)) : 0.0, w1 = ld(yv, cc.y);
                        const double l2 = 4 * q + 2 < (int)c ? ld(blk, rb * (4u + 4 * q)) : 0.0, w2 = ld(yv, cc.z);
                        const double l3 = 4 * q + 3 < (int)c ? ld(blk, rb * (5u + 4 * q)) : 0.0, w3 = ld(yv, cc.w);
                        st(yv, cc.x, fma(-l0, yj, w0));
                        if (4 * q + 1 < (int)c) st(yv, cc.y, fma(-l1, yj, w1));
                        if (4 * q + 2 < (int)c) st(yv, cc.z, fma(-l2, yj, w2));
                        if (4 * q + 3 < (int)c) st(yv, cc.w, fma(-l3, yj, w3));
                    }
                    ring.rd += h1.x;
                    continue;
                }
                double acc = ld(blk, rb);
#pragma unroll
                for (int q = 0; q < CM / 4; ++q) {
                    if (4 * q >= (int)c) break;
                    const uint4 cc = *reinterpret_cast<const uint4*>(cols + 16 * q);   // padded entries: slot 0, masked
                    const double l0 = ld(blk, rb * (2u + 4 * q)), w0 = ld(yv, cc.x);
                    const double l1 = 4 * q + 1 < (int)c ? ld(blk, rb * (3u + 4 * q)) : 0.0, w1 = ld(yv, cc.y);
                    const double l2 = 4 * q + 2 < (int)c ? ld(blk, rb * (4u + 4 * q)) : 0.0, w2 = ld(yv, cc.z);
                    const double l3 = 4 * q + 3 < (int)c ? ld(blk, rb * (5u + 4 * q)) : 0.0, w3 = ld(yv, cc.w);
                    acc = fma(-l0, w0, acc);
                    acc = fma(-l1, w1, acc);
                    acc = fma(-l2, w2, acc);
                    acc = fma(-l3, w3, acc);
                }
                const double wj = acc * rinv;
                st(yv, h0.y, wj);
                bw = fma(__longlong_as_double(((long long)h1.w << 32) | h1.z), wj, bw);
                if (io.w_out && valid) io.w_out[(size_t)s * n + h0.w] = wj;
                const unsigned char* oval = cols + 4 * ((c + 3u) & ~3u);
                const unsigned char* orow = oval + ((8 * nobs + 15) & ~15u);
                for (unsigned o = 0; o < nobs; ++o) {
                    const unsigned row = *reinterpret_cast<const unsigned*>(orow + 4 * o);
                    st(qacc, row, fma(*reinterpret_cast<const double*>(oval + 8 * o), wj, ld(qacc, row)));
                }
            }
            ring.rd += h1.x;
        }
        if (valid && PHASE != FRONTAL_PHASE_FSUB) {
            if (io.qoi_out)
                for (int o = 0; o < io.n_obs; ++o) io.qoi_out[(size_t)s * io.n_obs + o] = ld(qacc, rb * o);
            const bool nan = !(bw == bw);
            if (io.status_out) io.status_out[s] = (bad || nan) ? TFIN_STATUS_BREAKDOWN : TFIN_STATUS_CONVERGED;
            if (io.iters_out) io.iters_out[s] = 0;
            if (io.relres_out) io.relres_out[s] = fabs(bw - yy) / yy;
        }
        cp_async_wait<0>();
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
        L.qacc = take((size_t)n_obs * 16);   // observables | residual of the adjoint right-hand side
        L.cvec = take((size_t)ncv_smem * 8);
        L.cslot = take((size_t)cmax * 4);
        L.ring = take((size_t)ring_bytes);
        L.total = o;
        return L;
    }
};

template <int MODE, int NU>   // NU: the column fits 32 * NU entries
__global__ void __launch_bounds__(NU >= 8 ? 512 : (NU >= 4 ? 768 : 1024), 1) frontal_cta_kernel(FrontalDev P, FrontalIO io, FrontalCtaSmem L) {
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
        if (loader) cp_async_commit();
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
        double* res = qacc + io.n_obs;
        const bool adj = MODE == FRONTAL_MODE_SOLVE && io.adj != 0;
        if (adj)
            for (int o = tid; o < io.n_obs; o += NT)
                res[o] = io.adj == 2 ? (o == io.unit_row ? 1.0 : 0.0)
                                     : io.qoi_in[(size_t)s * io.n_obs + o] - io.data[(size_t)s * io.data_stride + o];
        if (loader) cp_async_wait<0>();
        __syncthreads();
        if (adj && io.cost_out && tid == 0) {
            double cost = 0.0;
            for (int o = 0; o < io.n_obs; ++o) cost = fma(res[o], res[o], cost);
            io.cost_out[s] = 0.5 * cost;
        }

        bool bad = false;
        double myq = 0.0;   // QOI mode: thread rtid = 1 + o accumulates observable o; rtid = 0 accumulates y.y
        size_t cp = 0;
        for (int j = -1; j < n; ++j) {
            // ---- G_j: pivot, scaled column (zeroing the consumed entries), pivot row of the right-hand sides,
            //      and the assembly of column j + 1 (disjoint entries)
            ring.fill(lane, loader);
            if (loader) cp_async_commit();
            const unsigned char* rec = ring.buf + (ring.rd & ring.mask);   // records never straddle the wrap point
            const uint4 h0 = *reinterpret_cast<const uint4*>(rec);         // c | pivot slot | npos | nent
            const uint4 h1 = *reinterpret_cast<const uint4*>(rec + 16);    // nobs | bytes | rhs
            const unsigned c = h0.x, p = h0.y, npos = h0.z, nent = h0.w, nobs = h1.x, reclen = h1.y;
            const unsigned short* slots = reinterpret_cast<const unsigned short*>(rec + 32);
            const unsigned* posv = reinterpret_cast<const unsigned*>(rec + 32 + ((2 * c + 7) & ~7u));
            const double* coefv = reinterpret_cast<const double*>(posv + 2 * npos);
            const unsigned* termv = reinterpret_cast<const unsigned*>(coefv + nent);
            const double* ovalv = reinterpret_cast<const double*>(reinterpret_cast<const unsigned char*>(termv) + ((4 * nent + 7) & ~7u));
            const unsigned* orowv = reinterpret_cast<const unsigned*>(ovalv + nobs);
            const unsigned pd = tri_u(p) + p;
            double rinv = 0.0;
            // only the threads that scale something need 1/sqrt(pivot): the gatherers, the right-hand-side threads, and
            // thread 0 (which keeps the breakdown flag)
            if (j >= 0 && (tid < (int)c || rtid < R || tid == 0)) {
                const double dd = F[pd];
                bad |= !(dd > 0.0);
                rinv = rsqrt(dd);
                for (int a = tid; a < (int)c; a += NT) {
                    const unsigned sa = slots[a];
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
                        if (adj) {   // -B_obs^T res at this pivot's dof
                            for (unsigned o = 0; o < nobs; ++o) v = fma(-ovalv[o], res[orowv[o]], v);
                        } else {
                            v += __longlong_as_double(((long long)h1.w << 32) | h1.z);
                        }
                    } else {
                        for (unsigned o = 0; o < nobs; ++o)
                            if ((int)orowv[o] == rtid - 1) v += ovalv[o];
                    }
                    ypiv[rtid] = v * rinv;
                }
            }
            // assembly of column j + 1 by ONE warp: lane e forms the product of entry e, lane q sums the entries of
            // position q with shuffles (the serial per-position loop was the critical path of this phase)
            if (warp == (nw >= 2 ? nw - 2 : 0)) {   // not the warp that forms the pivot row of the right-hand sides
                const unsigned rl = 31u - (unsigned)lane;
                if (nent <= 32u && npos <= 32u) {
                    const double prod = rl < nent ? coefv[rl] * cv[termv[rl]] : 0.0;
                    const unsigned pw = rl < npos ? posv[2 * rl + 1] : 0u;
                    const unsigned cnt = pw & 255u, e0 = pw >> 8;
                    unsigned cmax_w = cnt;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) cmax_w = max(cmax_w, __shfl_xor_sync(0xffffffffu, cmax_w, o));
                    double sum = 0.0;
                    for (unsigned k = 0; k < cmax_w; ++k) {
                        const double pk = __shfl_sync(0xffffffffu, prod, 31 - (int)min(e0 + k, 31u));
                        if (k < cnt) sum += pk;
                    }
                    if (rl < npos) F[posv[2 * rl]] += sum;
                } else {
                    for (unsigned q = rl; q < npos; q += 32) {
                        const unsigned pw = posv[2 * q + 1], cnt = pw & 255u;
                        unsigned e = pw >> 8;
                        double sum = 0.0;
                        for (unsigned k = 0; k < cnt; ++k, ++e) sum = fma(coefv[e], cv[termv[e]], sum);
                        F[posv[2 * q]] += sum;
                    }
                }
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
                // row a of the triangle (b = 0..a) goes to warp a mod nw; lane holds the column entries b = lane + 32 u in
                // registers, so an update costs one shared load, one FMA and one store.  Two rows are in flight at a time
                // (all loads before the stores) to overlap the shared-memory latency.
                double lb[NU];
                unsigned sb[NU];
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    if (32 * u >= (int)c) break;
                    const int b = lane + 32 * u;
                    lb[u] = b < (int)c ? lcol[b] : 0.0;
                    sb[u] = b < (int)c ? cslot[b] : 0u;
                }
                for (int a = warp; a < (int)c; a += 2 * nw) {
                    const int a2 = a + nw;
                    const bool two = a2 < (int)c;
                    const double la = lcol[a], la2 = two ? lcol[a2] : 0.0;
                    const unsigned tr = tri_u(cslot[a]), tr2 = two ? tri_u(cslot[a2]) : 0u;
                    double f[NU], f2[NU];
                    const int amax = two ? a2 : a;   // rows ascend: a2 > a
#pragma unroll
                    for (int u = 0; u < NU; ++u) {
                        if (32 * u > amax) break;
                        if (lane + 32 * u <= a) f[u] = F[tr + sb[u]];
                        if (two && lane + 32 * u <= a2) f2[u] = F[tr2 + sb[u]];
                    }
#pragma unroll
                    for (int u = 0; u < NU; ++u) {
                        if (32 * u > amax) break;
                        if (lane + 32 * u <= a) F[tr + sb[u]] = fma(-la, lb[u], f[u]);
                        if (two && lane + 32 * u <= a2) F[tr2 + sb[u]] = fma(-la2, lb[u], f2[u]);
                    }
                }
            }
            if (loader) cp_async_wait<1>();   // chunks requested one step ago must have landed; this step's may still fly
            __syncthreads();
            cp += (j >= 0 ? c : 0u);
            ring.rd += reclen;
        }
        if (loader) cp_async_wait<0>();   // nothing may land in the ring after it is re-armed
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
        if (loader) {
            cp_async_commit();
            cp_async_wait<0>();
        }
        __syncthreads();
        ring.rd += ring.u32(16);   // skip the prologue record (D1's prefetch schedule)
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
            if (loader) cp_async_commit();
            const unsigned char* rec = ring.buf + (ring.rd & ring.mask);
            const uint4 h0 = *reinterpret_cast<const uint4*>(rec);         // c | pivot slot | nobs | dof
            const uint4 h1 = *reinterpret_cast<const uint4*>(rec + 16);    // bytes | next c | rhs
            const unsigned c = h0.x, p = h0.y, nobs = h0.z, dof = h0.w, reclen = h1.x, cnext = h1.y;
            const unsigned short* slots = reinterpret_cast<const unsigned short*>(rec + 40);
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
                if (tid + u * NT < (int)c) part = fma(lpre[u], yv[slots[tid + u * NT]], part);
            for (int a = tid + LPF * NT; a < (int)c; a += NT) part = fma(Lw[cpj + a], yv[slots[a]], part);
            part = warp_sum(part);
            double* rd = red + (j & 1) * 32;
            if (lane == 0) rd[warp] = part;
            __syncthreads();
            double acc = 0.0;
            for (int w = 0; w < nw; ++w) acc += rd[w];
            const double wj = (ycur - acc) * rcur;
            if (tid == 0) {
                yv[p] = wj;
                bw = fma(__longlong_as_double(((long long)h1.w << 32) | h1.z), wj, bw);
                if (io.w_out) io.w_out[(size_t)s * n + dof] = wj;
                const double* oval = reinterpret_cast<const double*>(rec + 40 + ((2 * c + 7) & ~7u));
                const unsigned* orow = reinterpret_cast<const unsigned*>(oval + nobs);
                for (unsigned o = 0; o < nobs; ++o) qacc[orow[o]] += oval[o] * wj;
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
            if (io.qoi_out && !adj)
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
                                        double* __restrict__ cv) {   // lane_major: 0 = [sample][ncv], else samples per group
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
    if (lane_major) {   // x = sample, y = cell; groups of `lane_major` samples: [group][ncv][lane]
        const long long s = g * 32 + tx;
        const size_t grp = (size_t)(s / lane_major), ln = (size_t)(s % lane_major);
        for (int cy = ty; cy < 32; cy += 8) {
            const int c = c0 + cy;
            if (c < n_cells) cv[(grp * ncv + 1 + c) * lane_major + ln] = tile[tx][cy];
        }
        if (blockIdx.x == 0 && ty == 0) cv[grp * ncv * lane_major + ln] = 1.0;
    } else {            // x = cell, y = sample
        for (int sy = ty; sy < 32; sy += 8) {
            const long long s = g * 32 + sy;
            const int c = c0 + tx;
            if (s < N && c < n_cells) cv[(size_t)s * ncv + 1 + c] = tile[sy][tx];
            if (s < N && blockIdx.x == 0 && tx == 0) cv[(size_t)s * ncv] = 1.0;
        }
    }
}

// ------------------------------------------------------------------------------------------------ gradient form
// g[s][i] = assemble(k_hat_i * c'(k) * inner(grad w, grad v) * dx) (fom/forward_solve.py:313-314; exp(k): forward_solve_exp.py:299)
//         = sum over the cells e around vertex i of  weight(e, i) * (w_e^T K_e v_e),
// weight = 1/3 for the plain parametrisation (k_hat is a hat function, the rest is constant per cell), the k_hat-weighted
// degree-4 quadrature of exp(k) otherwise (vertex_weight_exp, pcg_small.cuh).  One thread per (sample, vertex).
__global__ void __launch_bounds__(256) frontal_gradform_kernel(const double* __restrict__ w, const double* __restrict__ v,
                                                               const double* __restrict__ k, long long N, int n,
                                                               const int* __restrict__ dptr, const int* __restrict__ dcell,
                                                               const int* __restrict__ cells, const double* __restrict__ Ke,
                                                               int coef_mode, double* __restrict__ g, long long g_stride) {
    const long long total = N * n;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long s = idx / n;
        const int i = (int)(idx - s * n);
        const double* ws = w + s * n;
        const double* vs = v + s * n;
        const double* ks = k + s * n;
        double acc = 0.0;
        for (int j = dptr[i]; j < dptr[i + 1]; ++j) {
            const int e = dcell[j];
            const int ca = cells[3 * e], cb = cells[3 * e + 1], cc = cells[3 * e + 2];
            const double* K = Ke + 9 * (size_t)e;
            const double va = vs[ca], vb = vs[cb], vc = vs[cc];
            const double t0 = fma(K[0], va, fma(K[1], vb, K[2] * vc));
            const double t1 = fma(K[3], va, fma(K[4], vb, K[5] * vc));
            const double t2 = fma(K[6], va, fma(K[7], vb, K[8] * vc));
            const double se = fma(ws[ca], t0, fma(ws[cb], t1, ws[cc] * t2));
            if (coef_mode == 0) {
                acc = fma(se, 1.0 / 3.0, acc);
            } else {
                const int ob = ca == i ? cb : ca, oc = cc == i ? cb : cc;   // the two other vertices
                acc = fma(se, vertex_weight_exp(ks[i], ks[ob], ks[oc]), acc);
            }
        }
        g[s * g_stride + i] = acc;   // g_stride = n (gradient) or n_obs n (one row of the Jacobian)
    }
}

}  // namespace tfin
