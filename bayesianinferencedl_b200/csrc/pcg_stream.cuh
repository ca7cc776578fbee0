// K4: batched Jacobi-PCG for meshes too large for one SM (refined meshes, n ~ 1e5): a genuinely HBM-streaming
// kernel.  Included by tfin_api.cu only.
//
//   * a TILE of S samples is solved by ONE persistent CTA (512 threads, 128 registers); tiles are pulled from a
//     global counter, so there is no grid-wide synchronisation at all: every reduction is CTA-local
//   * the CG vectors x, z, p, q^, w^ of a tile live in HBM as S/2 PLANES of double2 (two samples per element,
//     plane stride ldr): every vector access of a warp is a contiguous run of 16-byte elements
//   * the operator is applied MATRIX-FREE from shared term-tagged entries (col | term << 24, coef), stored slot-major
//     (ELL) and sorted by term inside a row:
//         (A(theta) v)_i = sum_e theta[term_e] * coef_e * v[col_e]          (averaged_affine_ROM.py:156-162)
//     so no per-sample matrix values exist anywhere.  A lane owns ONE row and FOUR samples (two planes); 32/(S/4)
//     consecutive rows per warp make the operator loads coalesced (one wavefront per slot), the per-sample
//     conductivities of the current term are cached in registers and reloaded only when the term changes, and the
//     operator slice of the warp's next row group is staged in shared memory by cp.async while the current one runs
//   * single-reduction (Chronopoulos-Gear) PCG written in terms of z = D^-1 r, so that the diagonal D is needed
//     ONLY inside the SpMV pass, where it falls out of the entries with col == row:
//         pass A   w^ = D^-1 A z,  gamma = z.D z (= r.z),  delta = z.A z        read z (gather), write w^     16 n B
//         scalars  beta = gamma/gamma_old,  alpha = gamma / (delta - beta gamma / alpha_old)
//         pass B   p = z + beta p,  q^ = w^ + beta q^,  x += alpha p,  z -= alpha q^    read 5, write 4       72 n B
//     => the 88 n bytes per iteration and sample of the SURVEY 8d model, in two passes and one reduction; pass B is a
//     pure streaming pass (no operator, no diagonal)
//   * rows are renumbered on the host by reverse Cuthill-McKee (tfin_api.cu), so every gather of a row sweep falls in
//     a window of +-B rows around the row.  For S <= 8 that window of z is kept in a SHARED-MEMORY RING filled by TMA
//     bulk copies (cp.async.bulk + mbarrier, one 4 KB copy per plane and 256-row chunk) running LA chunks ahead of
//     the sweep: HBM is read exactly once, strictly sequentially, and every gather is an LDS.  Warps are decoupled:
//     "full" mbarriers (transaction count) gate the consumers, a "done" mbarrier per step (one arrival per warp)
//     gates the reuse of a ring slot by the producer thread.  Wider tiles / wider bandwidths use direct gathers
//   * converged samples of a tile are frozen (alpha = beta = 0) until the whole tile is done
#pragma once

#include "pcg_small.cuh"

namespace tfin {

struct StreamOp {
    int n, n_terms;
    int ldr;                     // padded row count (multiple of 32) = plane stride of the vectors
    int We;                      // ELL width, multiple of STREAM_KC
    const uint32_t* colterm;     // [We][ldr]  col | term << 24;  padding: col = min(row, n-1), coef = 0
    const double* coef;          // [We][ldr]
    const unsigned char* cnt;    // [ldr]      max entries in use over the aligned 32-row block of the row (>= 1)
    const double* rhs;           // [n]        (renumbered)
    const int* perm;             // [n]        renumbered row -> caller's dof index (for w_out)
    int hb, la;                  // ring mode: halo in chunks (ceil(B / chunk rows)), TMA look-ahead in chunks
    int Ws;                      // operator slots staged through shared memory per row group
};

#define STREAM_RING_BYTES (128 * 1024)  // shared-memory ring of the gathered vector (ring mode)
#define STREAM_RING_SLOTS 8             // chunks in the ring

// streaming 16-byte load of read-write data that is touched once per pass: do not allocate an L1 line
__device__ __forceinline__ double2 ld_stream2(const double2* p) {
    double2 v;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: a barrier that does not complete within ~1 s is a protocol bug -> report and abort instead of hanging
__device__ __forceinline__ void mbar_wait_checked(uint64_t* bar, uint32_t parity, int what, int step) {
    const uint32_t addr = smem_u32(bar);
    const long long t0 = clock64();
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if (clock64() - t0 > 2000000000LL) {
            printf("tfin stream ring: barrier timeout what=%d step=%d parity=%u block=%d thread=%d raw=%llx\n", what, step, parity,
                   (int)blockIdx.x, (int)threadIdx.x, *(volatile unsigned long long*)bar);
            __trap();
        }
    }
}
__device__ __forceinline__ void mbar_inval(uint64_t* bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

#define STREAM_KC 4  // the ELL width is padded to a multiple of this (slot loop unroll)

// 1/d to full double precision from the hardware seed (MUFU.RCP64H) and two Newton steps; d > 0
__device__ __forceinline__ double fast_rcp(double d) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    x = x * (2.0 - d * x);
    x = x * (2.0 - d * x);
    return x;
}

enum { SW_DIAG = 0, SW_DIRECT = 1, SW_RING = 2 };
template <int M>
struct SweepMode { static constexpr int value = M; };

template <int S, bool RING>
__global__ void __launch_bounds__(512, 1) pcg_stream_kernel(StreamOp op, CsrRows obs, PcgIO io,
                                                            double* __restrict__ work,
                                                            unsigned long long* prof /* nullable: CTA 0 clocks */) {
    static_assert(S == 4 || S == 8 || S == 16 || S == 32, "tile width");
    constexpr int NPL = S / 2;     // planes (double2 = two samples)
    constexpr int HL = S / 4;      // lanes per row (a lane owns 4 samples = 2 planes)
    constexpr int RPW = 32 / HL;   // rows per warp step
    __shared__ double s_theta[TFIN_MAX_TERMS * S];
    __shared__ double s_red[16][2][S];  // per-warp partial sums
    __shared__ double s_tot[2][S];
    __shared__ double s_alpha[S], s_beta[S], s_gamma[S], s_aold[S], s_thresh[S];
    __shared__ int s_active[S], s_iters[S], s_status[S], s_nactive;
    __shared__ long long s_tile;
    __shared__ uint64_t s_full[STREAM_RING_SLOTS], s_done[8];  // warps drift by <= la + 1 <= 5 steps: 8 "done" phases in flight suffice
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    constexpr int CH = 16 * RPW;                                   // rows per CTA step = rows per ring chunk
    constexpr int RING_ROWS = STREAM_RING_BYTES / (NPL * 16);      // rows per plane in the ring (power of two)
    constexpr int NR = RING_ROWS / CH;
    constexpr uint32_t RMASK = RING_ROWS - 1;
    static_assert(!RING || NR == STREAM_RING_SLOTS, "ring geometry");
    double2* ring = reinterpret_cast<double2*>(dyn_smem);
    using GatherMode = SweepMode<RING ? SW_RING : SW_DIRECT>;
    // The barriers are initialised ONCE; their phases run on across sweeps.  Slot s is used by chunks (steps)
    // s, s + 8, ... of every sweep, i.e. uses(s) = ceil((n_chunks - s) / 8) times per sweep, so the phase parity of
    // the k-th use inside ring sweep number q is (q * uses(s) + k) & 1  (ring_q is uniform over the CTA).
    if (RING && threadIdx.x == 0) {
        for (int b = 0; b < STREAM_RING_SLOTS; ++b) mbar_init(&s_full[b], 1);
        for (int b = 0; b < 8; ++b) mbar_init(&s_done[b], blockDim.x >> 5);
        mbar_fence_init();
    }
    unsigned ring_q = 0;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int n = op.n, ldr = op.ldr;
    const int hl = lane / RPW, rw = lane % RPW;  // which 4 samples of the row; row within the warp's row group
    const size_t vec2 = (size_t)NPL * ldr;        // double2 elements per vector
    double2* x2 = reinterpret_cast<double2*>(work) + (size_t)blockIdx.x * 5 * vec2;
    double2* z2 = x2 + vec2;
    double2* p2 = z2 + vec2;
    double2* q2 = p2 + vec2;
    double2* w2 = q2 + vec2;
    const double2* th2 = reinterpret_cast<const double2*>(s_theta);
    const long long n_tiles = (io.N + S - 1) / S;
    const int n_blk = ldr / RPW;
    const double2 zero2 = make_double2(0.0, 0.0);

    // per-lane partials of 2 quantities x 4 samples -> per-sample CTA totals in s_tot[which][S]
    auto reduce8 = [&](double (&a)[4], double (&b)[4]) {
#pragma unroll
        for (int o = 1; o < RPW; o <<= 1) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                a[c] += __shfl_xor_sync(0xffffffffu, a[c], o);
                b[c] += __shfl_xor_sync(0xffffffffu, b[c], o);
            }
        }
        if (rw == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                s_red[warp][0][4 * hl + c] = a[c];
                s_red[warp][1][4 * hl + c] = b[c];
            }
        }
        __syncthreads();
        if (tid < 2 * S) {
            const int which = tid / S, s = tid % S;
            double t = 0.0;
            for (int w = 0; w < nwarps; ++w) t += s_red[w][which][s];
            s_tot[which][s] = t;
        }
        __syncthreads();
    };

    // Row sweep.  For every row i of this lane and its 4 samples:  av = (A v)_i  and  own = v_i  (gather modes only),
    // dd = diag_i;  row_fn(i, valid, av, dd, own).
    // The operator slice of the warp's NEXT row group (its first Ws slots: RPW packed columns + RPW coefficients per
    // slot, contiguous in the slot-major arrays) is copied into a per-warp double buffer in shared memory with
    // cp.async while the current group is computed, so no L2 round trip sits in the dependent chain
    // entry -> gather -> FMA; slots beyond Ws (very wide rows) are read from global memory directly.
    const int n_chunks = ldr / CH;  // ring mode: ldr is a multiple of CH, every warp has exactly n_chunks row groups
    const int Ws = op.Ws;
    const int stage_bytes = Ws * RPW * 12;
    unsigned char* op_stage = dyn_smem + (RING ? STREAM_RING_BYTES : 0) + (size_t)warp * 2 * stage_bytes;
    auto stage_issue = [&](int blk, int st, int c) {
        constexpr int PPS = 3 * RPW / 4;  // 16-byte pieces per slot: RPW/4 of columns, RPW/2 of coefficients
        const int pieces = min(c, Ws) * PPS;
        unsigned char* dst = op_stage + (size_t)st * stage_bytes;
        for (int pz = lane; pz < pieces; pz += 32) {
            const int k = pz / PPS, r = pz % PPS;
            if (r < RPW / 4)
                cp_async16(dst + ((size_t)k * RPW + r * 4) * 4, op.colterm + (size_t)k * ldr + (size_t)blk * RPW + r * 4);
            else
                cp_async16(dst + (size_t)Ws * RPW * 4 + ((size_t)k * RPW + (r - RPW / 4) * 2) * 8,
                           op.coef + (size_t)k * ldr + (size_t)blk * RPW + (r - RPW / 4) * 2);
        }
    };
    auto ring_issue = [&](const double2* v, int c) {  // one thread
        uint64_t* bar = &s_full[c & (NR - 1)];
        mbar_expect_tx(bar, (uint32_t)(NPL * CH * 16));
        const uint32_t slot_row = ((uint32_t)c * CH) & RMASK;
#pragma unroll
        for (int pl = 0; pl < NPL; ++pl)
            tma_bulk_g2s(ring + (size_t)pl * RING_ROWS + slot_row, v + (size_t)pl * ldr + (size_t)c * CH, CH * 16, bar);
    };
    auto sweep = [&](auto mode, const double2* __restrict__ v, auto&& row_fn) {
        constexpr int MODE = decltype(mode)::value;
        constexpr bool GATHER = MODE != SW_DIAG;
        const double2* va = v + (size_t)(2 * hl) * ldr;               // this lane's first plane
        const double2* ra = ring + (size_t)(2 * hl) * RING_ROWS;      // ... and its image in the ring
        if (MODE == SW_RING) {
            __syncthreads();  // nobody is still using the ring; all writes of v are done
            if (tid == 0) {
                asm volatile("fence.proxy.async;" ::: "memory");
                for (int c = 0; c < op.hb + op.la && c < n_chunks; ++c) ring_issue(v, c);
            }
        }
        const unsigned q = ring_q;
        auto parity_of = [&](int c) -> uint32_t {  // c = chunk or step index inside this sweep
            const unsigned uses = (unsigned)(n_chunks - (c & 7) + 7) >> 3;
            return (q * uses + ((unsigned)c >> 3)) & 1u;
        };
        int blk = warp, st = 0;
        int cnt = blk < n_blk ? (int)op.cnt[blk * RPW] : 0;
        int cnt_n = blk + nwarps < n_blk ? (int)op.cnt[(blk + nwarps) * RPW] : 0;
        if (blk < n_blk) stage_issue(blk, 0, cnt);
        cp_async_commit();
        for (int t = 0; blk < n_blk; ++t, st ^= 1) {
            const int nblk = blk + nwarps, nnblk = nblk + nwarps;
            const int cnt_nn = nnblk < n_blk ? (int)op.cnt[nnblk * RPW] : 0;  // consumed two row groups from now
            if (nblk < n_blk) stage_issue(nblk, st ^ 1, cnt_n);
            cp_async_commit();
            const int i = blk * RPW + rw;
            if (MODE == SW_RING) {
                if (tid == 0) {  // producer: chunk t + hb + la reuses the slot whose chunk was last needed in step t - 2
                    const int cn = t + op.hb + op.la;
                    if (cn < n_chunks) {
                        if (t >= 2) mbar_wait_checked(&s_done[(t - 2) & 7], parity_of(t - 2), 0, t);
                        ring_issue(v, cn);
                    }
                }
                if (t == 0) {
                    for (int c = 0; c <= op.hb && c < n_chunks; ++c) mbar_wait_checked(&s_full[c & (NR - 1)], parity_of(c), 1, c);
                } else {
                    const int c = t + op.hb;
                    if (c < n_chunks) mbar_wait_checked(&s_full[c & (NR - 1)], parity_of(c), 2, t);
                }
            }
            double2 own[2] = {zero2, zero2};
            if (MODE == SW_DIRECT) {
                own[0] = va[i];
                own[1] = va[(size_t)ldr + i];
            }
            cp_async_wait<1>();
            __syncwarp();
            const uint32_t* sct = reinterpret_cast<const uint32_t*>(op_stage + (size_t)st * stage_bytes) + rw;
            const double* scf = reinterpret_cast<const double*>(op_stage + (size_t)st * stage_bytes + (size_t)Ws * RPW * 4) + rw;
            double2 av[2] = {zero2, zero2}, dd[2] = {zero2, zero2}, th[2] = {zero2, zero2};
            uint32_t cur = 0xffffffffu;
            auto consume = [&](uint32_t ct, double cf) {
                const uint32_t term = ct >> 24, col = ct & 0xffffffu;
                double2 pa = zero2, pb = zero2;
                if (MODE == SW_DIRECT) {
                    pa = va[col];
                    pb = va[(size_t)ldr + col];
                } else if (MODE == SW_RING) {
                    pa = ra[col & RMASK];
                    pb = ra[RING_ROWS + (col & RMASK)];
                }
                if (term != cur) {
                    th[0] = th2[term * NPL + 2 * hl];
                    th[1] = th2[term * NPL + 2 * hl + 1];
                    cur = term;
                }
                const double c0 = th[0].x * cf, c1 = th[0].y * cf, c2 = th[1].x * cf, c3 = th[1].y * cf;
                if (GATHER) {
                    av[0].x = fma(c0, pa.x, av[0].x);
                    av[0].y = fma(c1, pa.y, av[0].y);
                    av[1].x = fma(c2, pb.x, av[1].x);
                    av[1].y = fma(c3, pb.y, av[1].y);
                }
                if (col == (uint32_t)i) {
                    dd[0].x += c0;
                    dd[0].y += c1;
                    dd[1].x += c2;
                    dd[1].y += c3;
                }
            };
            const int cs = min(cnt, Ws);
#pragma unroll 4
            for (int k = 0; k < cs; ++k) consume(sct[k * RPW], scf[k * RPW]);
            for (int k = cs; k < cnt; ++k)  // rows wider than the staged slots (rare)
                consume(__ldg(op.colterm + (size_t)k * ldr + i), __ldg(op.coef + (size_t)k * ldr + i));
            if (MODE == SW_RING) {
                own[0] = ra[(uint32_t)i & RMASK];
                own[1] = ra[RING_ROWS + ((uint32_t)i & RMASK)];
            }
            row_fn(i, i < n, av, dd, own);
            __syncwarp();  // everyone is done with this operator stage before the next cp.async overwrites it
            if (MODE == SW_RING && lane == 0) mbar_arrive(&s_done[t & 7]);
            blk = nblk;
            cnt = cnt_n;
            cnt_n = cnt_nn;
        }
        cp_async_wait<0>();
        if (MODE == SW_RING) ++ring_q;
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
            s_alpha[tid] = s_beta[tid] = 0.0;
            s_gamma[tid] = s_aold[tid] = 1.0;
        }
        __syncthreads();

        // ---- init: x = 0, p = 0, q^ = 0, z = D^-1 b   (w^ is written by pass A before it is read)
        sweep(SweepMode<SW_DIAG>{}, x2, [&](int i, bool valid, const double2 (&)[2], const double2 (&dd)[2], const double2 (&)[2]) {
            const double b = valid ? op.rhs[i] : 0.0;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const size_t o = (size_t)(2 * hl + j) * ldr + i;
                x2[o] = zero2;
                p2[o] = zero2;
                q2[o] = zero2;
                z2[o] = valid ? make_double2(b * fast_rcp(dd[j].x), b * fast_rcp(dd[j].y)) : zero2;
            }
        });
        if (RING) {
            __threadfence();
            asm volatile("fence.proxy.async;" ::: "memory");
        }
        __syncthreads();

        const bool timing = prof != nullptr && blockIdx.x == 0 && tid == 0;
        long long t_mark = timing ? clock64() : 0;
        auto lap = [&](int slot) {
            if (timing) {
                const long long t = clock64();
                prof[slot] += (unsigned long long)(t - t_mark);
                t_mark = t;
            }
        };
        for (int it = 1;; ++it) {
            // ---------------- pass A: w^ = D^-1 A z, gamma = z.D z, delta = z.A z
            double ga[4] = {0.0, 0.0, 0.0, 0.0}, de[4] = {0.0, 0.0, 0.0, 0.0};
            sweep(GatherMode{}, z2, [&](int i, bool valid, const double2 (&av)[2], const double2 (&dd)[2], const double2 (&own)[2]) {
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const size_t o = (size_t)(2 * hl + j) * ldr + i;
                        const double2 zv = own[j];
                        w2[o] = make_double2(av[j].x * fast_rcp(dd[j].x), av[j].y * fast_rcp(dd[j].y));
                        ga[2 * j] = fma(dd[j].x * zv.x, zv.x, ga[2 * j]);
                        ga[2 * j + 1] = fma(dd[j].y * zv.y, zv.y, ga[2 * j + 1]);
                        de[2 * j] = fma(zv.x, av[j].x, de[2 * j]);
                        de[2 * j + 1] = fma(zv.y, av[j].y, de[2 * j + 1]);
                    }
                }
            });
            reduce8(ga, de);
            lap(0);
            if (tid < S) {
                const double g = s_tot[0][tid], d = s_tot[1][tid];
                double a = 0.0, b = 0.0;
                if (s_active[tid]) {
                    if (it == 1) s_thresh[tid] = io.tol2 * g;
                    bool stop = true;
                    if (g <= s_thresh[tid]) s_status[tid] = TFIN_STATUS_CONVERGED;  // includes b == 0
                    else if (!(g == g) || !(d > 0.0)) s_status[tid] = TFIN_STATUS_BREAKDOWN;
                    else if (it > io.maxit) s_status[tid] = TFIN_STATUS_MAXIT;
                    else {
                        b = it == 1 ? 0.0 : g / s_gamma[tid];
                        const double denom = it == 1 ? d : d - b * g / s_aold[tid];
                        if (denom > 0.0) {
                            a = g / denom;
                            s_gamma[tid] = g;
                            s_aold[tid] = a;
                            stop = false;
                        } else {
                            b = 0.0;
                            s_status[tid] = TFIN_STATUS_BREAKDOWN;
                        }
                    }
                    if (stop) {
                        s_active[tid] = 0;
                        s_iters[tid] = it - 1;
                    }
                }
                s_alpha[tid] = a;
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
            // ---------------- pass B: p = z + beta p, q^ = w^ + beta q^, x += alpha p, z -= alpha q^   (pure streaming)
            for (int pl = 0; pl < NPL; ++pl) {
                const double2 al = make_double2(s_alpha[2 * pl], s_alpha[2 * pl + 1]);
                const double2 be = make_double2(s_beta[2 * pl], s_beta[2 * pl + 1]);
                const size_t base = (size_t)pl * ldr;
                for (int r0 = tid; r0 < n; r0 += 2 * blockDim.x) {  // two elements in flight per thread
                    const int r1 = r0 + blockDim.x;
                    const bool v1 = r1 < n;
                    const size_t o0 = base + r0, o1 = base + (v1 ? r1 : r0);
                    const double2 za = ld_stream2(z2 + o0), pa = ld_stream2(p2 + o0), wa = ld_stream2(w2 + o0),
                                  qa = ld_stream2(q2 + o0), xa = ld_stream2(x2 + o0);
                    const double2 zb = ld_stream2(z2 + o1), pb = ld_stream2(p2 + o1), wb = ld_stream2(w2 + o1),
                                  qb = ld_stream2(q2 + o1), xb = ld_stream2(x2 + o1);
                    {
                        const double2 pn = make_double2(fma(be.x, pa.x, za.x), fma(be.y, pa.y, za.y));
                        const double2 qn = make_double2(fma(be.x, qa.x, wa.x), fma(be.y, qa.y, wa.y));
                        p2[o0] = pn;
                        q2[o0] = qn;
                        x2[o0] = make_double2(fma(al.x, pn.x, xa.x), fma(al.y, pn.y, xa.y));
                        z2[o0] = make_double2(fma(-al.x, qn.x, za.x), fma(-al.y, qn.y, za.y));
                    }
                    if (v1) {
                        const double2 pn = make_double2(fma(be.x, pb.x, zb.x), fma(be.y, pb.y, zb.y));
                        const double2 qn = make_double2(fma(be.x, qb.x, wb.x), fma(be.y, qb.y, wb.y));
                        p2[o1] = pn;
                        q2[o1] = qn;
                        x2[o1] = make_double2(fma(al.x, pn.x, xb.x), fma(al.y, pn.y, xb.y));
                        z2[o1] = make_double2(fma(-al.x, qn.x, zb.x), fma(-al.y, qn.y, zb.y));
                    }
                }
            }
            if (RING) asm volatile("fence.proxy.async;" ::: "memory");  // z: generic-proxy writes, read next by TMA
            __syncthreads();
            lap(1);
            if (timing) prof[3] += 1;
        }

        // ---------------- epilogue: true residual ||b - A x|| / ||b||, observables, optional w
        if (io.relres_out || io.status_out) {
            double rr[4] = {0.0, 0.0, 0.0, 0.0}, bb[4] = {0.0, 0.0, 0.0, 0.0};
            sweep(GatherMode{}, x2, [&](int i, bool valid, const double2 (&av)[2], const double2 (&)[2], const double2 (&)[2]) {
                if (valid) {
                    const double b = op.rhs[i];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        rr[2 * j] = fma(b - av[j].x, b - av[j].x, rr[2 * j]);
                        rr[2 * j + 1] = fma(b - av[j].y, b - av[j].y, rr[2 * j + 1]);
                    }
                    bb[0] = fma(b, b, bb[0]);
                }
            });
            bb[1] = bb[2] = bb[3] = bb[0];
            reduce8(rr, bb);
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
                double acc[4] = {0.0, 0.0, 0.0, 0.0}, dummy[4] = {0.0, 0.0, 0.0, 0.0};
                const int j0 = obs.ptr[o], j1 = obs.ptr[o + 1];
                for (int j = j0 + warp * RPW + rw; j < j1; j += nwarps * RPW) {
                    const double w = obs.val[j];
                    const size_t e = (size_t)(2 * hl) * ldr + obs.idx[j];
                    const double2 xa = x2[e], xb = x2[e + ldr];
                    acc[0] = fma(w, xa.x, acc[0]);
                    acc[1] = fma(w, xa.y, acc[1]);
                    acc[2] = fma(w, xb.x, acc[2]);
                    acc[3] = fma(w, xb.y, acc[3]);
                }
                reduce8(acc, dummy);
                if (tid < S && s_base + tid < io.N) io.qoi_out[(s_base + tid) * obs.rows + o] = s_tot[0][tid];
                __syncthreads();
            }
        }
        if (io.w_out) {
            for (int pl = 0; pl < NPL; ++pl) {
                const long long sa = s_base + 2 * pl, sb = sa + 1;
                for (int i = tid; i < n; i += blockDim.x) {
                    const double2 xv = x2[(size_t)pl * ldr + i];
                    const int dof = op.perm[i];
                    if (sa < io.N) io.w_out[sa * (long long)n + dof] = xv.x;
                    if (sb < io.N) io.w_out[sb * (long long)n + dof] = xv.y;
                }
            }
        }
    }
}

}  // namespace tfin
