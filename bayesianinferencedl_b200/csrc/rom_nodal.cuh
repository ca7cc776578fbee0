// R4: batched nodal-conductivity LSPG reduction (Fin.r_fwd_no_full -> Fin.reduced_forward, fom/forward_solve.py:421-464):
//     psi = A(k) phi,   A_r = psi^T (A phi) = psi^T psi,   B_r = psi^T b        (per sample, A(k) assembled in-kernel)
// written into the augmented packed layout of rom.cuh so that rom_chol_kernel finishes the solve and the projections.
//
// One CTA per sample (persistent over the chunk).  The n rows are swept in tiles of 16: phase A forms the tile of psi
// in shared memory (per-row operator entries from the nodal ELL of K2, phi rows through L1/L2), phase B accumulates
// the Gram matrix in registers, 6 x 6 entries per thread, lower-triangular tiles only (TT (TT+1) / 2 threads,
// TT = ceil((n_r + 1) / 6)).  Column n_r of the psi tile carries the right-hand side b, so B_r = psi^T b is row n_r of
// the same Gram matrix -- exactly the augmented row the Cholesky kernel expects.  psi tiles and operator rows are
// double buffered: one barrier per tile.
#pragma once

#include "pcg_small.cuh"
#include "rom.cuh"

namespace tfin {

constexpr int RN_TR = 16;  // rows of psi per tile

struct RomNodalSmem {
    size_t kbar_off, psi_off, a_off, col_off, total;
    int slots;  // W + 1 (diagonal first)
    __host__ __device__ static RomNodalSmem make(int n_cells, int nrp, int W) {
        RomNodalSmem s;
        s.slots = W + 1;
        size_t o = 0;
        s.kbar_off = o; o += (((size_t)n_cells + 2 + 1) & ~size_t(1)) * sizeof(double);
        s.psi_off = o;  o += (size_t)2 * RN_TR * nrp * sizeof(double);
        s.a_off = o;    o += (size_t)2 * RN_TR * s.slots * sizeof(double);
        s.col_off = o;  o += (size_t)2 * RN_TR * s.slots * sizeof(int);
        s.total = (o + 15) & ~size_t(15);
        return s;
    }
};

__global__ void __launch_bounds__(256, 2) rom_nodal_gram_kernel(PcgOp op, const double* __restrict__ kin,  // (N, n)
                                                             long long s_begin, long long s_end,
                                                             const double* __restrict__ phi,  // [n][nrp], zero padded
                                                             int nr, int TT, double* __restrict__ C /*[chunk][Taug]*/) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int NT = blockDim.x, tid = threadIdx.x;
    const int n = op.n, ld = op.ld, W = op.W, nc = op.n_cells, nrp = 6 * TT, Taug = rom_taug(nr);
    const RomNodalSmem L = RomNodalSmem::make(nc, nrp, W);
    double* s_kbar = reinterpret_cast<double*>(smem + L.kbar_off);
    double* s_psi = reinterpret_cast<double*>(smem + L.psi_off);   // [2][TR][nrp]
    double* s_a = reinterpret_cast<double*>(smem + L.a_off);       // [2][TR][slots]
    int* s_col = reinterpret_cast<int*>(smem + L.col_off);         // [2][TR][slots]
    const int slots = L.slots;
    const size_t plane = (size_t)W * ld;

    // lower-triangular tile of this thread: tid -> (ty >= tx)
    int ty = 0, tx = 0;
    const bool has_tile = tid < TT * (TT + 1) / 2;
    if (has_tile) {
        ty = (int)((sqrt(8.0 * tid + 1.0) - 1.0) * 0.5);
        while ((ty + 1) * (ty + 2) / 2 <= tid) ++ty;
        while (ty * (ty + 1) / 2 > tid) --ty;
        tx = tid - ty * (ty + 1) / 2;
    }
    const int n_tiles = (n + RN_TR - 1) / RN_TR;

    // operator entries of the rows of tile t: slot 0 = diagonal, slots 1..W = the ELL slots
    auto rows_of = [&](int t) {
        double* a = s_a + (size_t)(t & 1) * RN_TR * slots;
        int* cl = s_col + (size_t)(t & 1) * RN_TR * slots;
        for (int e = tid; e < RN_TR * slots; e += NT) {
            const int r = e / slots, w = e - r * slots, i = t * RN_TR + r;
            double v = 0.0;
            int c = 0;
            if (i < n) {
                if (w == 0) {
                    v = op.dcst[i];
                    for (int j = op.dptr[i]; j < op.dptr[i + 1]; ++j) v = fma(op.dcoef[j], s_kbar[op.dcell[j]], v);
                    c = i;
                } else {
                    const size_t o = (size_t)(w - 1) * ld + i;
                    v = op.cst[o];
                    v = fma(op.coef[o], s_kbar[op.cell[o]], v);
                    v = fma(op.coef[plane + o], s_kbar[op.cell[plane + o]], v);
                    c = op.col[o];
                }
            }
            a[e] = v;
            cl[e] = c;
        }
    };

    for (long long s = s_begin + blockIdx.x; s < s_end; s += gridDim.x) {
        const double* kk = kin + s * (long long)n;
        __syncthreads();  // previous sample finished with every buffer
        for (int e = tid; e < nc; e += NT) {
            const int a = op.cells[3 * e], b = op.cells[3 * e + 1], c = op.cells[3 * e + 2];
            s_kbar[e] = cell_coefficient(op.coef_mode, __ldg(kk + a), __ldg(kk + b), __ldg(kk + c));  // as in K2
        }
        if (tid == 0) s_kbar[nc] = 0.0;  // sentinel "no cell"
        __syncthreads();
        rows_of(0);
        __syncthreads();

        double acc[6][6];
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = 0; j < 6; ++j) acc[i][j] = 0.0;
        for (int t = 0; t < n_tiles; ++t) {
            // ---- phase A: psi tile t (uses the operator rows staged one iteration earlier), operator rows of t + 1
            {
                double* psi = s_psi + (size_t)(t & 1) * RN_TR * nrp;
                const double* a = s_a + (size_t)(t & 1) * RN_TR * slots;
                const int* cl = s_col + (size_t)(t & 1) * RN_TR * slots;
                for (int e = tid; e < RN_TR * nrp; e += NT) {
                    const int r = e / nrp, c = e - r * nrp;
                    double v = 0.0;
                    for (int w = 0; w < slots; ++w) v = fma(a[r * slots + w], __ldg(phi + (size_t)cl[r * slots + w] * nrp + c), v);
                    if (c == nr) v = (t * RN_TR + r < n) ? op.rhs[t * RN_TR + r] : 0.0;  // augmented column: b
                    psi[e] = v;
                }
                if (t + 1 < n_tiles) rows_of(t + 1);
            }
            __syncthreads();
            // ---- phase B: Gram update from tile t
            if (has_tile) {
                const double* psi = s_psi + (size_t)(t & 1) * RN_TR * nrp;
#pragma unroll 2
                for (int r = 0; r < RN_TR; ++r) {
                    const double2* py = reinterpret_cast<const double2*>(psi + r * nrp + 6 * ty);
                    const double2* px = reinterpret_cast<const double2*>(psi + r * nrp + 6 * tx);
                    const double2 y0 = py[0], y1 = py[1], y2 = py[2], x0 = px[0], x1 = px[1], x2 = px[2];
                    const double yv[6] = {y0.x, y0.y, y1.x, y1.y, y2.x, y2.y};
                    const double xv[6] = {x0.x, x0.y, x1.x, x1.y, x2.x, x2.y};
#pragma unroll
                    for (int i = 0; i < 6; ++i)
#pragma unroll
                        for (int j = 0; j < 6; ++j) acc[i][j] = fma(yv[i], xv[j], acc[i][j]);
                }
            }
        }
        if (has_tile) {
            double* row = C + (size_t)(s - s_begin) * Taug;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const int gi = 6 * ty + i;
                if (gi > nr) continue;  // gi == nr is the B_r row
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const int gj = 6 * tx + j;
                    if (gj <= gi && gj < nr) row[rom_col_off(gj, nr) + (gi - gj)] = acc[i][j];
                }
            }
        }
    }
}

// Unpack the augmented packed [A_r; B_r^T] rows of a chunk into dense outputs (both optional).
__global__ void __launch_bounds__(256) rom_unpack_kernel(const double* __restrict__ C, long long s_begin, long long s_end,
                                                         int nr, double* __restrict__ Ar_out /* (N, nr, nr) */,
                                                         double* __restrict__ Br_out /* (N, nr) */) {
    const int Taug = rom_taug(nr);
    const long long per = (long long)nr * (nr + 1);
    const long long total = (s_end - s_begin) * per;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long sl = e / per;
        const int rem = (int)(e - sl * per), i = rem / nr, j = rem - i * nr;  // i == nr: the B_r row
        const double* row = C + (size_t)sl * Taug;
        const long long s = s_begin + sl;
        if (i == nr) {
            if (Br_out) Br_out[s * nr + j] = row[rom_col_off(j, nr) + (nr - j)];
        } else if (Ar_out) {
            const int lo = min(i, j), hi = max(i, j);
            Ar_out[(s * nr + i) * nr + j] = row[rom_col_off(lo, nr) + (hi - lo)];
        }
    }
}

}  // namespace tfin
