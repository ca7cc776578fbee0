// Micro-benchmark for the K4 SpMV pass (P1): 148 persistent CTAs x 1024 threads, each CTA owns vectors p, q of
// n rows x S=8 samples (interleaved, lane = 2 samples of one row, warp = 8 rows).  q_i = sum_k c_k p[i + off_k]
// with the 7-point structured offsets.  Variants:
//   0 direct gathers          1 + prefetch.global.L2 PF steps ahead     2 + prefetch.global.L1 PF steps ahead
//   3 two row groups per warp in flight        4 p window staged in a shared-memory ring by TMA bulk copies
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int S = 8, LPR = 4, RPW = 8, L = 261;
__constant__ int c_off[7];
__constant__ double c_coef[7];

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int VAR>
__global__ void __launch_bounds__(1024, 1) gather_kernel(const double2* __restrict__ pw, double2* __restrict__ qw, int n, int reps) {
    const size_t vec2 = (size_t)n * LPR;
    const double2* p = pw + (size_t)blockIdx.x * vec2;
    double2* q = qw + (size_t)blockIdx.x * vec2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = 32;
    const int sp = lane % LPR, rw = lane / LPR;
    const int n_groups = n / RPW;
    constexpr int PF = 4;
    for (int it = 0; it < reps; ++it) {
        if (VAR <= 2) {
            for (int g = warp; g < n_groups; g += nwarps) {
                const int i = g * RPW + rw;
                if (VAR == 1 && i + PF * 256 + L + 1 < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + (size_t)(i + PF * 256 + L + 1) * LPR + sp));
                if (VAR == 2 && i + PF * 256 + L + 1 < n) asm volatile("prefetch.global.L1 [%0];" ::"l"(p + (size_t)(i + PF * 256 + L + 1) * LPR + sp));
                double2 acc = make_double2(0, 0);
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    int c = i + c_off[k];
                    c = c < 0 ? 0 : (c >= n ? n - 1 : c);
                    const double2 pv = p[(size_t)c * LPR + sp];
                    acc.x = fma(c_coef[k], pv.x, acc.x);
                    acc.y = fma(c_coef[k], pv.y, acc.y);
                }
                q[(size_t)i * LPR + sp] = acc;
            }
        } else if (VAR == 3) {
            for (int g = warp; g < n_groups; g += 2 * nwarps) {
                const int ia = g * RPW + rw, ib = ia + nwarps * RPW;
                double2 pa[7], pb[7];
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    int c = ia + c_off[k];
                    c = c < 0 ? 0 : (c >= n ? n - 1 : c);
                    pa[k] = p[(size_t)c * LPR + sp];
                    c = ib + c_off[k];
                    c = c < 0 ? 0 : (c >= n ? n - 1 : c);
                    pb[k] = p[(size_t)c * LPR + sp];
                }
                double2 a = make_double2(0, 0), b = make_double2(0, 0);
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    a.x = fma(c_coef[k], pa[k].x, a.x); a.y = fma(c_coef[k], pa[k].y, a.y);
                    b.x = fma(c_coef[k], pb[k].x, b.x); b.y = fma(c_coef[k], pb[k].y, b.y);
                }
                q[(size_t)ia * LPR + sp] = a;
                if (ib < n) q[(size_t)ib * LPR + sp] = b;
            }
        } else {
            // ring of NCH chunks x 256 rows x 64 B in shared memory, filled by TMA bulk copies LA chunks ahead
            constexpr int CH = 256, NCH = 8, LA = 4, CHB = CH * S * 8;
            extern __shared__ __align__(128) unsigned char ring[];
            __shared__ uint64_t full[NCH];
            const int n_chunks = n / CH;  // n multiple of 256 in this benchmark
            if (tid == 0) {
                for (int c = 0; c < NCH; ++c) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[c])));
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncthreads();
            auto issue = [&](int c) {
                const int slot = c % NCH;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[slot])), "r"(CHB) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(ring + (size_t)slot * CHB)),
                             "l"(p + (size_t)c * CH * LPR), "r"(CHB), "r"(smem_u32(&full[slot])) : "memory");
            };
            auto wait = [&](int c) {
                const uint32_t parity = (uint32_t)((c / NCH) & 1), bar = smem_u32(&full[c % NCH]);
                uint32_t done = 0;
                while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
            };
            if (tid == 0) for (int c = 0; c < LA + 2 && c < n_chunks; ++c) issue(c);
            const double2* rp = reinterpret_cast<const double2*>(ring);
            for (int t = 0; t < n_chunks; ++t) {
                // chunk t+2 (covers i + L + 1 for L < 256) must have landed; all warps passed step t-1 => slot of
                // chunk t+LA+2-NCH... is free: keep it simple with one CTA barrier per step
                if (t + 2 < n_chunks) wait(t + 2); else wait(n_chunks - 1);
                const int i = t * CH + warp * RPW + rw;
                double2 acc = make_double2(0, 0);
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    int c = i + c_off[k];
                    c = c < 0 ? 0 : (c >= n ? n - 1 : c);
                    const double2 pv = rp[(size_t)(c % (NCH * CH)) * LPR + sp];
                    acc.x = fma(c_coef[k], pv.x, acc.x);
                    acc.y = fma(c_coef[k], pv.y, acc.y);
                }
                q[(size_t)i * LPR + sp] = acc;
                __syncthreads();
                if (tid == 0 && t + LA + 2 < n_chunks) issue(t + LA + 2);   // slot of chunk t+LA+2-NCH = t-2: no longer needed (t-2 < t+1-1-1)
            }
            __syncthreads();
        }
        __syncthreads();
    }
}

template <int VAR>
int run(const double2* p, double2* q, int n, const char* name) {
    const int reps = 20;
    const size_t smem = VAR == 4 ? 8 * 256 * 64 : 0;
    CK(cudaFuncSetAttribute(gather_kernel<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather_kernel<VAR><<<148, 1024, smem>>>(p, q, n, 2);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    gather_kernel<VAR><<<148, 1024, smem>>>(p, q, n, reps);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = 2.0 * n * S * 8 * 148 * reps;
    printf("%-28s: %.2f ms  %.0f GB/s (algorithmic 16 n S bytes per pass)\n", name, ms, bytes / ms / 1e6);
    return 0;
}

int main() {
    const int n = 99840;  // multiple of 256
    const size_t vec2 = (size_t)n * LPR;
    double2 *p, *q;
    CK(cudaMalloc(&p, 148 * vec2 * 16)); CK(cudaMalloc(&q, 148 * vec2 * 16));
    CK(cudaMemset(p, 0, 148 * vec2 * 16)); CK(cudaMemset(q, 0, 148 * vec2 * 16));
    const int off[7] = {-L - 1, -L, -1, 0, 1, L, L + 1};
    const double coef[7] = {-0.5, -1, -1, 4, -1, -1, -0.5};
    CK(cudaMemcpyToSymbol(c_off, off, sizeof(off))); CK(cudaMemcpyToSymbol(c_coef, coef, sizeof(coef)));
    run<0>(p, q, n, "direct gathers");
    run<1>(p, q, n, "+ prefetch.L2 4 steps ahead");
    run<2>(p, q, n, "+ prefetch.L1 4 steps ahead");
    run<3>(p, q, n, "2 groups / warp in flight");
    run<4>(p, q, n, "smem ring via TMA bulk");
    return 0;
}
