// Micro-benchmark: what HBM bandwidth does the K4 access pattern reach?  148 persistent CTAs, each streaming its own
// 4 vectors (x, r, p, q) of n*S doubles:  P2-like pass (read 4, write 2), P3-like pass (read 2, write 1),
// copy-like pass (read 1, write 1), with U row-groups in flight per thread.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int U, int MODE>
__global__ void __launch_bounds__(1024, 1) pass_kernel(double2* work, size_t vec2, int reps) {
    double2* x = work + (size_t)blockIdx.x * 4 * vec2;
    double2 *r = x + vec2, *p = r + vec2, *q = p + vec2;
    const int T = blockDim.x;
    for (int it = 0; it < reps; ++it) {
        for (size_t o = threadIdx.x; o < vec2; o += (size_t)U * T) {
            double2 xv[U], rv[U], pv[U], qv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const size_t oo = o + (size_t)u * T;
                if (oo < vec2) {
                    if (MODE == 0) { xv[u] = x[oo]; rv[u] = r[oo]; pv[u] = p[oo]; qv[u] = q[oo]; }
                    if (MODE == 1) { rv[u] = r[oo]; pv[u] = p[oo]; }
                    if (MODE == 2) { pv[u] = p[oo]; }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const size_t oo = o + (size_t)u * T;
                if (oo < vec2) {
                    if (MODE == 0) {
                        x[oo] = make_double2(fma(0.5, pv[u].x, xv[u].x), fma(0.5, pv[u].y, xv[u].y));
                        r[oo] = make_double2(fma(-0.5, qv[u].x, rv[u].x), fma(-0.5, qv[u].y, rv[u].y));
                    }
                    if (MODE == 1) p[oo] = make_double2(fma(0.5, pv[u].x, rv[u].x), fma(0.5, pv[u].y, rv[u].y));
                    if (MODE == 2) q[oo] = pv[u];
                }
            }
        }
        __syncthreads();
    }
}

template <int U, int MODE>
int run(double2* work, size_t vec2, int threads, const char* name) {
    const int reps = 20;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    pass_kernel<U, MODE><<<148, threads>>>(work, vec2, 2);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    pass_kernel<U, MODE><<<148, threads>>>(work, vec2, reps);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double vecs = MODE == 0 ? 6 : MODE == 1 ? 3 : 2;
    const double bytes = vecs * vec2 * 16.0 * 148 * reps;
    printf("%-8s U=%d threads=%4d: %.2f ms  %.0f GB/s\n", name, U, threads, ms, bytes / ms / 1e6);
    return 0;
}

int main() {
    const size_t n = 99945, S = 8, vec2 = n * S / 2;
    double2* work;
    CK(cudaMalloc(&work, 148 * 4 * vec2 * 16));
    CK(cudaMemset(work, 0, 148 * 4 * vec2 * 16));
    for (int threads : {1024, 512}) {
        run<1, 0>(work, vec2, threads, "P2-like"); run<2, 0>(work, vec2, threads, "P2-like"); run<4, 0>(work, vec2, threads, "P2-like");
        run<1, 1>(work, vec2, threads, "P3-like"); run<2, 1>(work, vec2, threads, "P3-like"); run<4, 1>(work, vec2, threads, "P3-like");
        run<1, 2>(work, vec2, threads, "copy"); run<2, 2>(work, vec2, threads, "copy"); run<4, 2>(work, vec2, threads, "copy"); run<8, 2>(work, vec2, threads, "copy");
    }
    return 0;
}
