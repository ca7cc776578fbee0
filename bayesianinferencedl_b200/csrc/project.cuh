// K0: batched sparse row projection (sub-fin averages of nodal fields).  Included by tfin_api.cu only.
#pragma once

#include "pcg_small.cuh"

namespace tfin {

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
