// C: many-chain preconditioned Crank-Nicolson Metropolis on the Gaussian-field prior.
//   state z_c ~ N(0, I_n) coordinates of chain c,  k_c = exp(0.5 chol^T z_c)   (generate_fin_dataset.py:87-88)
//   misfit  Phi(z) = 0.5 ||qoi(k(z)) - data||^2 / sigma^2                      (pymc_func_bayes_inverse.py:76, 201)
//   proposal z' = sqrt(1 - beta^2) z + beta xi,  accept with probability min(1, exp(Phi(z) - Phi(z')))
// The forward solves are the batched kernels of this library (one launch for all chains); the two kernels below are the
// glue that keeps the whole step on the device: proposal (fused Philox normals) and accept/reject + running statistics.
// Every random number of chain g at step t comes from the Philox stream (seed, g, t), see field.cuh.
#pragma once

#include "field.cuh"

namespace tfin {

__global__ void __launch_bounds__(256) pcn_propose_kernel(const double* __restrict__ z, double beta,
                                                          unsigned long long seed, uint32_t step, long long chain0,
                                                          long long C, int n, double* __restrict__ zp) {
    const int ppr = (n + 1) / 2;
    const long long total = C * ppr;
    const double a = sqrt(1.0 - beta * beta);
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long c = e / ppr;
        const int p = (int)(e - c * ppr);
        double x0, x1;
        philox_normal2(seed, (unsigned long long)(chain0 + c), (uint32_t)p, step, x0, x1);
        const long long o = c * n + 2 * p;
        zp[o] = fma(beta, x0, a * z[o]);
        if (2 * p + 1 < n) zp[o + 1] = fma(beta, x1, a * z[o + 1]);
    }
}

struct ChainState {
    double* z;        // (C, n) current coordinates
    double* k;        // (C, n) current conductivity field | nullptr (only needed for k_sum)
    double* qoi;      // (C, n_obs) observables of the current state
    double* phi;      // (C) misfit of the current state
    unsigned long long* accepted;  // (C)
    double* qoi_sum;  // (C, n_obs) running sum over steps of the current observables
    double* qoi_sq;   // (C, n_obs) running sum of squares
    double* k_sum;    // (C, n) running sum of the current field | nullptr
};

// One warp per chain.  init != 0: adopt the proposal unconditionally (first evaluation of the start state).
__global__ void __launch_bounds__(256) chain_accept_kernel(ChainState st, const double* zp,   // zp may alias st.z (init)
                                                           const double* kp, const double* __restrict__ qp,
                                                           const int* __restrict__ status_p, const double* __restrict__ data,
                                                           double inv_sigma2, unsigned long long seed, uint32_t step,
                                                           long long chain0, long long C, int n, int n_obs, int init) {
    const int lane = threadIdx.x & 31;
    const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long c = wid; c < C; c += nw) {
        double s = 0.0;
        for (int o = lane; o < n_obs; o += 32) {
            const double r = qp[c * n_obs + o] - data[o];
            s = fma(r, r, s);
        }
        s = warp_sum(s);
        double phi_p = 0.5 * s * inv_sigma2;
        const bool valid = status_p[c] == TFIN_STATUS_CONVERGED && isfinite(phi_p);
        if (!valid) phi_p = INFINITY;
        bool accept;
        if (init) {
            accept = true;
        } else {
            double u1, u2;
            philox_uniform2(seed, (unsigned long long)(chain0 + c), PHILOX_PAIR_UNIFORM, step, u1, u2);
            // lane 0 reads the current misfit and broadcasts it: lane 0 also overwrites it below, and the other lanes
            // must not depend on the warp staying converged in between
            double phi_cur = lane == 0 ? st.phi[c] : 0.0;
            phi_cur = __shfl_sync(0xffffffffu, phi_cur, 0);
            accept = valid && log(u1) < phi_cur - phi_p;
        }
        if (accept) {
            const bool copy_z = zp != st.z;   // the start state is evaluated in place (init): nothing to copy
            for (int i = lane; i < n; i += 32) {
                if (copy_z) st.z[c * n + i] = zp[c * n + i];
                if (st.k) st.k[c * n + i] = kp[c * n + i];
            }
            for (int o = lane; o < n_obs; o += 32) st.qoi[c * n_obs + o] = qp[c * n_obs + o];
            if (lane == 0) {
                st.phi[c] = phi_p;
                if (!init) st.accepted[c] += 1ULL;
            }
        }
        __syncwarp();
        if (!init) {
            for (int o = lane; o < n_obs; o += 32) {
                const double q = st.qoi[c * n_obs + o];
                st.qoi_sum[c * n_obs + o] += q;
                st.qoi_sq[c * n_obs + o] = fma(q, q, st.qoi_sq[c * n_obs + o]);
            }
            if (st.k_sum)
                for (int i = lane; i < n; i += 32) st.k_sum[c * n + i] += st.k[c * n + i];
        }
    }
}

}  // namespace tfin
