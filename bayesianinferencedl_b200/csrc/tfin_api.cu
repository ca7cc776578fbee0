// libtfin: C ABI (include/tfin.h) over the sm_100a kernels.  Host side: CSR -> ELL conversion, per-handle
// device workspaces, launch geometry, host<->device staging.  No CPU compute path exists here.
#include <algorithm>
#include <cmath>
#include <climits>
#include <cstring>
#include <map>

#include "common.cuh"
#include "pcg_small.cuh"
#include "pcg_variants.h"
#include "project.cuh"
#include "pcg_stream.cuh"
#include "pcg_f32.cuh"
#include "rom.cuh"
#include "rom_nodal.cuh"
#include "field.cuh"
#include "chains.cuh"
#include "frontal_host.h"
#include "frontal.cuh"

using namespace tfin;

namespace tfin {
#define DECL_G(g) const PcgVariant* pcg_variants_g##g##_n0(int*); const PcgVariant* pcg_variants_g##g##_n1(int*);
DECL_G(0) DECL_G(1) DECL_G(2) DECL_G(3) DECL_G(4)
#undef DECL_G
const PcgVariant* pcg_variants(int group, int nodal, int* count) {
    switch (group * 2 + (nodal ? 1 : 0)) {
#define CASE_G(g) case g * 2: return pcg_variants_g##g##_n0(count); case g * 2 + 1: return pcg_variants_g##g##_n1(count);
        CASE_G(0) CASE_G(1) CASE_G(2) CASE_G(3) CASE_G(4)
#undef CASE_G
    }
    *count = 0;
    return nullptr;
}
}  // namespace tfin

// Reverse Cuthill-McKee ordering of the (symmetric) CSR pattern: order[new] = old.  Every connected component is
// started from a pseudo-peripheral node (repeated BFS from the farthest minimum-degree node).
static std::vector<int> rcm_order(int n, const int32_t* rp, const int32_t* ci) {
    std::vector<int> order, level(n, -1), deg(n);
    order.reserve(n);
    for (int i = 0; i < n; ++i) deg[i] = rp[i + 1] - rp[i];
    std::vector<char> done(n, 0);
    std::vector<int> comp, nbr;
    auto bfs = [&](int start, std::vector<int>& out) -> int {  // BFS over not-yet-ordered nodes; returns eccentricity
        out.clear();
        out.push_back(start);
        level[start] = 0;
        for (size_t h = 0; h < out.size(); ++h) {
            const int u = out[h];
            nbr.clear();
            for (int j = rp[u]; j < rp[u + 1]; ++j) {
                const int v = ci[j];
                if (v != u && !done[v] && level[v] < 0) {
                    level[v] = level[u] + 1;
                    nbr.push_back(v);
                }
            }
            std::sort(nbr.begin(), nbr.end(), [&](int a, int b) { return deg[a] != deg[b] ? deg[a] < deg[b] : a < b; });
            out.insert(out.end(), nbr.begin(), nbr.end());
        }
        const int ecc = level[out.back()];
        return ecc;
    };
    for (int seed = 0; seed < n; ++seed) {
        if (done[seed]) continue;
        int start = seed, ecc = -1;
        for (int round = 0; round < 8; ++round) {
            const int e = bfs(start, comp);
            int far = comp.back();
            for (int v : comp)
                if (level[v] == e && deg[v] < deg[far]) far = v;
            for (int v : comp) level[v] = -1;
            if (e <= ecc) break;
            ecc = e;
            start = far;
        }
        bfs(start, comp);
        for (int v : comp) {
            done[v] = 1;
            order.push_back(v);
        }
    }
    std::reverse(order.begin(), order.end());
    return order;
}

// Slot assignment of the on-chip ELL.  A warp owns 32 CONSECUTIVE rows and gathers slot w of all of them with one
// shared-memory load, which is conflict-free when the 32 columns are consecutive too, i.e. when slot w means the same
// relative offset (col - row) for the whole warp.  CSR order breaks that at every row with a missing neighbour
// (boundaries, junctions): ncu showed 24-41 wavefronts instead of 20 on the gather loads.  So a row inherits, offset
// by offset, the slots of the previous row; entries with a new offset take the free slots.  keep[i] comes in as CSR
// positions of the kept off-diagonals and leaves as a slot-aligned list (-1 = padding).
static void align_ell_slots(int n, int W, const int32_t* col_idx, std::vector<std::vector<int>>& keep) {
    const int NONE = INT32_MIN;
    std::vector<int> prev_off(W, NONE), cur_off(W), slot(W);
    for (int i = 0; i < n; ++i) {
        std::fill(slot.begin(), slot.end(), -1);
        std::fill(cur_off.begin(), cur_off.end(), NONE);
        std::vector<int> rest;
        for (int j : keep[i]) {
            const int off = col_idx[j] - i;
            int w = 0;
            while (w < W && !(prev_off[w] == off && slot[w] < 0)) ++w;
            if (w < W) {
                slot[w] = j;
                cur_off[w] = off;
            } else {
                rest.push_back(j);
            }
        }
        for (int j : rest) {  // new offsets: prefer a slot the previous row left empty, else any free one
            int w = 0;
            while (w < W && !(slot[w] < 0 && prev_off[w] == NONE)) ++w;
            if (w == W) {
                w = 0;
                while (slot[w] >= 0) ++w;
            }
            slot[w] = j;
            cur_off[w] = col_idx[j] - i;
        }
        keep[i].assign(slot.begin(), slot.end());
        for (int w = 0; w < W; ++w)
            if (cur_off[w] != NONE) prev_off[w] = cur_off[w];  // a gap keeps the offset of the last row that had one
    }
}

#define FRONTAL_WORK_BYTES ((long long)6 << 30)   // HBM workspace of the split launch: factor blocks of one chunk
// Device limits and tuning knobs FrontalSet::upload sizes the sample-per-thread kernels for.
struct FrontalCfg {
    int smem_optin, smem_per_sm, sm_count;
    int want_lanes;   // "frontal_lanes": samples per warp, 0 = auto
    int ring_rows;    // "frontal_ring_rows": rows of the substitution kernel's factor-block ring, 0 = auto
};

// Device copy of one frontal program (frontal_host.h) -- the sparse-direct solver D1 / D2.
struct FrontalSet {
    bool ok = false;
    std::string why = "no operator";
    FrontalProgram host;     // D2 (slots of column j + 1 allocated one step early)
    FrontalProgram host1;    // D1 (no lookahead: one slot fewer, i.e. ~8 % less shared memory per sample at n = 1597)
    FrontalStreams streams;
    int ncv = 0;
    DevBuf<unsigned char> fwd, bwd, fwd1, bwd1, fsub1;
    void release() {
        fwd.release();
        bwd.release();
        fwd1.release();
        bwd1.release();
        fsub1.release();
        ok = false;
    }
    // (re)pack the instruction streams from the host program and copy them to the device.  D1 geometry is fixed here
    // (n_obs and the affine coefficient rows are known): samples per warp `lanes` (rows of 8 * lanes bytes) and the rows of
    // the factor-block ring, which gets the shared memory that the resident warps leave free.
    int upload(cudaStream_t st, int n_obs, const FrontalCfg& cfg) {
        const int smem_optin = cfg.smem_optin, smem_per_sm = cfg.smem_per_sm, want_lanes = cfg.want_lanes;
        const int ntri = host1.nslots * (host1.nslots + 1) / 2;
        const int base_rows = ntri + host1.nslots + n_obs + (ncv <= TFIN_MAX_TERMS ? ncv : 0);
        // Throughput of the (latency-bound) kernel is samples in flight / pass latency, and samples in flight = resident
        // warps x samples per warp are bounded by the SM's shared memory (228 KiB, 1 KiB reserved per CTA): pick the
        // (warps, lanes) pair with the largest product, fewer warps on a tie.  n = 1597: 3 x 27 = 81 instead of 2 x 32.
        // factor-block ring of the substitution kernel (split launch: no front in its shared memory): ~4 blocks ahead
        long long lr = std::max<long long>(host.cmax + 2, std::min<long long>(64, 4LL * (host.cmax + 2)));
        if (cfg.ring_rows > 0) lr = std::max<long long>(host.cmax + 2, cfg.ring_rows);
        frontal_pack_streams(host, (int)lr, FRONTAL_DMAX, 32, &streams, &host1);
        const long long rows_f = base_rows - n_obs;   // factor kernel: front + rhs + coefficients (+ its instruction ring)
        int lanes = 0, warps_f = 1;
        if (host.cmax <= 32) {
            if (want_lanes >= 4 && want_lanes <= 32) {
                if (rows_f * 8 * want_lanes + 16 + streams.ring_fwd1 <= (long long)smem_optin) lanes = want_lanes;
                if (lanes) warps_f = (int)std::max<long long>(1, smem_per_sm / (rows_f * 8 * lanes + 16 + streams.ring_fwd1 + 1024));
            } else {
                long long best = 0;
                for (int w = 1; w <= 6; ++w) {
                    const long long per_cta = std::min<long long>(smem_optin, smem_per_sm / w - 1024);
                    const long long l = std::min<long long>(32, (per_cta - streams.ring_fwd1 - 16) / (rows_f * 8));
                    if (l >= 8 && w * l > best) {
                        best = w * l;
                        lanes = (int)l;
                        warps_f = w;
                    }
                }
            }
        }
        if (lanes == 0) lr = 0;   // D1 cannot serve this front
        if (lanes && cfg.ring_rows == 0) {
            // One pass of the substitution kernel costs about half a factor pass however few warps it carries, so it should
            // cover a whole chunk (the factor waves whose blocks fit the HBM workspace, launch_frontal): shrink its block ring
            // (less prefetch distance; >= ~3 blocks stay ahead) until enough of its warps are resident for that.
            const long long per_sample = ((long long)host.nnzL + 2LL * host.n + 2) * 8;
            const long long wave = (long long)cfg.sm_count * warps_f * lanes;
            const long long mem_waves = std::max<long long>(1, std::min<long long>(4, FRONTAL_WORK_BYTES / per_sample / wave));
            const long long lr_min = std::max<long long>(host.cmax + 2, std::min<long long>(lr, 3LL * (host.cmax / 2 + 2)));
            for (; lr > lr_min; lr -= 4) {
                const long long sm_b = (host1.nslots + n_obs + lr) * 8LL * lanes + 16 + streams.ring_bytes + 1024;
                if (smem_per_sm / sm_b >= warps_f * mem_waves) break;
            }
            lr = std::max(lr, lr_min);
        }
        if (lanes != 32 || lr != streams.lr_rows)
            frontal_pack_streams(host, (int)lr, FRONTAL_DMAX, lanes ? lanes : 32, &streams, &host1);
        if (int e = fwd.upload(streams.fwd, st)) return e;
        if (int e = bwd.upload(streams.bwd, st)) return e;
        if (int e = fwd1.upload(streams.fwd1, st)) return e;
        if (int e = fsub1.upload(streams.fsub1, st)) return e;
        return bwd1.upload(streams.bwd1, st);
    }
    FrontalDev dev(bool lane_kernel) const {
        FrontalDev d{};
        d.n = host.n;
        d.nslots = lane_kernel ? host1.nslots : host.nslots;
        d.cmax = host.cmax;
        d.ncv = ncv;
        d.ntri = d.nslots * (d.nslots + 1) / 2;
        d.ring_bytes = streams.ring_bytes;
        d.lr_rows = streams.lr_rows;
        d.lanes = streams.lanes;
        d.ring_fwd1 = streams.ring_fwd1;
        d.nnzL = host.nnzL;
        d.fwd = lane_kernel ? fwd1.p : fwd.p;
        d.bwd = lane_kernel ? bwd1.p : bwd.p;
        d.fsub = fsub1.p;
        return d;
    }
};

struct tfin_ctx {
    int device = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    int smem_per_sm = 0;
    cudaStream_t stream = nullptr;
    int64_t launches = 0;

    // ---- affine operator (K1)
    int n = 0, ld = 0, n_terms = 0, W = 0;
    std::vector<int32_t> h_row_ptr, h_col_idx;    // kept for tfin_set_cells
    std::vector<double> h_const;                  // vals[0] on the CSR pattern
    DevBuf<uint16_t> d_col;
    DevBuf<double> d_val, d_diag, d_rhs;
    bool small_ok = false;                        // on-chip kernels usable (n <= 8191)
    // ---- term-tagged CSR for the streaming kernel (K4)
    int s_ldr = 0, s_We = 0, s_bandwidth = 0;
    bool stream_ok = false;
    std::vector<int> h_sinv;       // caller's dof -> renumbered row of the streaming path
    DevBuf<int> d_sperm, d_sobs_idx;
    int stream_ring = -1;          // -1 auto (= direct gathers), 0 direct gathers, 1 shared-memory ring (opt-in)
    int last_ring = 0;
    DevBuf<uint32_t> d_scolterm;
    DevBuf<unsigned char> d_scnt;
    DevBuf<double> d_scoef, d_srhs, d_swork;
    DevBuf<unsigned long long> d_sprof;  // CTA-0 clocks per pass (P1, P2, P3, iterations) when stream_prof is set
    int stream_prof = 0;
    int stream_pad_smem = 0;  // experiment knob: extra (unused) dynamic shared memory in KB
    int pcg_path = 0;     // 0 auto, 1 on-chip, 2 streaming
    int precision = 64;   // 64, or 32: optional single-precision on-chip affine PCG (K1f)
    int stream_tile = 0;  // 0 auto, else 8 / 16 / 32
    int last_path = 0, last_tile = 0;
    // ---- observation / averaging
    int n_obs = 0, n_avg = 0;
    DevBuf<int> d_obs_ptr, d_obs_idx, d_avg_ptr, d_avg_idx;
    DevBuf<double> d_obs_val, d_avg_val;
    // ---- nodal operator (K2)
    int n_cells = 0, Wn = 0;
    int coef_mode = 0;  // 0: k, 1: exp(k) (forward_solve_exp.py)
    DevBuf<uint16_t> d_ncol;
    DevBuf<int> d_ncell, d_dptr, d_dcell, d_cells;
    DevBuf<double> d_ncoef, d_ncst, d_dcoef, d_dcst, d_Ke;
    // ---- transpose of the observation operator (adjoint right-hand sides)
    DevBuf<int> d_obsT_ptr, d_obsT_idx;
    DevBuf<double> d_obsT_val, d_data, d_grad, d_cost;
    // ---- ROM (K3)
    int n_r = 0, rom_terms = 0, rom_obs = 0;
    DevBuf<double> d_S, d_obs_phi, d_romC;
    int64_t rom_chunk = 0;  // 0 = auto
    // ---- many-chain pCN (C)
    DevBuf<double> d_cz, d_czp, d_ck, d_ckp, d_cq, d_cqp, d_cphi, d_cqs, d_cqq, d_cks, d_cdata;
    DevBuf<unsigned long long> d_cacc;
    DevBuf<int> d_cstat;
    // ---- Gaussian-field sampler (F): lower Cholesky factor of the covariance, normals
    int f_n = 0;
    DevBuf<double> d_fL, d_fz, d_fk, d_fxy;
    DevBuf<int> d_finfo;
    // ---- nodal LSPG (R4): padded basis [n][6 TT], projection rows
    int b_nr = 0, b_nout = 0, b_TT = 0;
    DevBuf<double> d_bphi, d_bout, d_Ar, d_Br, d_y;
    // ---- ROM gradient (R3): Gram blocks Psi_t^T Psi_q, transposed averaging operator
    int rg_ob = 0;
    DevBuf<double> d_NG, d_vr, d_gtheta, d_avgT_val;
    DevBuf<int> d_avgT_ptr, d_avgT_idx;
    // ---- batch staging / scratch
    DevBuf<double> d_in, d_theta, d_w, d_qoi, d_relres, d_wr;
    // ---- pipelined host path of the nodal solve: second stream, two input / solution staging buffers, events
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    DevBuf<double> d_pin[2], d_pw[2];
    int64_t host_chunk = 8192;  // samples per pipelined chunk (0 disables the pipeline)
    DevBuf<int> d_iters, d_status;
    DevBuf<unsigned long long> d_counter;
    // ---- sparse-direct solver (D1 / D2): programs of the affine and of the nodal operator, workspaces, knobs
    FrontalSet fr_aff, fr_nod;
    std::vector<int32_t> h_obs_ptr, h_obs_idx;
    std::vector<double> h_obs_val;
    DevBuf<double> d_fwork, d_fcv, d_fw, d_fv;
    int fom_solver = 0;      // 0 auto (direct where the front fits on chip), 1 PCG, 2 direct
    int frontal_kernel = 0;  // 0 auto, 1 = D1 (sample per thread), 2 = D2 (sample per CTA)
    int frontal_threads = 0; // D2 threads per CTA, 0 = auto
    int frontal_mode = -1;   // D2: -1 auto, 0 = QOI (extra right-hand sides), 1 = SOLVE (factor to HBM + backward)
    int frontal_lanes = 0;   // D1: samples per warp, 0 = auto (most samples in flight), else 4..32
    int frontal_ring_rows = 0;   // D1: rows of the substitution kernel's factor-block ring, 0 = auto
    FrontalCfg frontal_cfg() const { return FrontalCfg{max_smem_optin, smem_per_sm, sm_count, frontal_lanes, frontal_ring_rows}; }
    int frontal_split = 1;   // D1: 1 = factor kernel + substitution kernel, 0 = one fused kernel
    int last_focc_b = 0;
    int last_solver = 0, last_fkernel = 0, last_fthreads = 0, last_focc = 0;
    size_t last_fsmem = 0;
    // ---- tuning
    int pcg_R = 0;    // rows per thread, 0 = auto
    int pcg_WR = -1;  // -1 auto, 0 = ELL values in shared memory, 1 = in registers
    int last_T = 0, last_R = 0, last_occ = 0, last_WT = 0, last_WR = 0;
    size_t last_smem = 0;
};

// ------------------------------------------------------------------------------------------------
extern "C" int tfin_version(void) { return 120; }
extern "C" const char* tfin_last_error(void) { return last_error().c_str(); }

extern "C" int tfin_create(int device, tfin_handle_t* out) {
    if (!out) return fail(TFIN_E_ARG, "tfin_create: out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(TFIN_E_CUDA, "tfin_create: no CUDA device (%s); libtfin has no CPU fallback",
                    cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(TFIN_E_ARG, "tfin_create: device %d out of range", device);
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    TFIN_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(TFIN_E_CUDA, "tfin_create: device %d is sm_%d%d; libtfin is built for sm_100a only", device,
                    prop.major, prop.minor);
    tfin_ctx* c = new tfin_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    c->smem_per_sm = (int)prop.sharedMemPerMultiprocessor;
    TFIN_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    if (int err = c->d_counter.reserve(1)) return err;
    *out = c;
    return 0;
}

extern "C" int tfin_destroy(tfin_handle_t h) {
    if (!h) return 0;
    DeviceGuard guard(h->device);
    cudaDeviceSynchronize();
    for (auto* b : {&h->d_val, &h->d_diag, &h->d_rhs, &h->d_obs_val, &h->d_avg_val, &h->d_ncoef, &h->d_ncst,
                    &h->d_dcoef, &h->d_dcst, &h->d_S, &h->d_obs_phi, &h->d_romC, &h->d_in, &h->d_theta, &h->d_w,
                    &h->d_qoi, &h->d_relres, &h->d_wr})
        b->release();
    for (auto* b : {&h->d_obs_ptr, &h->d_obs_idx, &h->d_avg_ptr, &h->d_avg_idx, &h->d_ncell, &h->d_dptr,
                    &h->d_dcell, &h->d_cells, &h->d_iters, &h->d_status})
        b->release();
    for (auto* b : {&h->d_scoef, &h->d_srhs, &h->d_swork, &h->d_Ke, &h->d_obsT_val, &h->d_data, &h->d_grad, &h->d_cost})
        b->release();
    h->d_obsT_ptr.release();
    h->d_obsT_idx.release();
    for (auto* b : {&h->d_fL, &h->d_fz, &h->d_fk, &h->d_fxy}) b->release();
    for (auto* b : {&h->d_cz, &h->d_czp, &h->d_ck, &h->d_ckp, &h->d_cq, &h->d_cqp, &h->d_cphi, &h->d_cqs, &h->d_cqq, &h->d_cks,
                    &h->d_cdata})
        b->release();
    h->d_cacc.release();
    h->d_cstat.release();
    h->d_finfo.release();
    for (auto* b : {&h->d_NG, &h->d_vr, &h->d_gtheta, &h->d_avgT_val, &h->d_bphi, &h->d_bout, &h->d_Ar, &h->d_Br, &h->d_y})
        b->release();
    h->d_avgT_ptr.release();
    h->d_avgT_idx.release();
    h->d_scolterm.release();
    h->d_scnt.release();
    h->d_sperm.release();
    h->d_sobs_idx.release();
    h->d_sprof.release();
    h->d_col.release();
    h->d_ncol.release();
    h->d_counter.release();
    for (int b = 0; b < 2; ++b) {
        h->d_pin[b].release();
        h->d_pw[b].release();
        if (h->ev_in[b]) cudaEventDestroy(h->ev_in[b]);
        if (h->ev_comp[b]) cudaEventDestroy(h->ev_comp[b]);
        if (h->ev_out[b]) cudaEventDestroy(h->ev_out[b]);
    }
    h->fr_aff.release();
    h->fr_nod.release();
    h->d_fwork.release();
    h->d_fcv.release();
    h->d_fw.release();
    h->d_fv.release();
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

#define CHECK_HANDLE(h)                                             \
    if (!(h)) return fail(TFIN_E_ARG, "%s: NULL handle", __func__); \
    DeviceGuard _device_guard((h)->device);

// ------------------------------------------------------------------------------------------------ setup
extern "C" int tfin_set_operator(tfin_handle_t h, int32_t n, int32_t nnz, const int32_t* row_ptr,
                                 const int32_t* col_idx, int32_t n_terms, const double* vals, const double* rhs,
                                 int32_t prune_zeros) {
    CHECK_HANDLE(h);
    if (n <= 0 || nnz <= 0 || !row_ptr || !col_idx || !vals || !rhs)
        return fail(TFIN_E_ARG, "tfin_set_operator: bad argument");
    if (n_terms < 1 || n_terms > TFIN_MAX_TERMS)
        return fail(TFIN_E_ARG, "tfin_set_operator: n_terms must be in [1, %d]", TFIN_MAX_TERMS);
    const bool small_ok = n <= 8191;  // uint16 byte offsets of the on-chip kernels
    if (row_ptr[0] != 0 || row_ptr[n] != nnz) return fail(TFIN_E_ARG, "tfin_set_operator: malformed row_ptr");
    const int ld = (n + 31) & ~31;
    // pass 1: diagonal positions, off-diagonal widths
    std::vector<int> diag_pos(n, -1);
    std::vector<std::vector<int>> keep(n);
    int W = 0;
    for (int i = 0; i < n; ++i) {
        if (row_ptr[i + 1] < row_ptr[i]) return fail(TFIN_E_ARG, "tfin_set_operator: malformed row_ptr");
        for (int j = row_ptr[i]; j < row_ptr[i + 1]; ++j) {
            const int c = col_idx[j];
            if (c < 0 || c >= n) return fail(TFIN_E_ARG, "tfin_set_operator: column index out of range");
            if (c == i) {
                diag_pos[i] = j;
                continue;
            }
            bool nz = !prune_zeros;
            for (int t = 0; t < n_terms && !nz; ++t) nz = vals[(size_t)t * nnz + j] != 0.0;
            if (nz) keep[i].push_back(j);
        }
        if (diag_pos[i] < 0) return fail(TFIN_E_ARG, "tfin_set_operator: row %d has no diagonal entry", i);
        W = std::max(W, (int)keep[i].size());
    }
    if (W == 0) W = 1;
    align_ell_slots(n, W, col_idx, keep);
    std::vector<uint16_t> col((size_t)W * ld);
    std::vector<double> val((size_t)n_terms * W * ld, 0.0), diag((size_t)n_terms * ld, 0.0), b(ld, 0.0);
    for (int w = 0; w < W; ++w)
        for (int i = 0; i < ld; ++i) col[(size_t)w * ld + i] = (uint16_t)std::min(i, n - 1);
    for (int i = 0; i < n; ++i) {
        b[i] = rhs[i];
        for (int t = 0; t < n_terms; ++t) diag[(size_t)t * ld + i] = vals[(size_t)t * nnz + diag_pos[i]];
        for (int w = 0; w < W; ++w) {
            const int j = keep[i][w];
            if (j < 0) {  // padding: own row, value 0
                col[(size_t)w * ld + i] = (uint16_t)i;
                continue;
            }
            col[(size_t)w * ld + i] = (uint16_t)col_idx[j];
            for (int t = 0; t < n_terms; ++t) val[((size_t)t * W + w) * ld + i] = vals[(size_t)t * nnz + j];
        }
    }
    h->n = n;
    h->ld = ld;
    h->n_terms = n_terms;
    h->W = W;
    h->h_row_ptr.assign(row_ptr, row_ptr + n + 1);
    h->h_col_idx.assign(col_idx, col_idx + nnz);
    h->h_const.assign(vals, vals + nnz);
    h->small_ok = small_ok;
    if (small_ok) {
        if (int e = h->d_col.upload(col, h->stream)) return e;
        if (int e = h->d_val.upload(val, h->stream)) return e;
        if (int e = h->d_diag.upload(diag, h->stream)) return e;
    }
    if (int e = h->d_rhs.upload(b, h->stream)) return e;
    h->stream_ok = n < (1 << 24);
    if (h->stream_ok) {  // term-tagged slot-major ELL for the streaming kernel: only non-zero (entry, term) pairs
        struct Ent {
            int col, term;
            double coef;
        };
        // bandwidth-reducing renumbering (internal to the streaming path; callers keep their dof numbering)
        const std::vector<int> perm = rcm_order(n, row_ptr, col_idx);  // new -> old
        std::vector<int> inv(n);
        for (int i = 0; i < n; ++i) inv[perm[i]] = i;
        std::vector<std::vector<Ent>> rows(n);
        int We = 1, bandwidth = 1;
        for (int i = 0; i < n; ++i) {
            const int old = perm[i];
            for (int t = 0; t < n_terms; ++t)  // sorted by term inside the row (theta is cached per term run)
                for (int j = row_ptr[old]; j < row_ptr[old + 1]; ++j) {
                    const double v = vals[(size_t)t * nnz + j];
                    if (v != 0.0) rows[i].push_back(Ent{inv[col_idx[j]], t, v});
                }
            std::stable_sort(rows[i].begin(), rows[i].end(),
                             [](const Ent& a, const Ent& b) { return a.term != b.term ? a.term < b.term : a.col < b.col; });
            for (const Ent& e : rows[i]) bandwidth = std::max(bandwidth, std::abs(e.col - i));
            We = std::max(We, (int)rows[i].size());
        }
        if (We > 255) return fail(TFIN_E_ARG, "tfin_set_operator: more than 255 term-tagged entries in one row");
        We = (We + STREAM_KC - 1) / STREAM_KC * STREAM_KC;
        const int ldr = (n + 511) & ~511;  // whole ring chunks for every tile width
        std::vector<uint32_t> colterm((size_t)We * ldr);
        std::vector<double> coef((size_t)We * ldr, 0.0);
        std::vector<unsigned char> cnt(ldr, 1);
        for (int i = 0; i < ldr; ++i) {
            const int used = i < n ? (int)rows[i].size() : 0;
            const uint32_t pad_term = used ? (uint32_t)rows[i][used - 1].term : 0u;
            for (int k = 0; k < We; ++k) {
                const size_t o = (size_t)k * ldr + i;
                if (k < used) {
                    colterm[o] = (uint32_t)rows[i][k].col | ((uint32_t)rows[i][k].term << 24);
                    coef[o] = rows[i][k].coef;
                } else {
                    colterm[o] = (uint32_t)i | (pad_term << 24);  // own row: always inside the gather window
                }
            }
            if (i < n) cnt[i] = (unsigned char)std::max(used, 1);
        }
        for (int b0 = 0; b0 < ldr; b0 += 32) {  // the kernel wants the max over the aligned 32-row block
            unsigned char m = 1;
            for (int i = b0; i < b0 + 32; ++i) m = std::max(m, cnt[i]);
            for (int i = b0; i < b0 + 32; ++i) cnt[i] = m;
        }
        std::vector<double> srhs(n);
        for (int i = 0; i < n; ++i) srhs[i] = rhs[perm[i]];
        h->s_ldr = ldr;
        h->s_We = We;
        h->s_bandwidth = bandwidth;
        h->h_sinv = inv;
        if (int e = h->d_scolterm.upload(colterm, h->stream)) return e;
        if (int e = h->d_scoef.upload(coef, h->stream)) return e;
        if (int e = h->d_scnt.upload(cnt, h->stream)) return e;
        if (int e = h->d_srhs.upload(srhs, h->stream)) return e;
        if (int e = h->d_sperm.upload(perm, h->stream)) return e;
    }
    TFIN_CUDA(cudaStreamSynchronize(h->stream));
    // a new operator invalidates everything that was indexed by the previous one
    h->n_cells = 0;
    h->n_obs = 0;
    h->n_avg = 0;
    h->b_nr = 0;
    h->n_r = 0;
    h->rg_ob = 0;
    h->f_n = 0;
    h->h_obs_ptr.clear();
    h->fr_nod.release();
    h->fr_nod.why = "tfin_set_cells not called";
    // sparse-direct solver: symbolic analysis of the shared pattern, affine terms as the assembly list
    h->fr_aff.release();
    {
        const int32_t nnz_ = nnz;
        auto terms = [&](int e, std::vector<FrontalTermEntry>& out) {
            for (int t = 0; t < n_terms; ++t) {
                const double v = vals[(size_t)t * nnz_ + e];
                if (v != 0.0) out.push_back(FrontalTermEntry{t, v});
            }
        };
        h->fr_aff.why = frontal_build(n, row_ptr, col_idx, rhs, terms, &h->fr_aff.host);
        if (h->fr_aff.why.empty()) h->fr_aff.why = frontal_build(n, row_ptr, col_idx, rhs, terms, &h->fr_aff.host1, false);
        if (h->fr_aff.why.empty()) {
            h->fr_aff.ncv = n_terms;
            if (int e = h->fr_aff.upload(h->stream, 0, h->frontal_cfg())) return e;
            TFIN_CUDA(cudaStreamSynchronize(h->stream));
            h->fr_aff.ok = true;
        }
    }
    return 0;
}

static int upload_csr(tfin_ctx* h, const char* who, int rows, const int32_t* ptr, const int32_t* idx,
                      const double* val, DevBuf<int>& dptr, DevBuf<int>& didx, DevBuf<double>& dval) {
    if (h->n <= 0) return fail(TFIN_E_STATE, "%s: call tfin_set_operator first", who);
    if (rows <= 0 || !ptr || !idx || !val) return fail(TFIN_E_ARG, "%s: bad argument", who);
    if (ptr[0] != 0) return fail(TFIN_E_ARG, "%s: ptr[0] must be 0", who);
    const int nnz = ptr[rows];
    for (int r = 0; r < rows; ++r)
        if (ptr[r + 1] < ptr[r]) return fail(TFIN_E_ARG, "%s: malformed ptr", who);
    for (int j = 0; j < nnz; ++j)
        if (idx[j] < 0 || idx[j] >= h->n) return fail(TFIN_E_ARG, "%s: index out of range", who);
    std::vector<int> p(ptr, ptr + rows + 1), ix(idx, idx + nnz);
    std::vector<double> v(val, val + nnz);
    if (int e = dptr.upload(p, h->stream)) return e;
    if (int e = didx.upload(ix, h->stream)) return e;
    if (int e = dval.upload(v, h->stream)) return e;
    TFIN_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

// Transpose of a (rows x n) CSR operator as CSR over the n dofs.
static int upload_csr_transpose(tfin_ctx* h, int rows, const int32_t* ptr, const int32_t* idx, const double* val,
                                DevBuf<int>& dptr, DevBuf<int>& didx, DevBuf<double>& dval) {
    const int n = h->n, nnz = ptr[rows];
    std::vector<int> tp(n + 1, 0), ti(nnz);
    std::vector<double> tv(nnz);
    for (int j = 0; j < nnz; ++j) tp[idx[j] + 1]++;
    for (int i = 0; i < n; ++i) tp[i + 1] += tp[i];
    std::vector<int> fill(tp.begin(), tp.end() - 1);
    for (int o = 0; o < rows; ++o)
        for (int j = ptr[o]; j < ptr[o + 1]; ++j) {
            const int dst = fill[idx[j]]++;
            ti[dst] = o;
            tv[dst] = val[j];
        }
    if (int e = dptr.upload(tp, h->stream)) return e;
    if (int e = didx.upload(ti, h->stream)) return e;
    if (int e = dval.upload(tv, h->stream)) return e;
    TFIN_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int tfin_set_observation(tfin_handle_t h, int32_t n_obs, const int32_t* ptr, const int32_t* idx,
                                    const double* val) {
    CHECK_HANDLE(h);
    if (int e = upload_csr(h, "tfin_set_observation", n_obs, ptr, idx, val, h->d_obs_ptr, h->d_obs_idx, h->d_obs_val))
        return e;
    h->n_obs = n_obs;
    h->h_obs_ptr.assign(ptr, ptr + n_obs + 1);
    h->h_obs_idx.assign(idx, idx + ptr[n_obs]);
    h->h_obs_val.assign(val, val + ptr[n_obs]);
    for (FrontalSet* fs : {&h->fr_aff, &h->fr_nod})
        if (fs->ok) {
            frontal_set_obs(fs->host, n_obs, ptr, idx, val);
            frontal_set_obs(fs->host1, n_obs, ptr, idx, val);
            if (int e = fs->upload(h->stream, n_obs, h->frontal_cfg())) return e;
            TFIN_CUDA(cudaStreamSynchronize(h->stream));
        }
    // B_obs^T as CSR over the n dofs: right-hand sides of the adjoint solves
    if (int e = upload_csr_transpose(h, n_obs, ptr, idx, val, h->d_obsT_ptr, h->d_obsT_idx, h->d_obsT_val)) return e;
    if (h->stream_ok) {  // the streaming path works in its own row numbering
        std::vector<int> ix(ptr[n_obs]);
        for (int j = 0; j < ptr[n_obs]; ++j) ix[j] = h->h_sinv[idx[j]];
        if (int e = h->d_sobs_idx.upload(ix, h->stream)) return e;
        TFIN_CUDA(cudaStreamSynchronize(h->stream));
    }
    return 0;
}

extern "C" int tfin_set_averaging(tfin_handle_t h, int32_t n_rows, const int32_t* ptr, const int32_t* idx,
                                  const double* val) {
    CHECK_HANDLE(h);
    if (int e = upload_csr(h, "tfin_set_averaging", n_rows, ptr, idx, val, h->d_avg_ptr, h->d_avg_idx, h->d_avg_val))
        return e;
    h->n_avg = n_rows;
    // Avg^T over the dofs: lifts d/d theta to d/d k (dsigma_dk, averaged_affine_ROM.py:210, 350)
    return upload_csr_transpose(h, n_rows, ptr, idx, val, h->d_avgT_ptr, h->d_avgT_idx, h->d_avgT_val);
}

extern "C" int tfin_set_cells(tfin_handle_t h, int32_t n_cells, const int32_t* cells, const double* Ke,
                              int32_t prune_zeros) {
    CHECK_HANDLE(h);
    if (h->n <= 0) return fail(TFIN_E_STATE, "tfin_set_cells: call tfin_set_operator first");
    if (n_cells <= 0 || !cells || !Ke) return fail(TFIN_E_ARG, "tfin_set_cells: bad argument");
    const int n = h->n, ld = h->ld;
    const std::vector<int32_t>& rp = h->h_row_ptr;
    const std::vector<int32_t>& ci = h->h_col_idx;
    const int nnz = (int)ci.size();
    // per CSR entry: list of (cell, coef)
    std::vector<std::vector<std::pair<int, double>>> contrib(nnz);
    auto find = [&](int r, int c) -> int {
        const int32_t* b = ci.data() + rp[r];
        const int32_t* e = ci.data() + rp[r + 1];
        const int32_t* it = std::lower_bound(b, e, c);
        if (it != e && *it == c) return (int)(it - ci.data());
        for (const int32_t* q = b; q != e; ++q)  // unsorted rows
            if (*q == c) return (int)(q - ci.data());
        return -1;
    };
    for (int e = 0; e < n_cells; ++e)
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
                const int r = cells[3 * e + a], c = cells[3 * e + b];
                if (r < 0 || r >= n || c < 0 || c >= n) return fail(TFIN_E_ARG, "tfin_set_cells: vertex out of range");
                const int j = find(r, c);
                if (j < 0) return fail(TFIN_E_ARG, "tfin_set_cells: entry (%d,%d) not in the operator pattern", r, c);
                contrib[j].push_back({e, Ke[9 * (size_t)e + 3 * a + b]});
            }
    // sparse-direct solver: same pattern, assembly list = constant (Robin) part + cell contributions
    h->fr_nod.release();
    {
        auto terms = [&](int e, std::vector<FrontalTermEntry>& out) {
            if (h->h_const[e] != 0.0) out.push_back(FrontalTermEntry{0, h->h_const[e]});
            for (auto& pc : contrib[e])
                if (pc.second != 0.0) out.push_back(FrontalTermEntry{1 + pc.first, pc.second});
        };
        std::vector<double> rhs_host(n);
        TFIN_CUDA(cudaMemcpy(rhs_host.data(), h->d_rhs.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
        h->fr_nod.why = frontal_build(n, rp.data(), ci.data(), rhs_host.data(), terms, &h->fr_nod.host);
        if (h->fr_nod.why.empty())
            h->fr_nod.why = frontal_build(n, rp.data(), ci.data(), rhs_host.data(), terms, &h->fr_nod.host1, false);
        if (h->fr_nod.why.empty()) {
            h->fr_nod.ncv = n_cells + 1;
            if (!h->h_obs_ptr.empty())
                for (FrontalProgram* fp : {&h->fr_nod.host, &h->fr_nod.host1})
                    frontal_set_obs(*fp, h->n_obs, h->h_obs_ptr.data(), h->h_obs_idx.data(), h->h_obs_val.data());
            if (int e = h->fr_nod.upload(h->stream, h->n_obs, h->frontal_cfg())) return e;
            TFIN_CUDA(cudaStreamSynchronize(h->stream));
            h->fr_nod.ok = true;
        }
    }
    {
        std::vector<int> cl(cells, cells + 3 * (size_t)n_cells);
        if (int e = h->d_cells.upload(cl, h->stream)) return e;
        std::vector<double> ke(Ke, Ke + 9 * (size_t)n_cells);
        if (int e = h->d_Ke.upload(ke, h->stream)) return e;
        TFIN_CUDA(cudaStreamSynchronize(h->stream));
    }
    if (!h->small_ok) {   // refined meshes: only the direct solver serves the nodal operator
        if (!h->fr_nod.ok)
            return fail(TFIN_E_STATE, "tfin_set_cells: n = %d is beyond the on-chip PCG (n <= 8191) and the direct solver is unavailable: %s",
                        h->n, h->fr_nod.why.c_str());
        // cells around every dof (those of its diagonal entry): the gradient-form kernel walks them
        std::vector<int> dptr(n + 1, 0), dcell;
        for (int i = 0; i < n; ++i) {
            dptr[i] = (int)dcell.size();
            for (int j = rp[i]; j < rp[i + 1]; ++j)
                if (ci[j] == i)
                    for (auto& pc : contrib[j]) dcell.push_back(pc.first);
        }
        dptr[n] = (int)dcell.size();
        if (int e = h->d_dptr.upload(dptr, h->stream)) return e;
        if (int e = h->d_dcell.upload(dcell, h->stream)) return e;
        TFIN_CUDA(cudaStreamSynchronize(h->stream));
        h->Wn = 0;
        h->n_cells = n_cells;
        return 0;
    }
    std::vector<std::vector<int>> keep(n);
    std::vector<int> diag_pos(n, -1);
    int W = 0;
    for (int i = 0; i < n; ++i) {
        for (int j = rp[i]; j < rp[i + 1]; ++j) {
            if (ci[j] == i) {
                diag_pos[i] = j;
                continue;
            }
            if (contrib[j].size() > 2)
                return fail(TFIN_E_ARG, "tfin_set_cells: edge (%d,%d) belongs to %d cells (non-manifold mesh)", i,
                            ci[j], (int)contrib[j].size());
            bool nz = !prune_zeros || h->h_const[j] != 0.0;
            for (auto& pc : contrib[j]) nz = nz || pc.second != 0.0;
            if (nz) keep[i].push_back(j);
        }
        W = std::max(W, (int)keep[i].size());
    }
    if (W == 0) W = 1;
    align_ell_slots(n, W, ci.data(), keep);
    const size_t plane = (size_t)W * ld;
    std::vector<uint16_t> col(plane);
    std::vector<int> cell(2 * plane, n_cells), dptr(ld + 1, 0), dcell, cl(cells, cells + 3 * (size_t)n_cells);
    std::vector<double> coef(2 * plane, 0.0), cst(plane, 0.0), dcoef, dcst(ld, 1.0);
    for (int w = 0; w < W; ++w)
        for (int i = 0; i < ld; ++i) col[(size_t)w * ld + i] = (uint16_t)std::min(i, n - 1);
    for (int i = 0; i < n; ++i) {
        for (int w = 0; w < W; ++w) {
            const int j = keep[i][w];
            const size_t o = (size_t)w * ld + i;
            if (j < 0) {  // padding: own row, no cells, constant 0
                col[o] = (uint16_t)i;
                continue;
            }
            col[o] = (uint16_t)ci[j];
            cst[o] = h->h_const[j];
            for (size_t c = 0; c < contrib[j].size(); ++c) {
                cell[c * plane + o] = contrib[j][c].first;
                coef[c * plane + o] = contrib[j][c].second;
            }
        }
        dptr[i] = (int)dcell.size();
        dcst[i] = h->h_const[diag_pos[i]];
        for (auto& pc : contrib[diag_pos[i]]) {
            dcell.push_back(pc.first);
            dcoef.push_back(pc.second);
        }
    }
    for (int i = n; i <= ld; ++i) dptr[i] = (int)dcell.size();
    h->Wn = W;
    if (int e = h->d_ncol.upload(col, h->stream)) return e;
    if (int e = h->d_ncell.upload(cell, h->stream)) return e;
    if (int e = h->d_ncoef.upload(coef, h->stream)) return e;
    if (int e = h->d_ncst.upload(cst, h->stream)) return e;
    if (int e = h->d_dptr.upload(dptr, h->stream)) return e;
    if (int e = h->d_dcell.upload(dcell, h->stream)) return e;
    if (int e = h->d_dcoef.upload(dcoef, h->stream)) return e;
    if (int e = h->d_dcst.upload(dcst, h->stream)) return e;
    if (int e = h->d_cells.upload(cl, h->stream)) return e;
    {
        std::vector<double> ke(Ke, Ke + 9 * (size_t)n_cells);
        if (int e = h->d_Ke.upload(ke, h->stream)) return e;
    }
    TFIN_CUDA(cudaStreamSynchronize(h->stream));
    h->n_cells = n_cells;
    return 0;
}

extern "C" int tfin_set_rom(tfin_handle_t h, int32_t n_r, int32_t n_terms, int32_t n_obs, const double* S,
                            const double* G, const double* obs_phi) {
    CHECK_HANDLE(h);
    if (n_r <= 0 || n_r > 127 || !S || !G || !obs_phi || n_obs <= 0)
        return fail(TFIN_E_ARG, "tfin_set_rom: bad argument (n_r must be in [1,127])");
    if (n_terms < 1 || n_terms > TFIN_MAX_TERMS) return fail(TFIN_E_ARG, "tfin_set_rom: bad n_terms");
    const int P2 = n_terms * (n_terms + 1) / 2, T = n_r * (n_r + 1) / 2, Taug = rom_taug(n_r);
    const int ldS = (Taug + ROM_BN - 1) / ROM_BN * ROM_BN;  // whole tiles of the combine sweep, zero padded
    // repack: row-major packed lower (i>=j -> i(i+1)/2+j)  ->  augmented column-major, G in the extra row
    std::vector<double> Saug((size_t)P2 * ldS, 0.0);
    int pq = 0;
    for (int p = 0; p < n_terms; ++p)
        for (int q = p; q < n_terms; ++q, ++pq) {
            double* dst = Saug.data() + (size_t)pq * ldS;
            const double* src = S + (size_t)pq * T;
            for (int j = 0; j < n_r; ++j) {
                const int oj = rom_col_off(j, n_r);
                for (int i = j; i < n_r; ++i) dst[oj + (i - j)] = src[(size_t)i * (i + 1) / 2 + j];
                if (p == 0) dst[oj + (n_r - j)] = G[(size_t)q * n_r + j];  // coefficient th_0 th_q = th_q
            }
        }
    std::vector<double> op(obs_phi, obs_phi + (size_t)n_obs * n_r);
    if (int e = h->d_S.upload(Saug, h->stream)) return e;
    if (int e = h->d_obs_phi.upload(op, h->stream)) return e;
    TFIN_CUDA(cudaStreamSynchronize(h->stream));
    h->n_r = n_r;
    h->rom_terms = n_terms;
    h->rom_obs = n_obs;
    h->rg_ob = 0;  // a new basis invalidates the gradient tensors
    return 0;
}

extern "C" int tfin_set_rom_gradient(tfin_handle_t h, int32_t n_r, int32_t n_terms, const double* gram) {
    CHECK_HANDLE(h);
    if (h->n_r <= 0) return fail(TFIN_E_STATE, "tfin_set_rom_gradient: call tfin_set_rom first");
    if (!gram || n_r != h->n_r || n_terms != h->rom_terms)
        return fail(TFIN_E_ARG, "tfin_set_rom_gradient: n_r / n_terms must match tfin_set_rom (%d, %d)", h->n_r,
                    h->rom_terms);
    // gram[t][q-1][i][j] = (Psi_t^T Psi_q)[i][j]  ->  NG[ob][i*jp + j][96] with output o = t (n_terms-1) + (q-1); the j range
    // is padded to jp = multiple of 4 with zero rows (a DMMA k-step of 4 never straddles two i)
    const int n_par = n_terms - 1, O = n_terms * n_par, jp = rom_grad_jpad(n_r), K = n_r * jp, n_ob = (O + RG_OB - 1) / RG_OB;
    std::vector<double> ng((size_t)n_ob * K * RG_OB, 0.0);
    for (int o = 0; o < O; ++o) {
        const double* src = gram + (size_t)o * n_r * n_r;
        double* dst = ng.data() + (size_t)(o / RG_OB) * K * RG_OB + (o % RG_OB);
        for (int i = 0; i < n_r; ++i)
            for (int j = 0; j < n_r; ++j) dst[(size_t)(i * jp + j) * RG_OB] = src[i * n_r + j];
    }
    if (int e = h->d_NG.upload(ng, h->stream)) return e;
    TFIN_CUDA(cudaStreamSynchronize(h->stream));
    h->rg_ob = n_ob;
    return 0;
}

// ------------------------------------------------------------------------------------------------ launch helpers
struct PcgGeom {
    const PcgVariant* v;
    int T, occ;
    size_t smem;
};

static bool pcg_geom_for(tfin_ctx* h, const PcgVariant* v, int n_cells, PcgGeom* g, bool adjoint = false) {
    const int n = h->n, R = v->R;
    const void* func = adjoint ? v->func_adj : v->func;
    if (!func) return false;
    const int T = ((n + R - 1) / R + 31) & ~31;
    if (T > v->maxT || T > 1024 || R * T > 8192) return false;
    const PcgSmem L = PcgSmem::make(v->WT - v->WR, R * T, n_cells, n);
    if (L.total > (size_t)h->max_smem_optin) return false;
    if (cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, func, T, L.total) != cudaSuccess || occ < 1) {
        cudaGetLastError();
        return false;
    }
    g->v = v;
    g->T = T;
    g->occ = occ;
    g->smem = L.total;
    return true;
}

static int launch_pcg(tfin_ctx* h, bool nodal, const double* d_in, int in_stride, int64_t N, double tol, int maxit,
                      double* d_w, double* d_qoi, int* d_iters, int* d_status, double* d_relres, cudaStream_t st,
                      const PcgAdj* adjoint = nullptr) {
    const int W = nodal ? h->Wn : h->W;
    const int nc = nodal ? h->n_cells : 0;
    // candidates: variants with the smallest compiled ELL width >= W; user knobs filter further
    int wt_min = 1 << 30;
    for (int g = 0; g < PCG_NUM_GROUPS; ++g) {
        int cnt = 0;
        const PcgVariant* tab = pcg_variants(g, nodal, &cnt);
        for (int i = 0; i < cnt; ++i)
            if (tab[i].WT >= W) wt_min = std::min(wt_min, tab[i].WT);
    }
    PcgGeom best{};
    double best_score = -1.0;
    for (int g = 0; g < PCG_NUM_GROUPS; ++g) {
        int cnt = 0;
        const PcgVariant* tab = pcg_variants(g, nodal, &cnt);
        for (int i = 0; i < cnt; ++i) {
            const PcgVariant* v = &tab[i];
            if (v->WT < W) continue;
            if (h->pcg_R > 0 && v->R != h->pcg_R) continue;
            if (h->pcg_WR >= 0 && (v->WR > 0) != (h->pcg_WR > 0)) continue;
            PcgGeom geo{};
            if (!pcg_geom_for(h, v, nc, &geo, adjoint != nullptr)) continue;
            // score: resident rows per SM that do useful work, two CTAs per SM preferred (barrier latency of one
            // CTA overlaps the SpMV of the other), narrower compiled width preferred
            const double pad = (double)h->n / (v->R * geo.T);
            double score = pad * (geo.occ >= 2 ? 1.5 : 1.0) * (v->WT == wt_min ? 1.0 : 0.5 * wt_min / v->WT);
            if (v->R == 5) score *= 1.02;
            if (score > best_score) {
                best_score = score;
                best = geo;
            }
        }
    }
    if (best_score < 0)
        return fail(TFIN_E_STATE, "on-chip PCG: no compiled variant fits (n=%d, ell width=%d, rows/thread=%d)", h->n, W,
                    h->pcg_R);
    const int grid = (int)std::min<int64_t>(N, (int64_t)h->sm_count * best.occ);
    h->last_R = best.v->R;
    h->last_T = best.T;
    h->last_occ = best.occ;
    h->last_smem = best.smem;
    h->last_WT = best.v->WT;
    h->last_WR = best.v->WR;
    h->last_path = 1;

    TFIN_CUDA(cudaMemsetAsync(h->d_counter.p, 0, sizeof(unsigned long long), st));
    CsrRows obs{h->n_obs, h->d_obs_ptr.p, h->d_obs_idx.p, h->d_obs_val.p};
    PcgIO io{d_in, (long long)N, in_stride, tol * tol, maxit, d_w, d_qoi, d_iters, d_status, d_relres,
             h->d_counter.p};
    PcgOp op{};
    op.n = h->n;
    op.ld = h->ld;
    op.W = W;
    op.n_terms = h->n_terms;
    op.n_cells = nc;
    op.rhs = h->d_rhs.p;
    if (nodal) {
        op.col = h->d_ncol.p;
        op.cell = h->d_ncell.p;
        op.coef = h->d_ncoef.p;
        op.cst = h->d_ncst.p;
        op.dptr = h->d_dptr.p;
        op.dcell = h->d_dcell.p;
        op.dcoef = h->d_dcoef.p;
        op.dcst = h->d_dcst.p;
        op.cells = h->d_cells.p;
        op.coef_mode = h->coef_mode;
    } else {
        op.col = h->d_col.p;
        op.val = h->d_val.p;
        op.diag = h->d_diag.p;
    }
    PcgAdj adj{};
    if (adjoint) adj = *adjoint;
    void* args[] = {&op, &obs, &io, &adj};
    TFIN_CUDA(cudaLaunchKernel(adjoint ? best.v->func_adj : best.v->func, dim3(grid), dim3(best.T), args, best.smem, st));
    h->launches += 1;
    return 0;
}

// Optional fp32 variant of the on-chip affine PCG (K1f, pcg_f32.cuh).
static int launch_pcg_f32(tfin_ctx* h, const double* d_in, int in_stride, int64_t N, double tol, int maxit, double* d_w,
                          double* d_qoi, int* d_iters, int* d_status, double* d_relres, cudaStream_t st) {
    int cnt = 0;
    const PcgF32Variant* tab = pcg_f32_variants(&cnt);
    const PcgF32Variant* best = nullptr;
    int bestT = 0, best_occ = 0;
    size_t best_smem = 0;
    double best_score = -1.0;
    for (int i = 0; i < cnt; ++i) {
        const PcgF32Variant* v = &tab[i];
        if (v->WT < h->W) continue;
        const int T = ((h->n + v->R - 1) / v->R + 31) & ~31;
        if (T > v->maxT) continue;
        const size_t sm = PcgF32Smem::make(v->WT, v->R * T).total;
        if (sm > (size_t)h->max_smem_optin) continue;
        if (cudaFuncSetAttribute(v->func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, v->func, T, sm) != cudaSuccess || occ < 1) {
            cudaGetLastError();
            continue;
        }
        const double score = (double)h->n / (v->R * T) * std::min(occ, 3) / v->WT;  // useful rows, residency, padding
        if (score > best_score) {
            best_score = score;
            best = v;
            bestT = T;
            best_occ = occ;
            best_smem = sm;
        }
    }
    if (!best)
        return fail(TFIN_E_STATE, "fp32 PCG: no compiled variant fits (n=%d, ell width=%d); use pcg_precision = 64", h->n, h->W);
    const int grid = (int)std::min<int64_t>(N, (int64_t)h->sm_count * best_occ);
    h->last_R = best->R;
    h->last_T = bestT;
    h->last_occ = best_occ;
    h->last_smem = best_smem;
    h->last_WT = best->WT;
    h->last_WR = 0;
    h->last_path = 3;
    TFIN_CUDA(cudaMemsetAsync(h->d_counter.p, 0, sizeof(unsigned long long), st));
    CsrRows obs{h->n_obs, h->d_obs_ptr.p, h->d_obs_idx.p, h->d_obs_val.p};
    PcgIO io{d_in, (long long)N, in_stride, tol * tol, maxit, d_w, d_qoi, d_iters, d_status, d_relres, h->d_counter.p};
    PcgOp op{};
    op.n = h->n;
    op.ld = h->ld;
    op.W = h->W;
    op.n_terms = h->n_terms;
    op.rhs = h->d_rhs.p;
    op.col = h->d_col.p;
    op.val = h->d_val.p;
    op.diag = h->d_diag.p;
    void* args[] = {&op, &obs, &io};
    TFIN_CUDA(cudaLaunchKernel(best->func, dim3(grid), dim3(bestT), args, best_smem, st));
    h->launches += 1;
    return 0;
}

static int launch_pcg_stream(tfin_ctx* h, const double* d_in, int in_stride, int64_t N, double tol, int maxit,
                             double* d_w, double* d_qoi, int* d_iters, int* d_status, double* d_relres,
                             cudaStream_t st) {
    if (!h->stream_ok) return fail(TFIN_E_STATE, "streaming PCG supports n < 2^24 rows (n = %d)", h->n);
    int S = h->stream_tile;
    if (S == 0) S = N <= (int64_t)4 * h->sm_count ? 4 : 8;
    if (S != 4 && S != 8 && S != 16 && S != 32) return fail(TFIN_E_ARG, "stream_tile must be 4, 8, 16 or 32");
    const int64_t n_tiles = (N + S - 1) / S;
    const int grid = (int)std::min<int64_t>(n_tiles, h->sm_count);
    if (int e = h->d_swork.reserve((size_t)grid * 5 * h->s_ldr * S)) return e;
    TFIN_CUDA(cudaMemsetAsync(h->d_counter.p, 0, sizeof(unsigned long long), st));
    // shared-memory ring for the gathered vector: chunk = 16 warps x 32/(S/4) rows, 8 chunks; needs
    // 2 * halo + look-ahead + 2 <= 8 chunks
    const int chunk_rows = 16 * (32 / (S / 4));
    const int hb = (h->s_bandwidth + chunk_rows - 1) / chunk_rows, la = STREAM_RING_SLOTS - 2 * hb - 2;
    const bool ring_possible = S == 8 && la >= 1;
    if (h->stream_ring == 1 && !ring_possible)
        return fail(TFIN_E_STATE, "stream_ring = 1 needs tile 8 and bandwidth <= %d rows (bandwidth %d, tile %d)",
                    2 * chunk_rows, h->s_bandwidth, S);
    // measured (tools/gpu_probe_stream.py, n = 99 945): the 128 KB ring shrinks L1 so much that the streaming pass B
    // loses more than pass A gains (0.73 vs 0.85 of the HBM peak), so the ring is opt-in only
    const bool ring = ring_possible && h->stream_ring == 1;
    // per-warp double buffer for the operator slice of a row group: Ws slots x rows-per-warp x (4 + 8) bytes
    const int rpw = 32 / (S / 4);
    const size_t smem_avail = (size_t)h->max_smem_optin - 4096 - (ring ? STREAM_RING_BYTES : 0);
    const int Ws = (int)std::min<size_t>(h->s_We, smem_avail / ((size_t)16 * 2 * rpw * 12));
    if (Ws < 1) return fail(TFIN_E_STATE, "streaming PCG: no shared memory left for the operator stage");
    StreamOp op{h->n, h->n_terms, h->s_ldr, h->s_We, h->d_scolterm.p, h->d_scoef.p, h->d_scnt.p, h->d_srhs.p,
                h->d_sperm.p, hb, la, Ws};
    unsigned long long* prof = nullptr;
    if (h->stream_prof) {
        if (int e = h->d_sprof.reserve(4)) return e;
        TFIN_CUDA(cudaMemsetAsync(h->d_sprof.p, 0, 4 * sizeof(unsigned long long), st));
        prof = h->d_sprof.p;
    }
    CsrRows obs{h->n_obs, h->d_obs_ptr.p, h->d_sobs_idx.p, h->d_obs_val.p};
    PcgIO io{d_in, (long long)N, in_stride, tol * tol, maxit, d_w, d_qoi, d_iters, d_status, d_relres,
             h->d_counter.p};
    void (*kern)(StreamOp, CsrRows, PcgIO, double*, unsigned long long*) =
        S == 4 ? pcg_stream_kernel<4, false>
        : S == 8 ? (ring ? pcg_stream_kernel<8, true> : pcg_stream_kernel<8, false>)
        : S == 16 ? pcg_stream_kernel<16, false> : pcg_stream_kernel<32, false>;
    const size_t dyn = (ring ? STREAM_RING_BYTES : 0) + (size_t)16 * 2 * Ws * rpw * 12 + (size_t)h->stream_pad_smem * 1024;
    if (dyn) TFIN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    kern<<<grid, 512, dyn, st>>>(op, obs, io, h->d_swork.p, prof);
    h->last_ring = ring ? 1 : 0;
    TFIN_CUDA(cudaGetLastError());
    h->launches += 1;
    h->last_path = 2;
    h->last_tile = S;
    return 0;
}

// ---- sparse-direct solver (frontal.cuh).  Geometry of the kernel that would serve this operator; kernel = 0 if none.
static const void* frontal_lane_fn(int cmax, int phase) {
#define TFIN_LANE_FN(PH)                                                                                             \
    (cmax <= 8 ? (const void*)frontal_lane_kernel<8, PH> : cmax <= 16 ? (const void*)frontal_lane_kernel<16, PH>     \
     : cmax <= 24 ? (const void*)frontal_lane_kernel<24, PH> : (const void*)frontal_lane_kernel<32, PH>)
    return phase == FRONTAL_PHASE_FACTOR ? TFIN_LANE_FN(FRONTAL_PHASE_FACTOR)
           : phase == FRONTAL_PHASE_BSUB ? TFIN_LANE_FN(FRONTAL_PHASE_BSUB)
           : phase == FRONTAL_PHASE_FSUB ? TFIN_LANE_FN(FRONTAL_PHASE_FSUB) : TFIN_LANE_FN(FRONTAL_PHASE_BOTH);
#undef TFIN_LANE_FN
}

static const void* frontal_cta_fn(int mode, int cmax) {
#define TFIN_CTA_FN(M)                                                                                        \
    (cmax <= 32 ? (const void*)frontal_cta_kernel<M, 1> : cmax <= 64 ? (const void*)frontal_cta_kernel<M, 2>   \
     : cmax <= 128 ? (const void*)frontal_cta_kernel<M, 4> : cmax <= 256 ? (const void*)frontal_cta_kernel<M, 8> : nullptr)
    return mode == FRONTAL_MODE_QOI ? TFIN_CTA_FN(FRONTAL_MODE_QOI) : TFIN_CTA_FN(FRONTAL_MODE_SOLVE);
#undef TFIN_CTA_FN
}

struct FrontalGeom {
    int kernel = 0;   // 1 = D1 sample per thread, 2 = D2 sample per CTA
    int mode = 0;     // D2: FRONTAL_MODE_QOI / FRONTAL_MODE_SOLVE
    int threads = 0, occ = 0;
    size_t smem = 0;
    bool split = false;            // D1: factorisation and substitution as two kernels
    int occ_b = 0;                 // D1 split: resident warps per SM of the substitution kernel
    size_t smem_b = 0;
    FrontalCtaSmem cta{};
};

static FrontalGeom frontal_geom(tfin_ctx* h, bool nodal, bool want_w) {
    FrontalGeom g;
    const FrontalSet& fs = nodal ? h->fr_nod : h->fr_aff;
    if (!fs.ok) return g;
    const FrontalProgram& P = fs.host;
    const int ntri = P.nslots * (P.nslots + 1) / 2;
    const int ncv_smem = nodal ? 0 : fs.ncv;
    const int ring = fs.streams.ring_bytes;
    if (h->frontal_kernel != 2 && P.cmax <= 32 && fs.streams.lr_rows > 0) {
        auto fits = [&](int phase, size_t* sm_out, int* occ_out) {
            const void* fn = frontal_lane_fn(P.cmax, phase);
            const int ns1 = fs.host1.nslots;
            const size_t sm = frontal_lane_smem(ns1 * (ns1 + 1) / 2, ns1, h->n_obs, ncv_smem, fs.streams.lr_rows, fs.streams.lanes,
                                                phase == FRONTAL_PHASE_FACTOR ? fs.streams.ring_fwd1 : ring, phase);
            if (sm > (size_t)h->max_smem_optin ||
                cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess)
                return false;
            int occ = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 32, sm) != cudaSuccess || occ < 1) return false;
            *sm_out = sm;
            *occ_out = occ;
            return true;
        };
        if (h->frontal_split != 0 && fits(FRONTAL_PHASE_FACTOR, &g.smem, &g.occ) && fits(FRONTAL_PHASE_BSUB, &g.smem_b, &g.occ_b)) {
            g.kernel = 1;
            g.threads = 32;
            g.split = true;
            return g;
        }
        if (fits(FRONTAL_PHASE_BOTH, &g.smem, &g.occ)) {
            g.kernel = 1;
            g.threads = 32;
            return g;
        }
        cudaGetLastError();
    }
    if (h->frontal_kernel == 1) return g;
    // D2: observables-only mode unless the solution is wanted (or there are too many observation rows to carry along)
    int mode = (want_w || h->n_obs > 16) ? FRONTAL_MODE_SOLVE : FRONTAL_MODE_QOI;
    if (h->frontal_mode == 0 && !want_w) mode = FRONTAL_MODE_QOI;
    if (h->frontal_mode == 1) mode = FRONTAL_MODE_SOLVE;
    const int R = mode == FRONTAL_MODE_QOI ? 1 + h->n_obs : 1;
    const FrontalCtaSmem L = FrontalCtaSmem::make(ntri, P.nslots, std::max(P.cmax, 1), R, std::max(h->n_obs, 1), ncv_smem, ring);
    if (L.total + 64 > (size_t)h->max_smem_optin) return g;
    int T = h->frontal_threads;
    if (T <= 0) {
        const double pairs = 0.5 * P.cmax * (P.cmax + 1.0);
        T = 32 * (int)std::min(16.0, std::max(1.0, std::ceil(pairs / (32.0 * 16.0))));
    }
    T = std::max(32, std::min(512, (T + 31) & ~31));   // __launch_bounds__(512): 128 registers per thread
    const void* fn = frontal_cta_fn(mode, P.cmax);
    if (!fn) return g;   // columns of more than 256 entries: PCG
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total) != cudaSuccess) {
        cudaGetLastError();
        return g;
    }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, T, L.total) != cudaSuccess || occ < 1) {
        cudaGetLastError();
        return g;
    }
    g.kernel = 2;
    g.mode = mode;
    g.threads = T;
    g.occ = occ;
    g.smem = L.total;
    g.cta = L;
    return g;
}

// d_in: theta (N, in_stride) for the affine operator, nodal fields k (N, n) for the nodal one.
// Adjoint right-hand side of a D2 solve (FrontalIO::adj): pointers of sample 0 of the batch.
struct FrontalAdjArgs {
    int mode;                // 1 = -B_obs^T (qoi_in - data), 2 = -B_obs[unit_row]^T
    const double* qoi_in;    // (N, n_obs)
    const double* data;      // (1 | N, n_obs)
    int64_t data_stride;
    double* cost_out;        // (N) | null
    int unit_row;
};

static int launch_frontal(tfin_ctx* h, bool nodal, const FrontalGeom& g, const double* d_in, int in_stride, int64_t N,
                          double* d_w, double* d_qoi, int* d_iters, int* d_status, double* d_relres, cudaStream_t st,
                          const FrontalAdjArgs* adj = nullptr) {
    FrontalSet& fs = nodal ? h->fr_nod : h->fr_aff;
    if (!fs.ok || g.kernel == 0)
        return fail(TFIN_E_STATE, "direct solver unavailable for this operator: %s", fs.ok ? "front does not fit shared memory" : fs.why.c_str());
    const FrontalProgram& P = fs.host;
    const FrontalDev dev = fs.dev(g.kernel == 1);
    const size_t per_sample_work = (size_t)P.nnzL + 2 * (size_t)P.n;
    // the nodal operator needs the coefficient vector of every sample in HBM: bound that workspace by chunking
    int64_t chunk = N;
    if (g.kernel == 1 && g.split) {
        // split launch: the factor blocks of every group of a chunk stay in HBM (<= ~6 GiB).  A chunk is a whole number of
        // WAVES of the factor kernel (resident warps x samples per warp): a partial last wave would idle most of the GPU
        // for a full pass (2 ms at n = 1597)
        const int64_t wave = (int64_t)h->sm_count * g.occ * fs.streams.lanes;
        const int64_t mem_waves = std::max<int64_t>(1, ((int64_t)FRONTAL_WORK_BYTES / ((int64_t)(per_sample_work + 2) * 8)) / wave);
        // ... and the substitution kernel runs ceil(k occ / occ_b) passes of its own on a chunk of k waves (a pass of it
        // takes ~0.55 of a factor pass, however empty): take the k <= mem_waves with the most samples per unit time
        int64_t k = 1;
        double best = 0.0;
        for (int64_t kk = 1; kk <= std::min<int64_t>(mem_waves, 64); ++kk) {
            const double passes_b = (double)((kk * g.occ + g.occ_b - 1) / g.occ_b);
            const double eff = (double)kk / ((double)kk + 0.55 * passes_b);
            if (eff > best * 1.01) {
                best = eff;
                k = kk;
            }
        }
        chunk = std::min<int64_t>(chunk, k * wave);
    }
    if (nodal) {
        const int64_t cap = std::max<int64_t>(32, ((int64_t)1 << 30) / ((int64_t)fs.ncv * 8) / 32 * 32);
        chunk = std::min<int64_t>(chunk, cap);
        if (int e = h->d_fcv.reserve((size_t)(chunk + 96) * fs.ncv)) return e;   // whole 32-sample tiles, whole groups
    }
    for (int64_t s0 = 0; s0 < N; s0 += chunk) {
        const int64_t m = std::min<int64_t>(chunk, N - s0);
        FrontalIO io{};
        io.in = d_in + (size_t)s0 * in_stride;
        io.N = m;
        io.in_stride = in_stride;
        io.n_obs = h->n_obs;
        io.w_out = d_w ? d_w + (size_t)s0 * P.n : nullptr;
        io.qoi_out = d_qoi ? d_qoi + (size_t)s0 * h->n_obs : nullptr;
        io.iters_out = d_iters ? d_iters + s0 : nullptr;
        io.status_out = d_status ? d_status + s0 : nullptr;
        io.relres_out = d_relres ? d_relres + s0 : nullptr;
        io.counter = h->d_counter.p;
        if (adj) {
            if (g.kernel != 2 || g.mode != FRONTAL_MODE_SOLVE) return fail(TFIN_E_STATE, "adjoint right-hand side: sample-per-CTA solve mode only");
            io.adj = adj->mode;
            io.qoi_in = adj->qoi_in ? adj->qoi_in + (size_t)s0 * h->n_obs : nullptr;
            io.data = adj->data ? adj->data + (size_t)s0 * adj->data_stride : nullptr;
            io.data_stride = adj->data_stride;
            io.cost_out = adj->cost_out ? adj->cost_out + s0 : nullptr;
            io.unit_row = adj->unit_row;
        }
        if (nodal) {
            const dim3 cgrid((unsigned)((h->n_cells + 31) / 32), (unsigned)((m + 31) / 32));
            frontal_cellcoef_kernel<<<cgrid, dim3(32, 8), 0, st>>>(io.in, (long long)m, P.n, h->n_cells, h->d_cells.p,
                                                                  h->coef_mode, g.kernel == 1 ? fs.streams.lanes : 0, h->d_fcv.p);
            h->launches += 1;
            io.cv_global = h->d_fcv.p;
        }
        TFIN_CUDA(cudaMemsetAsync(h->d_counter.p, 0, sizeof(unsigned long long), st));
        if (g.kernel == 1) {
            const int lanes = fs.streams.lanes;
            const int64_t groups = (m + lanes - 1) / lanes;
            const int grid = (int)std::min<int64_t>(groups, (int64_t)h->sm_count * g.occ);
            void* args[] = {(void*)&dev, (void*)&io};
            if (g.split) {
                if (int e = h->d_fwork.reserve((size_t)groups * (per_sample_work + 2) * lanes)) return e;
                io.work = h->d_fwork.p;
                TFIN_CUDA(cudaLaunchKernel(frontal_lane_fn(P.cmax, FRONTAL_PHASE_FACTOR), dim3(grid), dim3(32), args, g.smem, st));
                TFIN_CUDA(cudaMemsetAsync(h->d_counter.p, 0, sizeof(unsigned long long), st));
                const int grid_b = (int)std::min<int64_t>(groups, (int64_t)h->sm_count * g.occ_b);
                TFIN_CUDA(cudaLaunchKernel(frontal_lane_fn(P.cmax, FRONTAL_PHASE_BSUB), dim3(grid_b), dim3(32), args, g.smem_b, st));
                h->launches += 1;
            } else {
                if (int e = h->d_fwork.reserve((size_t)grid * per_sample_work * lanes)) return e;
                io.work = h->d_fwork.p;
                TFIN_CUDA(cudaLaunchKernel(frontal_lane_fn(P.cmax, FRONTAL_PHASE_BOTH), dim3(grid), dim3(32), args, g.smem, st));
            }
        } else {
            const int grid = (int)std::min<int64_t>(m, (int64_t)h->sm_count * g.occ);
            if (g.mode == FRONTAL_MODE_SOLVE) {
                if (int e = h->d_fwork.reserve((size_t)grid * per_sample_work)) return e;
                io.work = h->d_fwork.p;
            }
            FrontalCtaSmem cta = g.cta;
            void* args[] = {(void*)&dev, (void*)&io, (void*)&cta};
            TFIN_CUDA(cudaLaunchKernel(frontal_cta_fn(g.mode, P.cmax), dim3(grid), dim3(g.threads), args, g.smem, st));
        }
        TFIN_CUDA(cudaGetLastError());
        h->launches += 1;
    }
    h->last_solver = 2;
    h->last_fkernel = g.kernel == 1 ? 1 : (g.mode == FRONTAL_MODE_QOI ? 2 : 3);
    h->last_fthreads = g.threads;
    h->last_focc = g.occ;
    h->last_focc_b = g.split ? g.occ_b : 0;
    h->last_fsmem = g.smem;
    h->last_path = 4;
    return 0;
}

// Fin.gradient (fom/forward_solve.py:293-322) with the direct solver: factorise once, then three substitution passes with
// the stored factor -- backward (w and the observables), forward with the adjoint right-hand side -B_obs^T (qoi - data),
// backward again (adjoint state v) -- and the gradient form.  Device pointers; d_data (1 | N, n_obs).
// sens != 0: Fin.sensitivity (:324-342) instead -- n_obs adjoint solves A v_o = -B_obs[o, :]^T, d_grad is the (N, n_obs, n) Jacobian.
static int launch_frontal_gradient(tfin_ctx* h, const FrontalGeom& g, const double* d_k, int64_t N, const double* d_data,
                                   int64_t data_stride, double* d_grad, double* d_cost, double* d_qoi, int* d_status,
                                   cudaStream_t st, bool sens = false) {
    FrontalSet& fs = h->fr_nod;
    const FrontalProgram& P = fs.host;
    const FrontalDev dev = fs.dev(true);
    const int lanes = fs.streams.lanes, n = P.n, nobs = h->n_obs;
    const size_t per_sample_work = (size_t)P.nnzL + 2 * (size_t)n;
    const int64_t wave = (int64_t)h->sm_count * g.occ * lanes;
    const int64_t mem_waves = std::max<int64_t>(1, ((int64_t)FRONTAL_WORK_BYTES / ((int64_t)(per_sample_work + 2) * 8)) / wave);
    int64_t kk = 1;   // waves per chunk: as launch_frontal, with three more substitution passes per chunk
    {
        double best = 0.0;
        for (int64_t c = 1; c <= std::min<int64_t>(mem_waves, 64); ++c) {
            const double passes_b = (double)((c * g.occ + g.occ_b - 1) / g.occ_b);
            const double eff = (double)c / ((double)c + 0.55 * (sens ? 1 + 2 * h->n_obs : 3) * passes_b);
            if (eff > best * 1.01) {
                best = eff;
                kk = c;
            }
        }
    }
    const int64_t cvcap = std::max<int64_t>(32, ((int64_t)1 << 30) / ((int64_t)fs.ncv * 8) / 32 * 32);
    const int64_t chunk = std::min<int64_t>(N, std::min<int64_t>(kk * wave, cvcap));
    if (int e = h->d_fcv.reserve((size_t)(chunk + 96) * fs.ncv)) return e;
    if (int e = h->d_fw.reserve((size_t)chunk * n)) return e;
    if (int e = h->d_fv.reserve((size_t)chunk * n)) return e;
    if (!d_qoi) {   // the adjoint right-hand side needs the observables
        if (int e = h->d_qoi.reserve((size_t)N * nobs)) return e;
        d_qoi = h->d_qoi.p;
    }
    const void* fn_f = frontal_lane_fn(P.cmax, FRONTAL_PHASE_FACTOR);
    const void* fn_b = frontal_lane_fn(P.cmax, FRONTAL_PHASE_BSUB);
    const void* fn_s = frontal_lane_fn(P.cmax, FRONTAL_PHASE_FSUB);
    TFIN_CUDA(cudaFuncSetAttribute(fn_s, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_b));
    for (int64_t s0 = 0; s0 < N; s0 += chunk) {
        const int64_t m = std::min<int64_t>(chunk, N - s0);
        const int64_t groups = (m + lanes - 1) / lanes;
        if (int e = h->d_fwork.reserve((size_t)groups * (per_sample_work + 2) * lanes)) return e;
        FrontalIO io{};
        io.in = d_k + (size_t)s0 * n;
        io.N = m;
        io.in_stride = n;
        io.n_obs = nobs;
        io.counter = h->d_counter.p;
        io.work = h->d_fwork.p;
        io.cv_global = h->d_fcv.p;
        io.data = d_data + (data_stride ? (size_t)s0 * data_stride : 0);
        io.data_stride = data_stride;
        const dim3 cgrid((unsigned)((h->n_cells + 31) / 32), (unsigned)((m + 31) / 32));
        frontal_cellcoef_kernel<<<cgrid, dim3(32, 8), 0, st>>>(io.in, (long long)m, n, h->n_cells, h->d_cells.p, h->coef_mode, lanes,
                                                              h->d_fcv.p);
        const int grid_f = (int)std::min<int64_t>(groups, (int64_t)h->sm_count * g.occ);
        const int grid_b = (int)std::min<int64_t>(groups, (int64_t)h->sm_count * g.occ_b);
        void* args[] = {(void*)&dev, (void*)&io};
        auto run = [&](const void* fn, int grid, size_t smem) -> int {
            TFIN_CUDA(cudaMemsetAsync(h->d_counter.p, 0, sizeof(unsigned long long), st));
            TFIN_CUDA(cudaLaunchKernel(fn, dim3(grid), dim3(32), args, smem, st));
            return 0;
        };
        if (int e = run(fn_f, grid_f, g.smem)) return e;                       // A = L L^T, y = L^-1 b
        io.w_out = h->d_fw.p;
        io.qoi_out = d_qoi + (size_t)s0 * nobs;
        io.status_out = d_status ? d_status + s0 : nullptr;
        if (int e = run(fn_b, grid_b, g.smem_b)) return e;                     // w = L^-T y, qoi = B_obs w
        const int64_t total = m * (int64_t)n;
        const int gb = (int)std::min<int64_t>((total + 255) / 256, (int64_t)h->sm_count * 16);
        double* const qoi_chunk = io.qoi_out;
        const int n_adj = sens ? nobs : 1;
        for (int o = 0; o < n_adj; ++o) {
            io.w_out = nullptr;
            io.qoi_out = qoi_chunk;             // read by the gradient's right-hand side
            io.status_out = nullptr;
            io.data = sens ? nullptr : d_data + (data_stride ? (size_t)s0 * data_stride : 0);
            io.unit_row = o;
            io.cost_out = (!sens && d_cost) ? d_cost + s0 : nullptr;
            if (int e = run(fn_s, grid_b, g.smem_b)) return e;                 // y' = L^-1 (-B_obs^T (qoi - data))  |  L^-1 (-B_obs[o]^T)
            io.w_out = h->d_fv.p;
            io.qoi_out = nullptr;
            io.cost_out = nullptr;
            if (int e = run(fn_b, grid_b, g.smem_b)) return e;                 // v = L^-T y'
            double* gout = sens ? d_grad + ((size_t)s0 * nobs + o) * n : d_grad + (size_t)s0 * n;
            frontal_gradform_kernel<<<gb, 256, 0, st>>>(h->d_fw.p, h->d_fv.p, io.in, (long long)m, n, h->d_dptr.p, h->d_dcell.p,
                                                        h->d_cells.p, h->d_Ke.p, h->coef_mode, gout,
                                                        sens ? (long long)nobs * n : (long long)n);
            h->launches += 3;
        }
        TFIN_CUDA(cudaGetLastError());
        h->launches += 3;
    }
    h->last_solver = 2;
    h->last_fkernel = 1;
    h->last_path = 4;
    return 0;
}

// Fin.gradient / Fin.sensitivity on meshes whose front is too wide for the sample-per-thread kernel: the sample-per-CTA
// kernel solves A w = b (solution + observables), then A v = -B_obs^T (qoi - data) (or one unit row per observable) by
// factorising again -- its factor blocks are per CTA, not per sample -- and the gradient form follows.
static int launch_frontal_gradient_cta(tfin_ctx* h, const FrontalGeom& g, const double* d_k, int64_t N, const double* d_data,
                                       int64_t data_stride, double* d_grad, double* d_cost, double* d_qoi, int* d_status,
                                       cudaStream_t st, bool sens) {
    const int n = h->n, nobs = h->n_obs;
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(N, ((int64_t)1 << 30) / ((int64_t)n * 8)));
    if (int e = h->d_fw.reserve((size_t)chunk * n)) return e;
    if (int e = h->d_fv.reserve((size_t)chunk * n)) return e;
    if (!d_qoi) {
        if (int e = h->d_qoi.reserve((size_t)N * nobs)) return e;
        d_qoi = h->d_qoi.p;
    }
    for (int64_t s0 = 0; s0 < N; s0 += chunk) {
        const int64_t m = std::min<int64_t>(chunk, N - s0);
        const double* kin = d_k + (size_t)s0 * n;
        if (int e = launch_frontal(h, true, g, kin, n, m, h->d_fw.p, d_qoi + (size_t)s0 * nobs, nullptr,
                                   d_status ? d_status + s0 : nullptr, nullptr, st))
            return e;
        const int64_t total = m * (int64_t)n;
        const int gb = (int)std::min<int64_t>((total + 255) / 256, (int64_t)h->sm_count * 16);
        const int n_adj = sens ? nobs : 1;
        for (int o = 0; o < n_adj; ++o) {
            FrontalAdjArgs a{};
            a.mode = sens ? 2 : 1;
            a.qoi_in = d_qoi + (size_t)s0 * nobs;
            a.data = sens ? nullptr : d_data + (data_stride ? (size_t)s0 * data_stride : 0);
            a.data_stride = data_stride;
            a.cost_out = (!sens && d_cost) ? d_cost + s0 : nullptr;
            a.unit_row = o;
            if (int e = launch_frontal(h, true, g, kin, n, m, h->d_fv.p, nullptr, nullptr, nullptr, nullptr, st, &a)) return e;
            double* gout = sens ? d_grad + ((size_t)s0 * nobs + o) * n : d_grad + (size_t)s0 * n;
            frontal_gradform_kernel<<<gb, 256, 0, st>>>(h->d_fw.p, h->d_fv.p, kin, (long long)m, n, h->d_dptr.p, h->d_dcell.p,
                                                        h->d_cells.p, h->d_Ke.p, h->coef_mode, gout,
                                                        sens ? (long long)nobs * n : (long long)n);
            TFIN_CUDA(cudaGetLastError());
            h->launches += 1;
        }
    }
    h->last_solver = 2;
    h->last_fkernel = 3;
    h->last_path = 4;
    return 0;
}

// Forward solve of one device-resident batch with whichever solver the handle is set to: the sparse-direct kernels where
// the front fits on chip (fom_solver 0 / 2), else the on-chip PCG.  Adjoint variants exist for the PCG only.
static int launch_fom(tfin_ctx* h, bool nodal, const double* d_in, int in_stride, int64_t N, double tol, int maxit,
                      double* d_w, double* d_qoi, int* d_iters, int* d_status, double* d_relres, cudaStream_t st,
                      const PcgAdj* adjoint = nullptr) {
    if (!adjoint && h->fom_solver != 1 && h->precision == 64) {
        const FrontalGeom g = frontal_geom(h, nodal, d_w != nullptr);
        if (g.kernel != 0)
            return launch_frontal(h, nodal, g, d_in, in_stride, N, d_w, d_qoi, d_iters, d_status, d_relres, st);
        if (h->fom_solver == 2) {
            const FrontalSet& fs = nodal ? h->fr_nod : h->fr_aff;
            return fail(TFIN_E_STATE, "fom_solver = 2 (direct) but the direct solver is unavailable: %s",
                        fs.ok ? "front does not fit shared memory" : fs.why.c_str());
        }
    }
    h->last_solver = 1;
    return launch_pcg(h, nodal, d_in, in_stride, N, tol, maxit, d_w, d_qoi, d_iters, d_status, d_relres, st, adjoint);
}

static int launch_project(tfin_ctx* h, const CsrRows& op, const double* d_k, int64_t N, double* d_out,
                          cudaStream_t st) {
    const int64_t warps = N * op.rows;
    const int blocks = (int)std::min<int64_t>((warps + 7) / 8, (int64_t)h->sm_count * 16);
    csr_project_kernel<<<std::max(blocks, 1), 256, 0, st>>>(op, d_k, (long long)N, h->n, d_out);
    TFIN_CUDA(cudaGetLastError());
    h->launches += 1;
    return 0;
}

// Host <-> device staging for one batch call.
struct Staged {
    tfin_ctx* h;
    cudaStream_t st;
    bool host;
    template <typename T>
    int in(const T* src, size_t count, DevBuf<T>& buf, const T** dev) {
        if (!host) {
            *dev = src;
            return 0;
        }
        if (int e = buf.reserve(count)) return e;
        TFIN_CUDA(cudaMemcpyAsync(buf.p, src, count * sizeof(T), cudaMemcpyHostToDevice, st));
        *dev = buf.p;
        return 0;
    }
    template <typename T>
    int out_alloc(T* dst, size_t count, DevBuf<T>& buf, T** dev) {
        if (!dst) {
            *dev = nullptr;
            return 0;
        }
        if (!host) {
            *dev = dst;
            return 0;
        }
        if (int e = buf.reserve(count)) return e;
        *dev = buf.p;
        return 0;
    }
    template <typename T>
    int out_copy(T* dst, size_t count, const T* dev) {
        if (!dst || !host) return 0;
        TFIN_CUDA(cudaMemcpyAsync(dst, dev, count * sizeof(T), cudaMemcpyDeviceToHost, st));
        return 0;
    }
};

// ------------------------------------------------------------------------------------------------ solves
extern "C" int tfin_subfin_avg(tfin_handle_t h, const double* k, int64_t N, int32_t mem, double* theta_out,
                               void* stream) {
    CHECK_HANDLE(h);
    if (h->n_avg <= 0) return fail(TFIN_E_STATE, "tfin_subfin_avg: call tfin_set_averaging first");
    if (N < 0 || (N > 0 && (!k || !theta_out))) return fail(TFIN_E_ARG, "tfin_subfin_avg: bad argument");
    if (N == 0) return 0;
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    Staged sg{h, st, mem == TFIN_MEM_HOST};
    const double* d_k;
    double* d_out;
    if (int e = sg.in(k, (size_t)N * h->n, h->d_in, &d_k)) return e;
    if (int e = sg.out_alloc(theta_out, (size_t)N * h->n_avg, h->d_theta, &d_out)) return e;
    CsrRows avg{h->n_avg, h->d_avg_ptr.p, h->d_avg_idx.p, h->d_avg_val.p};
    if (int e = launch_project(h, avg, d_k, N, d_out, st)) return e;
    if (int e = sg.out_copy(theta_out, (size_t)N * h->n_avg, d_out)) return e;
    if (sg.host) TFIN_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// Nodal solve from HOST buffers, pipelined: the fields are 8 n bytes per sample (12.8 KB at n = 1597), so a plain
// "copy everything, solve, copy back" leaves the SMs idle for the whole transfer -- and from pageable numpy memory the
// transfer is not much faster than the solve.  Chunks of `host_chunk` samples are double buffered instead: H2D of chunk
// c + 1 (and D2H of the solutions of chunk c - 1, if requested) run on a second stream while chunk c is being solved.
// With `adj` (gradient mode of the adjoint kernels) the per-sample vector that streams back is the gradient instead of
// the solution: w_out then receives grad, adj->data is a DEVICE pointer to the (1 | N, n_obs) observations.
static int fom_nodal_host_pipelined(tfin_ctx* h, const double* k, int64_t N, double tol, int maxit, double* w_out,
                                    double* qoi_out, int32_t* iters_out, int32_t* status_out, double* relres_out,
                                    cudaStream_t st, const PcgAdj* adj = nullptr, double* cost_out = nullptr,
                                    bool affine_avg = false) {
    // affine_avg: AffineROMFin.forward(k) -- the fields are averaged over the sub-fins (K0) and the AFFINE on-chip
    // kernel solves; otherwise the nodal kernel consumes the fields directly
    const int n = h->n, nobs = h->n_obs, nparam = h->n_terms - 1;
    const int64_t chunk = h->host_chunk;
    if (affine_avg)
        if (int e = h->d_theta.reserve((size_t)2 * chunk * nparam)) return e;
    if (!h->copy_stream) {
        TFIN_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            TFIN_CUDA(cudaEventCreateWithFlags(&h->ev_in[b], cudaEventDisableTiming));
            TFIN_CUDA(cudaEventCreateWithFlags(&h->ev_comp[b], cudaEventDisableTiming));
            TFIN_CUDA(cudaEventCreateWithFlags(&h->ev_out[b], cudaEventDisableTiming));
        }
    }
    cudaStream_t cs = h->copy_stream;
    for (int b = 0; b < 2; ++b) {
        if (int e = h->d_pin[b].reserve((size_t)chunk * n)) return e;
        if (w_out)
            if (int e = h->d_pw[b].reserve((size_t)chunk * n)) return e;
    }
    double *d_qoi = nullptr, *d_relres = nullptr;
    int *d_iters = nullptr, *d_status = nullptr;
    if (qoi_out) {
        if (int e = h->d_qoi.reserve((size_t)N * nobs)) return e;
        d_qoi = h->d_qoi.p;
    }
    if (iters_out) {
        if (int e = h->d_iters.reserve((size_t)N)) return e;
        d_iters = h->d_iters.p;
    }
    if (status_out) {
        if (int e = h->d_status.reserve((size_t)N)) return e;
        d_status = h->d_status.p;
    }
    if (relres_out) {
        if (int e = h->d_relres.reserve((size_t)N)) return e;
        d_relres = h->d_relres.p;
    }
    double* d_cost = nullptr;
    if (cost_out) {
        if (int e = h->d_cost.reserve((size_t)N)) return e;
        d_cost = h->d_cost.p;
    }
    const int64_t n_chunks = (N + chunk - 1) / chunk;
    auto drain_w = [&](int64_t c) -> int {  // solutions of chunk c: device staging -> host, on the copy stream
        const int b = (int)(c & 1);
        const int64_t s0 = c * chunk, m = std::min<int64_t>(chunk, N - s0);
        TFIN_CUDA(cudaStreamWaitEvent(cs, h->ev_comp[b], 0));
        TFIN_CUDA(cudaMemcpyAsync(w_out + (size_t)s0 * n, h->d_pw[b].p, (size_t)m * n * 8, cudaMemcpyDeviceToHost, cs));
        TFIN_CUDA(cudaEventRecord(h->ev_out[b], cs));
        return 0;
    };
    for (int64_t c = 0; c < n_chunks; ++c) {
        const int b = (int)(c & 1);
        const int64_t s0 = c * chunk, m = std::min<int64_t>(chunk, N - s0);
        if (c >= 2) TFIN_CUDA(cudaStreamWaitEvent(cs, h->ev_comp[b], 0));  // chunk c-2 has consumed this input buffer
        TFIN_CUDA(cudaMemcpyAsync(h->d_pin[b].p, k + (size_t)s0 * n, (size_t)m * n * 8, cudaMemcpyHostToDevice, cs));
        TFIN_CUDA(cudaEventRecord(h->ev_in[b], cs));
        TFIN_CUDA(cudaStreamWaitEvent(st, h->ev_in[b], 0));
        if (w_out && c >= 2) TFIN_CUDA(cudaStreamWaitEvent(st, h->ev_out[b], 0));  // solution staging b is free again
        PcgAdj a{};
        if (adj) {
            a = *adj;
            a.data = adj->data + (adj->data_stride ? (size_t)s0 * nobs : 0);
            a.grad_out = h->d_pw[b].p;
            a.cost_out = d_cost ? d_cost + s0 : nullptr;
        }
        const double* d_par = h->d_pin[b].p;
        int stride = n;
        if (affine_avg) {
            double* th = h->d_theta.p + (size_t)b * chunk * nparam;
            CsrRows avg{h->n_avg, h->d_avg_ptr.p, h->d_avg_idx.p, h->d_avg_val.p};
            if (int e = launch_project(h, avg, h->d_pin[b].p, m, th, st)) return e;
            d_par = th;
            stride = nparam;
        }
        if (int e = launch_fom(h, !affine_avg, d_par, stride, m, tol, maxit, (w_out && !adj) ? h->d_pw[b].p : nullptr,
                               d_qoi ? d_qoi + (size_t)s0 * nobs : nullptr, d_iters ? d_iters + s0 : nullptr,
                               d_status ? d_status + s0 : nullptr, d_relres ? d_relres + s0 : nullptr, st,
                               adj ? &a : nullptr))
            return e;
        TFIN_CUDA(cudaEventRecord(h->ev_comp[b], st));
        // vectors of chunk c-1 go back while chunk c is being solved.  Issued AFTER the launch: a copy to pageable host
        // memory blocks the calling thread until it has completed
        if (w_out && c >= 1)
            if (int e = drain_w(c - 1)) return e;
    }
    if (w_out)
        if (int e = drain_w(n_chunks - 1)) return e;
    if (qoi_out) TFIN_CUDA(cudaMemcpyAsync(qoi_out, d_qoi, (size_t)N * nobs * 8, cudaMemcpyDeviceToHost, st));
    if (iters_out) TFIN_CUDA(cudaMemcpyAsync(iters_out, d_iters, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
    if (status_out) TFIN_CUDA(cudaMemcpyAsync(status_out, d_status, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
    if (relres_out) TFIN_CUDA(cudaMemcpyAsync(relres_out, d_relres, (size_t)N * 8, cudaMemcpyDeviceToHost, st));
    if (cost_out) TFIN_CUDA(cudaMemcpyAsync(cost_out, d_cost, (size_t)N * 8, cudaMemcpyDeviceToHost, st));
    TFIN_CUDA(cudaStreamSynchronize(st));
    TFIN_CUDA(cudaStreamSynchronize(cs));
    return 0;
}

static int fom_common(tfin_handle_t h, bool nodal_op, const double* in, int64_t N, int32_t in_kind, int32_t mem,
                      double tol, int32_t maxit, double* w_out, double* qoi_out, int32_t* iters_out,
                      int32_t* status_out, double* relres_out, void* stream) {
    if (h->n <= 0) return fail(TFIN_E_STATE, "FOM solve: call tfin_set_operator first");
    if (N < 0 || (N > 0 && !in)) return fail(TFIN_E_ARG, "FOM solve: bad batch argument");
    if (!(tol > 0.0) || maxit < 1) return fail(TFIN_E_ARG, "FOM solve: tol must be > 0 and maxit >= 1");
    if (qoi_out && h->n_obs <= 0) return fail(TFIN_E_STATE, "FOM solve: qoi requested but no observation operator");
    if (nodal_op && h->n_cells <= 0) return fail(TFIN_E_STATE, "tfin_fom_nodal: call tfin_set_cells first");
    if (N == 0) return 0;
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    if (nodal_op && mem == TFIN_MEM_HOST && h->small_ok && h->host_chunk > 0 && N > h->host_chunk)
        return fom_nodal_host_pipelined(h, in, N, tol, maxit, w_out, qoi_out, iters_out, status_out, relres_out, st);
    if (!nodal_op && in_kind == TFIN_IN_NODAL && mem == TFIN_MEM_HOST && h->small_ok && h->precision == 64 &&
        h->pcg_path != 2 && h->W <= 4 && h->n <= 4096 && h->host_chunk > 0 && N > h->host_chunk) {
        // AffineROMFin.forward(k) from host fields: same pipeline, sub-fin averaging + affine kernel per chunk.  (Only
        // where an on-chip variant is certain to exist; other meshes take the single-shot path with its fallbacks.)
        if (h->n_avg != h->n_terms - 1)
            return fail(TFIN_E_STATE, "nodal input needs tfin_set_averaging with %d rows", h->n_terms - 1);
        return fom_nodal_host_pipelined(h, in, N, tol, maxit, w_out, qoi_out, iters_out, status_out, relres_out, st,
                                        nullptr, nullptr, true);
    }
    Staged sg{h, st, mem == TFIN_MEM_HOST};
    const int nparam = h->n_terms - 1;
    const bool nodal_in = nodal_op || in_kind == TFIN_IN_NODAL;
    const int in_cols = nodal_in ? h->n : nparam;
    const double* d_in;
    if (int e = sg.in(in, (size_t)N * in_cols, h->d_in, &d_in)) return e;
    const double* d_par = d_in;
    int stride = in_cols;
    if (!nodal_op && nodal_in) {  // AffineROMFin.forward(k): theta = subfin_avg_op(k)
        if (h->n_avg != nparam) return fail(TFIN_E_STATE, "nodal input needs tfin_set_averaging with %d rows", nparam);
        if (int e = h->d_theta.reserve((size_t)N * nparam)) return e;
        CsrRows avg{h->n_avg, h->d_avg_ptr.p, h->d_avg_idx.p, h->d_avg_val.p};
        if (int e = launch_project(h, avg, d_in, N, h->d_theta.p, st)) return e;
        d_par = h->d_theta.p;
        stride = nparam;
    }
    double *d_w, *d_qoi, *d_relres;
    int *d_iters, *d_status;
    if (int e = sg.out_alloc(w_out, (size_t)N * h->n, h->d_w, &d_w)) return e;
    if (int e = sg.out_alloc(qoi_out, (size_t)N * h->n_obs, h->d_qoi, &d_qoi)) return e;
    if (int e = sg.out_alloc(iters_out, (size_t)N, h->d_iters, &d_iters)) return e;
    if (int e = sg.out_alloc(status_out, (size_t)N, h->d_status, &d_status)) return e;
    if (int e = sg.out_alloc(relres_out, (size_t)N, h->d_relres, &d_relres)) return e;
    // sparse-direct solver first (fom_solver 0 = where available, 2 = required); pcg_path pins a PCG kernel
    if (h->fom_solver != 1 && h->precision == 64 && (h->pcg_path == 0 || h->fom_solver == 2)) {
        const FrontalGeom g = frontal_geom(h, nodal_op, d_w != nullptr);
        if (g.kernel != 0) {
            if (int e = launch_frontal(h, nodal_op, g, d_par, stride, N, d_w, d_qoi, d_iters, d_status, d_relres, st)) return e;
            if (int e = sg.out_copy(w_out, (size_t)N * h->n, d_w)) return e;
            if (int e = sg.out_copy(qoi_out, (size_t)N * h->n_obs, d_qoi)) return e;
            if (int e = sg.out_copy(iters_out, (size_t)N, d_iters)) return e;
            if (int e = sg.out_copy(status_out, (size_t)N, d_status)) return e;
            if (int e = sg.out_copy(relres_out, (size_t)N, d_relres)) return e;
            if (sg.host) TFIN_CUDA(cudaStreamSynchronize(st));
            return 0;
        }
        if (h->fom_solver == 2) {
            const FrontalSet& fs = nodal_op ? h->fr_nod : h->fr_aff;
            return fail(TFIN_E_STATE, "fom_solver = 2 (direct) but the direct solver is unavailable: %s",
                        fs.ok ? "front does not fit shared memory" : fs.why.c_str());
        }
    }
    h->last_solver = 1;
    bool use_stream = !nodal_op && (h->pcg_path == 2 || (h->pcg_path == 0 && !h->small_ok));
    if (!use_stream && !h->small_ok)
        return fail(TFIN_E_STATE, "on-chip PCG needs n <= 8191 (n = %d); use the streaming path", h->n);
    const bool use_f32 = !use_stream && !nodal_op && h->precision == 32;
    int rc = 0;
    if (!use_stream) {
        rc = use_f32 ? launch_pcg_f32(h, d_par, stride, N, tol, maxit, d_w, d_qoi, d_iters, d_status, d_relres, st)
                     : launch_pcg(h, nodal_op, d_par, stride, N, tol, maxit, d_w, d_qoi, d_iters, d_status, d_relres, st);
        // no compiled on-chip variant fits this mesh (CG state + per-sample operator exceed one SM, n >~ 4100):
        // the affine operator falls back to the streaming kernel; the nodal operator has no streaming kernel
        if (rc == TFIN_E_STATE && !nodal_op && !use_f32 && h->pcg_path == 0 && h->stream_ok) use_stream = true;
        else if (rc == TFIN_E_STATE && nodal_op)
            return fail(TFIN_E_STATE, "tfin_fom_nodal: the nodal-conductivity kernel keeps the CG state and the per-sample "
                        "operator on one SM and no compiled variant fits n = %d (ell width %d); meshes up to about 4100 "
                        "dofs are supported", h->n, h->Wn);
        else if (rc) return rc;
    }
    if (use_stream)
        rc = launch_pcg_stream(h, d_par, stride, N, tol, maxit, d_w, d_qoi, d_iters, d_status, d_relres, st);
    if (rc) return rc;
    if (int e = sg.out_copy(w_out, (size_t)N * h->n, d_w)) return e;
    if (int e = sg.out_copy(qoi_out, (size_t)N * h->n_obs, d_qoi)) return e;
    if (int e = sg.out_copy(iters_out, (size_t)N, d_iters)) return e;
    if (int e = sg.out_copy(status_out, (size_t)N, d_status)) return e;
    if (int e = sg.out_copy(relres_out, (size_t)N, d_relres)) return e;
    if (sg.host) TFIN_CUDA(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int tfin_fom_affine(tfin_handle_t h, const double* in, int64_t N, int32_t in_kind, int32_t mem,
                               double tol, int32_t maxit, double* w_out, double* qoi_out, int32_t* iters_out,
                               int32_t* status_out, double* relres_out, void* stream) {
    CHECK_HANDLE(h);
    if (in_kind != TFIN_IN_PARAMS && in_kind != TFIN_IN_NODAL) return fail(TFIN_E_ARG, "tfin_fom_affine: bad in_kind");
    return fom_common(h, false, in, N, in_kind, mem, tol, maxit, w_out, qoi_out, iters_out, status_out, relres_out,
                      stream);
}

extern "C" int tfin_fom_nodal(tfin_handle_t h, const double* k, int64_t N, int32_t mem, double tol, int32_t maxit,
                              double* w_out, double* qoi_out, int32_t* iters_out, int32_t* status_out,
                              double* relres_out, void* stream) {
    CHECK_HANDLE(h);
    return fom_common(h, true, k, N, TFIN_IN_NODAL, mem, tol, maxit, w_out, qoi_out, iters_out, status_out,
                      relres_out, stream);
}

// Producer of the packed reduced systems of a chunk: the affine Gram combination (R1, theta input) or the per-sample
// nodal assembly + Gram kernel (R4, nodal k input).
struct RomSrc {
    bool nodal = false;
    const double* d_in = nullptr;   // theta (N, n_terms-1) | k (N, n)
    double* d_Ar = nullptr;         // nodal only: dense A_r / B_r outputs (optional)
    double* d_Br = nullptr;
};

// Reduced systems + Cholesky over chunks of samples (R1|R4 + R2); with `adj` the Cholesky kernel also solves the
// reduced adjoint.
static int rom_run(tfin_ctx* h, const RomSrc& src, int nr, int nobs, const double* d_obs_phi, int64_t N,
                   cudaStream_t st, double* d_wr, double* d_qoi, int* d_status, const RomAdj* adj) {
    const int nt = h->rom_terms;
    const int Taug = rom_taug(nr), P2 = nt * (nt + 1) / 2;
    const int per_warp = ((Taug + 2 * nr + 2) + 1) & ~1;
    int wpb = std::min<int>(8, (int)((size_t)(h->max_smem_optin - 1024) / ((size_t)per_warp * 8)));
    if (wpb < 1) return fail(TFIN_E_STATE, "tfin_rom: n_r = %d does not fit shared memory", nr);
    const size_t chol_smem = (size_t)wpb * per_warp * 8 + 256;  // + slack for the unchecked panel-sweep loads
    const int64_t chunk = h->rom_chunk > 0 ? h->rom_chunk : (int64_t)h->sm_count * wpb * 16;   // 16 samples per warp: 516 MB of packed operators at n_r = 81
    if (int e = h->d_romC.reserve((size_t)std::min<int64_t>(chunk, N) * Taug)) return e;
    size_t comb_smem = 0, gram_smem = 0;
    int gram_threads = 0, gram_occ = 1;
    if (!src.nodal) {
        comb_smem = rom_combine_smem(nt);
        if (comb_smem > (size_t)h->max_smem_optin)
            return fail(TFIN_E_STATE, "tfin_rom: %d affine terms need %zu bytes of shared memory for the operator combination (max 12 terms)", nt, comb_smem);
        TFIN_CUDA(cudaFuncSetAttribute(rom_combine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)comb_smem));
    } else {
        const int TT = h->b_TT;
        gram_smem = RomNodalSmem::make(h->n_cells, 6 * TT, h->Wn).total;
        gram_threads = (TT * (TT + 1) / 2 + 31) & ~31;
        if (gram_smem > (size_t)h->max_smem_optin || gram_threads > 256)
            return fail(TFIN_E_STATE, "nodal LSPG: n_r = %d / n_cells = %d do not fit one CTA", nr, h->n_cells);
        TFIN_CUDA(cudaFuncSetAttribute(rom_nodal_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gram_smem));
        TFIN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&gram_occ, rom_nodal_gram_kernel, gram_threads, gram_smem));
        if (gram_occ < 1) return fail(TFIN_E_STATE, "nodal LSPG: kernel does not fit an SM");
    }
    const int maxm = (nr + 1 + 31) / 32;
    auto chol = adj ? (maxm == 1 ? rom_chol_kernel<1, true> : maxm == 2 ? rom_chol_kernel<2, true>
                       : maxm == 3 ? rom_chol_kernel<3, true> : rom_chol_kernel<4, true>)
                    : (maxm == 1 ? rom_chol_kernel<1, false> : maxm == 2 ? rom_chol_kernel<2, false>
                       : maxm == 3 ? rom_chol_kernel<3, false> : rom_chol_kernel<4, false>);
    TFIN_CUDA(cudaFuncSetAttribute(chol, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chol_smem));
    int chol_occ = 1;
    TFIN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&chol_occ, chol, wpb * 32, chol_smem));
    if (chol_occ < 1) return fail(TFIN_E_STATE, "tfin_rom: the Cholesky kernel does not fit an SM (n_r = %d)", nr);
    const RomAdj a = adj ? *adj : RomAdj{};
    for (int64_t s0 = 0; s0 < N; s0 += chunk) {
        const int64_t s1 = std::min<int64_t>(N, s0 + chunk);
        if (!src.nodal) {
            const int g1 = (int)std::min<int64_t>((s1 - s0 + ROM_BM - 1) / ROM_BM, (int64_t)h->sm_count);
            const int ldS = (Taug + ROM_BN - 1) / ROM_BN * ROM_BN;
            rom_combine_kernel<<<g1, 256, comb_smem, st>>>(src.d_in, s0, s1, nt, h->d_S.p, ldS, Taug, h->d_romC.p);
        } else {
            PcgOp op{};
            op.n = h->n; op.ld = h->ld; op.W = h->Wn; op.n_cells = h->n_cells; op.rhs = h->d_rhs.p;
            op.col = h->d_ncol.p; op.cell = h->d_ncell.p; op.coef = h->d_ncoef.p; op.cst = h->d_ncst.p;
            op.dptr = h->d_dptr.p; op.dcell = h->d_dcell.p; op.dcoef = h->d_dcoef.p; op.dcst = h->d_dcst.p;
            op.cells = h->d_cells.p;
            op.coef_mode = h->coef_mode;
            const int g1 = (int)std::min<int64_t>(s1 - s0, (int64_t)h->sm_count * gram_occ);
            rom_nodal_gram_kernel<<<g1, gram_threads, gram_smem, st>>>(op, src.d_in, s0, s1, h->d_bphi.p, nr, h->b_TT,
                                                                       h->d_romC.p);
            if (src.d_Ar || src.d_Br) {
                const int64_t total = (s1 - s0) * (int64_t)nr * (nr + 1);
                const int gb = (int)std::min<int64_t>((total + 255) / 256, (int64_t)h->sm_count * 16);
                rom_unpack_kernel<<<gb, 256, 0, st>>>(h->d_romC.p, s0, s1, nr, src.d_Ar, src.d_Br);
                h->launches += 1;
            }
        }
        // persistent grid: every warp keeps its observation rows in registers across the samples it takes
        const int g2 = (int)std::min<int64_t>((s1 - s0 + wpb - 1) / wpb, (int64_t)h->sm_count * chol_occ);
        chol<<<g2, wpb * 32, chol_smem, st>>>(h->d_romC.p, s0, s1, nr, nobs, d_obs_phi, d_wr, d_qoi, d_status, a);
        h->launches += 2;
    }
    TFIN_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tfin_rom(tfin_handle_t h, const double* in, int64_t N, int32_t in_kind, int32_t mem, double* wr_out,
                        double* qoi_out, int32_t* status_out, void* stream) {
    CHECK_HANDLE(h);
    if (h->n_r <= 0) return fail(TFIN_E_STATE, "tfin_rom: call tfin_set_rom first");
    if (in_kind != TFIN_IN_PARAMS && in_kind != TFIN_IN_NODAL) return fail(TFIN_E_ARG, "tfin_rom: bad in_kind");
    if (N < 0 || (N > 0 && !in)) return fail(TFIN_E_ARG, "tfin_rom: bad batch argument");
    if (N == 0) return 0;
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    Staged sg{h, st, mem == TFIN_MEM_HOST};
    const int nt = h->rom_terms, nparam = nt - 1, nr = h->n_r, nobs = h->rom_obs;
    const bool nodal_in = in_kind == TFIN_IN_NODAL;
    if (nodal_in && h->n <= 0) return fail(TFIN_E_STATE, "tfin_rom: nodal input needs tfin_set_operator");
    const int in_cols = nodal_in ? h->n : nparam;
    const double* d_in;
    if (int e = sg.in(in, (size_t)N * in_cols, h->d_in, &d_in)) return e;
    const double* d_par = d_in;
    if (nodal_in) {  // forward_reduced(k): theta = subfin_avg_op(k), averaged_affine_ROM.py:274
        if (h->n_avg != nparam) return fail(TFIN_E_STATE, "tfin_rom: nodal input needs tfin_set_averaging with %d rows", nparam);
        if (int e = h->d_theta.reserve((size_t)N * nparam)) return e;
        CsrRows avg{h->n_avg, h->d_avg_ptr.p, h->d_avg_idx.p, h->d_avg_val.p};
        if (int e = launch_project(h, avg, d_in, N, h->d_theta.p, st)) return e;
        d_par = h->d_theta.p;
    }
    double *d_wr, *d_qoi;
    int* d_status;
    if (int e = sg.out_alloc(wr_out, (size_t)N * nr, h->d_wr, &d_wr)) return e;
    if (int e = sg.out_alloc(qoi_out, (size_t)N * nobs, h->d_qoi, &d_qoi)) return e;
    if (int e = sg.out_alloc(status_out, (size_t)N, h->d_status, &d_status)) return e;

    RomSrc src;
    src.d_in = d_par;
    if (int e = rom_run(h, src, nr, nobs, h->d_obs_phi.p, N, st, d_wr, d_qoi, d_status, nullptr)) return e;
    TFIN_CUDA(cudaGetLastError());
    if (int e = sg.out_copy(wr_out, (size_t)N * nr, d_wr)) return e;
    if (int e = sg.out_copy(qoi_out, (size_t)N * nobs, d_qoi)) return e;
    if (int e = sg.out_copy(status_out, (size_t)N, d_status)) return e;
    if (sg.host) TFIN_CUDA(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int tfin_set_basis(tfin_handle_t h, int32_t n, int32_t n_r, const double* phi, int32_t n_out,
                              const double* out_phi) {
    CHECK_HANDLE(h);
    if (h->n <= 0) return fail(TFIN_E_STATE, "tfin_set_basis: call tfin_set_operator first");
    if (!phi || n != h->n || n_r <= 0 || n_r > 127 || n_out <= 0 || !out_phi)
        return fail(TFIN_E_ARG, "tfin_set_basis: bad argument (n must match the operator, n_r in [1,127])");
    const int TT = (n_r + 1 + 5) / 6, nrp = 6 * TT;  // one spare column for the right-hand side
    std::vector<double> pad((size_t)n * nrp, 0.0);
    for (int i = 0; i < n; ++i) std::copy(phi + (size_t)i * n_r, phi + (size_t)(i + 1) * n_r, pad.begin() + (size_t)i * nrp);
    std::vector<double> op(out_phi, out_phi + (size_t)n_out * n_r);
    if (int e = h->d_bphi.upload(pad, h->stream)) return e;
    if (int e = h->d_bout.upload(op, h->stream)) return e;
    TFIN_CUDA(cudaStreamSynchronize(h->stream));
    h->b_nr = n_r;
    h->b_nout = n_out;
    h->b_TT = TT;
    return 0;
}

extern "C" int tfin_rom_nodal(tfin_handle_t h, const double* k, int64_t N, int32_t mem, double* Ar_out,
                              double* Br_out, double* xr_out, double* y_out, int32_t* status_out, void* stream) {
    CHECK_HANDLE(h);
    if (h->b_nr <= 0) return fail(TFIN_E_STATE, "tfin_rom_nodal: call tfin_set_basis first");
    if (h->n_cells <= 0) return fail(TFIN_E_STATE, "tfin_rom_nodal: call tfin_set_cells first");
    if (N < 0 || (N > 0 && !k)) return fail(TFIN_E_ARG, "tfin_rom_nodal: bad batch argument");
    if (N == 0) return 0;
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    Staged sg{h, st, mem == TFIN_MEM_HOST};
    const int nr = h->b_nr, nout = h->b_nout;
    const double* d_k;
    if (int e = sg.in(k, (size_t)N * h->n, h->d_in, &d_k)) return e;
    double *d_Ar, *d_Br, *d_xr, *d_y;
    int* d_status;
    if (int e = sg.out_alloc(Ar_out, (size_t)N * nr * nr, h->d_Ar, &d_Ar)) return e;
    if (int e = sg.out_alloc(Br_out, (size_t)N * nr, h->d_Br, &d_Br)) return e;
    if (int e = sg.out_alloc(xr_out, (size_t)N * nr, h->d_wr, &d_xr)) return e;
    if (int e = sg.out_alloc(y_out, (size_t)N * nout, h->d_y, &d_y)) return e;
    if (int e = sg.out_alloc(status_out, (size_t)N, h->d_status, &d_status)) return e;
    RomSrc src;
    src.nodal = true;
    src.d_in = d_k;
    src.d_Ar = d_Ar;
    src.d_Br = d_Br;
    if (int e = rom_run(h, src, nr, nout, h->d_bout.p, N, st, d_xr, d_y, d_status, nullptr)) return e;
    if (int e = sg.out_copy(Ar_out, (size_t)N * nr * nr, d_Ar)) return e;
    if (int e = sg.out_copy(Br_out, (size_t)N * nr, d_Br)) return e;
    if (int e = sg.out_copy(xr_out, (size_t)N * nr, d_xr)) return e;
    if (int e = sg.out_copy(y_out, (size_t)N * nout, d_y)) return e;
    if (int e = sg.out_copy(status_out, (size_t)N, d_status)) return e;
    if (sg.host) TFIN_CUDA(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int tfin_rom_gradient(tfin_handle_t h, const double* in, int64_t N, int32_t in_kind, int32_t mem,
                                 const double* data, int64_t data_rows, int32_t grad_kind, double* grad_out,
                                 double* cost_out, double* qoi_out, double* wr_out, int32_t* status_out,
                                 void* stream) {
    CHECK_HANDLE(h);
    if (h->n_r <= 0 || h->rg_ob <= 0)
        return fail(TFIN_E_STATE, "tfin_rom_gradient: call tfin_set_rom and tfin_set_rom_gradient first");
    if ((in_kind != TFIN_IN_PARAMS && in_kind != TFIN_IN_NODAL) || (grad_kind != TFIN_IN_PARAMS && grad_kind != TFIN_IN_NODAL))
        return fail(TFIN_E_ARG, "tfin_rom_gradient: bad in_kind / grad_kind");
    if (N < 0 || (N > 0 && (!in || !grad_out || !data || (data_rows != 1 && data_rows != N))))
        return fail(TFIN_E_ARG, "tfin_rom_gradient: bad batch argument (data must have 1 or N rows)");
    if (N == 0) return 0;
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    Staged sg{h, st, mem == TFIN_MEM_HOST};
    const int nt = h->rom_terms, nparam = nt - 1, nr = h->n_r, nobs = h->rom_obs;
    const bool nodal_in = in_kind == TFIN_IN_NODAL, nodal_out = grad_kind == TFIN_IN_NODAL;
    if ((nodal_in || nodal_out) && (h->n <= 0 || h->n_avg != nparam))
        return fail(TFIN_E_STATE, "tfin_rom_gradient: nodal input/output needs tfin_set_operator and tfin_set_averaging with %d rows", nparam);
    const int in_cols = nodal_in ? h->n : nparam;
    const double *d_in, *d_data;
    if (int e = sg.in(in, (size_t)N * in_cols, h->d_in, &d_in)) return e;
    if (int e = sg.in(data, (size_t)data_rows * nobs, h->d_data, &d_data)) return e;
    const double* d_par = d_in;
    if (nodal_in) {  // grad_reduced(k) starts with forward_reduced(k): theta = subfin_avg_op(k), :336, :274
        if (int e = h->d_theta.reserve((size_t)N * nparam)) return e;
        CsrRows avg{h->n_avg, h->d_avg_ptr.p, h->d_avg_idx.p, h->d_avg_val.p};
        if (int e = launch_project(h, avg, d_in, N, h->d_theta.p, st)) return e;
        d_par = h->d_theta.p;
    }
    const int grad_cols = nodal_out ? h->n : nparam;
    double *d_wr, *d_qoi, *d_cost, *d_grad;
    int* d_status;
    if (int e = h->d_wr.reserve((size_t)N * nr)) return e;   // w_r is always needed by the contraction
    d_wr = (!sg.host && wr_out) ? wr_out : h->d_wr.p;
    if (int e = h->d_vr.reserve((size_t)N * nr)) return e;
    if (int e = sg.out_alloc(qoi_out, (size_t)N * nobs, h->d_qoi, &d_qoi)) return e;
    if (int e = sg.out_alloc(cost_out, (size_t)N, h->d_cost, &d_cost)) return e;
    if (int e = sg.out_alloc(status_out, (size_t)N, h->d_status, &d_status)) return e;
    if (int e = sg.out_alloc(grad_out, (size_t)N * grad_cols, h->d_grad, &d_grad)) return e;
    double* d_g = d_grad;
    if (nodal_out) {
        if (int e = h->d_gtheta.reserve((size_t)N * nparam)) return e;
        d_g = h->d_gtheta.p;
    }
    RomAdj adj{d_data, data_rows == 1 ? 0 : (long long)nobs, h->d_vr.p, d_cost};
    RomSrc src;
    src.d_in = d_par;
    if (int e = rom_run(h, src, nr, nobs, h->d_obs_phi.p, N, st, d_wr, d_qoi, d_status, &adj)) return e;
    const size_t gsm = rom_grad_smem(nr, nparam);
    if (gsm > (size_t)h->max_smem_optin) return fail(TFIN_E_STATE, "tfin_rom_gradient: n_r = %d does not fit shared memory", nr);
    TFIN_CUDA(cudaFuncSetAttribute(rom_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm));
    const int gg = (int)std::min<int64_t>((N + RG_BM - 1) / RG_BM, h->sm_count);
    rom_grad_kernel<<<gg, 256, gsm, st>>>(d_par, d_wr, h->d_vr.p, (long long)N, nr, nt, h->d_NG.p, h->rg_ob, d_g);
    h->launches += 1;
    if (nodal_out) {
        CsrRows avgT{h->n, h->d_avgT_ptr.p, h->d_avgT_idx.p, h->d_avgT_val.p};
        const int64_t total = N * h->n;
        const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)h->sm_count * 16);
        rom_grad_lift_kernel<<<blocks, 256, 0, st>>>(avgT, d_g, (long long)N, nparam, d_grad);
        h->launches += 1;
    }
    TFIN_CUDA(cudaGetLastError());
    if (int e = sg.out_copy(grad_out, (size_t)N * grad_cols, d_grad)) return e;
    if (int e = sg.out_copy(cost_out, (size_t)N, d_cost)) return e;
    if (int e = sg.out_copy(qoi_out, (size_t)N * nobs, d_qoi)) return e;
    if (int e = sg.out_copy(wr_out, (size_t)N * nr, (const double*)d_wr)) return e;
    if (int e = sg.out_copy(status_out, (size_t)N, d_status)) return e;
    if (sg.host) TFIN_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// ------------------------------------------------------------------------------------------------ adjoints
static int fom_adjoint(tfin_handle_t h, int mode, const double* k, int64_t N, int32_t mem, double tol, int32_t maxit,
                       const double* data, int64_t data_rows, double* grad_out, double* cost_out, double* qoi_out,
                       int32_t* iters_out, int32_t* status_out, void* stream) {
    if (h->n_cells <= 0) return fail(TFIN_E_STATE, "adjoint solve: call tfin_set_cells first");
    if (h->n_obs <= 0 || h->n_obs > 64) return fail(TFIN_E_STATE, "adjoint solve: needs an observation operator with <= 64 rows");
    if (N < 0 || (N > 0 && (!k || !grad_out))) return fail(TFIN_E_ARG, "adjoint solve: bad batch argument");
    if (!(tol > 0.0) || maxit < 1) return fail(TFIN_E_ARG, "adjoint solve: tol must be > 0 and maxit >= 1");
    if (mode == 1 && (!data || (data_rows != 1 && data_rows != N)))
        return fail(TFIN_E_ARG, "tfin_fom_nodal_gradient: data must have 1 or N rows");
    if (N == 0) return 0;
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    if ((mode == 1 || mode == 2) && h->fom_solver != 1 && h->precision == 64) {
        // direct solver: one factorisation serves the forward and the adjoint solve (sample-per-thread kernel), or one
        // factorisation per solve on wide fronts (sample-per-CTA kernel)
        const FrontalGeom g = frontal_geom(h, true, true);
        if ((g.kernel == 1 && g.split) || (g.kernel == 2 && g.mode == FRONTAL_MODE_SOLVE)) {
            Staged sg{h, st, mem == TFIN_MEM_HOST};
            const int n = h->n, nobs = h->n_obs;
            const double *d_k, *d_data = nullptr;
            const size_t grad_rows = mode == 2 ? (size_t)N * nobs : (size_t)N;
            if (int e = sg.in(k, (size_t)N * n, h->d_in, &d_k)) return e;
            if (mode == 1)
                if (int e = sg.in(data, (size_t)data_rows * nobs, h->d_data, &d_data)) return e;
            double *d_grad, *d_cost, *d_qoi;
            int *d_iters, *d_status;
            if (int e = sg.out_alloc(grad_out, grad_rows * n, h->d_grad, &d_grad)) return e;
            if (int e = sg.out_alloc(cost_out, (size_t)N, h->d_cost, &d_cost)) return e;
            if (int e = sg.out_alloc(qoi_out, (size_t)N * nobs, h->d_qoi, &d_qoi)) return e;
            if (int e = sg.out_alloc(iters_out, (size_t)N, h->d_iters, &d_iters)) return e;
            if (int e = sg.out_alloc(status_out, (size_t)N, h->d_status, &d_status)) return e;
            if (d_iters) TFIN_CUDA(cudaMemsetAsync(d_iters, 0, (size_t)N * sizeof(int), st));
            if (g.kernel == 1) {
                if (int e = launch_frontal_gradient(h, g, d_k, N, d_data, data_rows == 1 ? 0 : nobs, d_grad, d_cost, d_qoi, d_status,
                                                    st, mode == 2))
                    return e;
            } else {
                if (int e = launch_frontal_gradient_cta(h, g, d_k, N, d_data, data_rows == 1 ? 0 : nobs, d_grad, d_cost, d_qoi,
                                                        d_status, st, mode == 2))
                    return e;
            }
            if (int e = sg.out_copy(grad_out, grad_rows * n, d_grad)) return e;
            if (int e = sg.out_copy(cost_out, (size_t)N, d_cost)) return e;
            if (int e = sg.out_copy(qoi_out, (size_t)N * nobs, d_qoi)) return e;
            if (int e = sg.out_copy(iters_out, (size_t)N, d_iters)) return e;
            if (int e = sg.out_copy(status_out, (size_t)N, d_status)) return e;
            if (sg.host) TFIN_CUDA(cudaStreamSynchronize(st));
            return 0;
        }
        if (h->fom_solver == 2) return fail(TFIN_E_STATE, "fom_solver = 2 (direct) but the front is too wide for the direct gradient path");
    }
    if (mode == 1 && mem == TFIN_MEM_HOST && h->host_chunk > 0 && N > h->host_chunk) {  // pipelined host path
        std::vector<double> hd(data, data + (size_t)data_rows * h->n_obs);
        if (int e = h->d_data.upload(hd, st)) return e;
        PcgAdj adj{};
        adj.mode = 1;
        adj.data = h->d_data.p;
        adj.data_stride = data_rows == 1 ? 0 : h->n_obs;
        adj.obsT = CsrRows{h->n, h->d_obsT_ptr.p, h->d_obsT_idx.p, h->d_obsT_val.p};
        adj.Ke = h->d_Ke.p;
        return fom_nodal_host_pipelined(h, k, N, tol, maxit, grad_out, qoi_out, iters_out, status_out, nullptr, st, &adj,
                                        cost_out);
    }
    Staged sg{h, st, mem == TFIN_MEM_HOST};
    const int n = h->n, nobs = h->n_obs;
    const int64_t grad_rows = mode == 2 ? N * nobs : N;
    const double *d_k, *d_data = nullptr;
    if (int e = sg.in(k, (size_t)N * n, h->d_in, &d_k)) return e;
    if (mode == 1)
        if (int e = sg.in(data, (size_t)data_rows * nobs, h->d_data, &d_data)) return e;
    double *d_grad, *d_cost, *d_qoi;
    int *d_iters, *d_status;
    if (int e = sg.out_alloc(grad_out, (size_t)grad_rows * n, h->d_grad, &d_grad)) return e;
    if (int e = sg.out_alloc(cost_out, (size_t)N, h->d_cost, &d_cost)) return e;
    if (int e = sg.out_alloc(qoi_out, (size_t)N * nobs, h->d_qoi, &d_qoi)) return e;
    if (int e = sg.out_alloc(iters_out, (size_t)N, h->d_iters, &d_iters)) return e;
    if (int e = sg.out_alloc(status_out, (size_t)N, h->d_status, &d_status)) return e;
    PcgAdj adj{};
    adj.mode = mode;
    adj.data = d_data;
    adj.data_stride = data_rows == 1 ? 0 : nobs;
    adj.obsT = CsrRows{n, h->d_obsT_ptr.p, h->d_obsT_idx.p, h->d_obsT_val.p};
    adj.Ke = h->d_Ke.p;
    adj.grad_out = d_grad;
    adj.cost_out = d_cost;
    if (int e = launch_pcg(h, true, d_k, n, N, tol, maxit, nullptr, d_qoi, d_iters, d_status, nullptr, st, &adj)) return e;
    if (int e = sg.out_copy(grad_out, (size_t)grad_rows * n, d_grad)) return e;
    if (int e = sg.out_copy(cost_out, (size_t)N, d_cost)) return e;
    if (int e = sg.out_copy(qoi_out, (size_t)N * nobs, d_qoi)) return e;
    if (int e = sg.out_copy(iters_out, (size_t)N, d_iters)) return e;
    if (int e = sg.out_copy(status_out, (size_t)N, d_status)) return e;
    if (sg.host) TFIN_CUDA(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int tfin_fom_nodal_gradient(tfin_handle_t h, const double* k, int64_t N, int32_t mem, double tol,
                                       int32_t maxit, const double* data, int64_t data_rows, double* grad_out,
                                       double* cost_out, double* qoi_out, int32_t* iters_out, int32_t* status_out,
                                       void* stream) {
    CHECK_HANDLE(h);
    return fom_adjoint(h, 1, k, N, mem, tol, maxit, data, data_rows, grad_out, cost_out, qoi_out, iters_out, status_out,
                       stream);
}

extern "C" int tfin_fom_nodal_sensitivity(tfin_handle_t h, const double* k, int64_t N, int32_t mem, double tol,
                                          int32_t maxit, double* jac_out, double* qoi_out, int32_t* iters_out,
                                          int32_t* status_out, void* stream) {
    CHECK_HANDLE(h);
    return fom_adjoint(h, 2, k, N, mem, tol, maxit, nullptr, 0, jac_out, nullptr, qoi_out, iters_out, status_out, stream);
}

// ------------------------------------------------------------------------------------------------ field sampler
extern "C" int tfin_field_set_cov(tfin_handle_t h, int32_t n_pts, const double* coords, int32_t kern_type,
                                  double length, double* chol_out) {
    CHECK_HANDLE(h);
    if (n_pts <= 0 || !coords || !(length > 0.0) || kern_type < TFIN_KERN_SQ_EXP || kern_type > TFIN_KERN_M32)
        return fail(TFIN_E_ARG, "tfin_field_set_cov: bad argument");
    const int n = n_pts;
    cudaStream_t st = h->stream;
    std::vector<double> xy(coords, coords + 2 * (size_t)n);
    if (int e = h->d_fxy.upload(xy, st)) return e;
    if (int e = h->d_fL.reserve((size_t)n * n)) return e;
    if (int e = h->d_finfo.reserve(1)) return e;
    TFIN_CUDA(cudaMemsetAsync(h->d_finfo.p, 0, sizeof(int), st));
    const long long total = (long long)n * n;
    field_cov_kernel<<<(int)std::min<long long>((total + 255) / 256, (long long)h->sm_count * 16), 256, 0, st>>>(
        h->d_fxy.p, n, kern_type, length, h->d_fL.p);
    h->launches += 1;
    for (int p0 = 0; p0 < n; p0 += FC_NB) {
        field_potf2_kernel<<<1, 32, 0, st>>>(h->d_fL.p, n, p0, h->d_finfo.p);
        h->launches += 1;
        const int rest = n - p0 - FC_NB;
        if (rest > 0) {
            field_trsm_kernel<<<(rest + 127) / 128, 128, 0, st>>>(h->d_fL.p, n, p0);
            const int T = (rest + FC_NB - 1) / FC_NB;
            field_syrk_kernel<<<dim3(T, T), 256, 0, st>>>(h->d_fL.p, n, p0);
            h->launches += 2;
        }
    }
    TFIN_CUDA(cudaGetLastError());
    int info = 0;
    TFIN_CUDA(cudaMemcpyAsync(&info, h->d_finfo.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    TFIN_CUDA(cudaStreamSynchronize(st));
    if (info != 0) {
        h->f_n = 0;
        return fail(TFIN_E_STATE, "tfin_field_set_cov: covariance is not positive definite (pivot %d)", info);
    }
    h->f_n = n;
    if (chol_out) {  // upper factor chol = L^T, row-major: element (i, j) = L[j][i]
        std::vector<double> L((size_t)n * n);
        TFIN_CUDA(cudaMemcpy(L.data(), h->d_fL.p, L.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) chol_out[(size_t)i * n + j] = j >= i ? L[(size_t)j * n + i] : 0.0;
    }
    return 0;
}

extern "C" int tfin_field_set_chol(tfin_handle_t h, int32_t n_pts, const double* chol) {
    CHECK_HANDLE(h);
    if (n_pts <= 0 || !chol) return fail(TFIN_E_ARG, "tfin_field_set_chol: bad argument");
    const int n = n_pts;
    std::vector<double> L((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i)
        for (int j = i; j < n; ++j) L[(size_t)j * n + i] = chol[(size_t)i * n + j];
    if (int e = h->d_fL.upload(L, h->stream)) return e;
    TFIN_CUDA(cudaStreamSynchronize(h->stream));
    h->f_n = n;
    return 0;
}

static int launch_field_sample(tfin_ctx* h, const double* d_z, int64_t N, double* d_k, cudaStream_t st) {
    const int n = h->f_n;
    dim3 grid((unsigned)((n + FS_BN - 1) / FS_BN), (unsigned)((N + FS_BM - 1) / FS_BM));
    if (grid.y > 65535) return fail(TFIN_E_ARG, "field sampler: at most %d rows per call", 65535 * FS_BM);
    field_sample_kernel<<<grid, 256, 0, st>>>(d_z, (long long)N, n, h->d_fL.p, d_k);
    h->launches += 1;
    TFIN_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tfin_field_sample(tfin_handle_t h, const double* z, uint64_t seed, uint32_t subsequence,
                                 int64_t first_row, int64_t N, int32_t mem,
                                 double* k_out, double* z_out, void* stream) {
    CHECK_HANDLE(h);
    if (h->f_n <= 0) return fail(TFIN_E_STATE, "tfin_field_sample: call tfin_field_set_cov or tfin_field_set_chol first");
    if (N < 0 || (N > 0 && !k_out)) return fail(TFIN_E_ARG, "tfin_field_sample: bad batch argument");
    if (N == 0) return 0;
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    Staged sg{h, st, mem == TFIN_MEM_HOST};
    const int n = h->f_n;
    const size_t cnt = (size_t)N * n;
    const double* d_z;
    double* d_k;
    if (z) {
        if (int e = sg.in(z, cnt, h->d_fz, &d_z)) return e;
    } else {
        double* d_zw;
        if (!sg.host && z_out) d_zw = z_out;
        else {
            if (int e = h->d_fz.reserve(cnt)) return e;
            d_zw = h->d_fz.p;
        }
        const long long pairs = (long long)N * ((n + 1) / 2);
        field_normal_kernel<<<(int)std::min<long long>((pairs + 255) / 256, (long long)h->sm_count * 16), 256, 0, st>>>(
            (unsigned long long)seed, subsequence, (long long)first_row, (long long)N, n, d_zw);
        h->launches += 1;
        d_z = d_zw;
    }
    if (int e = sg.out_alloc(k_out, cnt, h->d_fk, &d_k)) return e;
    if (int e = launch_field_sample(h, d_z, N, d_k, st)) return e;
    if (int e = sg.out_copy(k_out, cnt, (const double*)d_k)) return e;
    if (z_out && d_z != z_out) {
        if (sg.host) {
            if (int e = sg.out_copy(z_out, cnt, d_z)) return e;
        } else {
            TFIN_CUDA(cudaMemcpyAsync(z_out, d_z, cnt * sizeof(double), cudaMemcpyDeviceToDevice, st));
        }
    }
    if (sg.host) TFIN_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// ------------------------------------------------------------------------------------------------ many-chain pCN
extern "C" int tfin_pcn_chains(tfin_handle_t h, int32_t model, int64_t C, int64_t first_chain, int32_t n_steps,
                               int32_t first_step, double beta, const double* data, double sigma, uint64_t seed,
                               double tol, int32_t maxit, int32_t mem, double* z_state, int32_t init_from_prior,
                               double* misfit_out, int64_t* accepted_out, double* qoi_out, double* qoi_sum_out,
                               double* qoi_sq_out, double* k_sum_out, void* stream) {
    CHECK_HANDLE(h);
    if (h->f_n <= 0) return fail(TFIN_E_STATE, "tfin_pcn_chains: call tfin_field_set_cov / tfin_field_set_chol first");
    if (h->f_n != h->n) return fail(TFIN_E_STATE, "tfin_pcn_chains: the prior has %d points but the operator %d dofs", h->f_n, h->n);
    if (model != 0 && model != 1) return fail(TFIN_E_ARG, "tfin_pcn_chains: model must be 0 (nodal FOM) or 1 (averaged ROM)");
    if (model == 0 && h->n_cells <= 0) return fail(TFIN_E_STATE, "tfin_pcn_chains: model 0 needs tfin_set_cells");
    if (model == 1 && (h->n_r <= 0 || h->n_avg != h->rom_terms - 1))
        return fail(TFIN_E_STATE, "tfin_pcn_chains: model 1 needs tfin_set_rom and tfin_set_averaging");
    const int nobs = model == 0 ? h->n_obs : h->rom_obs;
    if (nobs <= 0) return fail(TFIN_E_STATE, "tfin_pcn_chains: no observation operator");
    if (C < 0 || n_steps < 0 || first_step < 0 || !(beta > 0.0 && beta <= 1.0) || !(sigma > 0.0) || !data ||
        (C > 0 && !z_state) || !(tol > 0.0) || maxit < 1)
        return fail(TFIN_E_ARG, "tfin_pcn_chains: bad argument (0 < beta <= 1, sigma > 0, z_state required)");
    if (C == 0) return 0;
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    const bool host = mem == TFIN_MEM_HOST;
    const int n = h->n;
    const size_t cn = (size_t)C * n, co = (size_t)C * nobs;
    const bool want_k = k_sum_out != nullptr;
    // device state: caller's buffers in device mode, staging otherwise
    ChainState cs{};
    if (host) {
        if (int e = h->d_cz.reserve(cn)) return e;
        cs.z = h->d_cz.p;
        if (!init_from_prior) TFIN_CUDA(cudaMemcpyAsync(cs.z, z_state, cn * 8, cudaMemcpyHostToDevice, st));
    } else {
        cs.z = z_state;
    }
    auto pick = [&](double* user, DevBuf<double>& buf, size_t count, double** out) -> int {
        if (!host && user) {
            *out = user;
            return 0;
        }
        if (int e = buf.reserve(count)) return e;
        *out = buf.p;
        return 0;
    };
    if (int e = pick(misfit_out, h->d_cphi, (size_t)C, &cs.phi)) return e;
    if (int e = pick(qoi_out, h->d_cq, co, &cs.qoi)) return e;
    if (int e = pick(qoi_sum_out, h->d_cqs, co, &cs.qoi_sum)) return e;
    if (int e = pick(qoi_sq_out, h->d_cqq, co, &cs.qoi_sq)) return e;
    if (want_k) {
        if (int e = pick(k_sum_out, h->d_cks, cn, &cs.k_sum)) return e;
        if (int e = h->d_ck.reserve(cn)) return e;
        cs.k = h->d_ck.p;
    }
    if (!host && accepted_out) cs.accepted = reinterpret_cast<unsigned long long*>(accepted_out);
    else {
        if (int e = h->d_cacc.reserve((size_t)C)) return e;
        cs.accepted = h->d_cacc.p;
    }
    if (int e = h->d_czp.reserve(cn)) return e;
    if (int e = h->d_ckp.reserve(cn)) return e;
    if (int e = h->d_cqp.reserve(co)) return e;
    if (int e = h->d_cstat.reserve((size_t)C)) return e;
    std::vector<double> hd(data, data + nobs);   // data is a set-up array: always a host pointer
    if (int e = h->d_cdata.upload(hd, st)) return e;
    TFIN_CUDA(cudaMemsetAsync(cs.accepted, 0, (size_t)C * 8, st));
    TFIN_CUDA(cudaMemsetAsync(cs.qoi_sum, 0, co * 8, st));
    TFIN_CUDA(cudaMemsetAsync(cs.qoi_sq, 0, co * 8, st));
    if (want_k) TFIN_CUDA(cudaMemsetAsync(cs.k_sum, 0, cn * 8, st));
    const int ppr = (n + 1) / 2;
    const int gp = (int)std::min<int64_t>((C * ppr + 255) / 256, (int64_t)h->sm_count * 16);
    const int ga = (int)std::min<int64_t>((C + 7) / 8, (int64_t)h->sm_count * 8);
    const double inv_s2 = 1.0 / (sigma * sigma);
    auto forward = [&](const double* d_z) -> int {   // k' = T(z), qoi' = F(k'), status'
        if (int e = launch_field_sample(h, d_z, C, h->d_ckp.p, st)) return e;
        if (model == 0)
            return fom_common(h, true, h->d_ckp.p, C, TFIN_IN_NODAL, TFIN_MEM_DEVICE, tol, maxit, nullptr, h->d_cqp.p,
                              nullptr, h->d_cstat.p, nullptr, st);
        return tfin_rom(h, h->d_ckp.p, C, TFIN_IN_NODAL, TFIN_MEM_DEVICE, nullptr, h->d_cqp.p, h->d_cstat.p, st);
    };
    // ---- start state: z from the caller or from the prior (Philox draw 0 of every chain), evaluated once
    if (init_from_prior) {
        field_normal_kernel<<<gp, 256, 0, st>>>((unsigned long long)seed, 0u, (long long)first_chain, (long long)C, n, cs.z);
        h->launches += 1;
    }
    if (int e = forward(cs.z)) return e;
    chain_accept_kernel<<<ga, 256, 0, st>>>(cs, cs.z, h->d_ckp.p, h->d_cqp.p, h->d_cstat.p, h->d_cdata.p, inv_s2,
                                           (unsigned long long)seed, 0u, (long long)first_chain, (long long)C, n, nobs, 1);
    h->launches += 1;
    // ---- steps first_step+1 .. first_step+n_steps (draw index = step number, so runs can be continued)
    for (int t = 1; t <= n_steps; ++t) {
        const uint32_t step = (uint32_t)(first_step + t);
        pcn_propose_kernel<<<gp, 256, 0, st>>>(cs.z, beta, (unsigned long long)seed, step, (long long)first_chain,
                                              (long long)C, n, h->d_czp.p);
        h->launches += 1;
        if (int e = forward(h->d_czp.p)) return e;
        chain_accept_kernel<<<ga, 256, 0, st>>>(cs, h->d_czp.p, h->d_ckp.p, h->d_cqp.p, h->d_cstat.p, h->d_cdata.p, inv_s2,
                                               (unsigned long long)seed, step, (long long)first_chain, (long long)C, n,
                                               nobs, 0);
        h->launches += 1;
    }
    TFIN_CUDA(cudaGetLastError());
    if (host) {
        auto back = [&](void* dst, const void* src, size_t bytes) -> int {
            if (dst) TFIN_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
            return 0;
        };
        if (int e = back(z_state, cs.z, cn * 8)) return e;
        if (int e = back(misfit_out, cs.phi, (size_t)C * 8)) return e;
        if (int e = back(accepted_out, cs.accepted, (size_t)C * 8)) return e;
        if (int e = back(qoi_out, cs.qoi, co * 8)) return e;
        if (int e = back(qoi_sum_out, cs.qoi_sum, co * 8)) return e;
        if (int e = back(qoi_sq_out, cs.qoi_sq, co * 8)) return e;
        if (want_k)
            if (int e = back(k_sum_out, cs.k_sum, cn * 8)) return e;
        TFIN_CUDA(cudaStreamSynchronize(st));
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------ introspection
extern "C" int64_t tfin_kernel_launches(tfin_handle_t h) { return h ? h->launches : -1; }

extern "C" int64_t tfin_get_int(tfin_handle_t h, const char* key) {
    if (!h || !key) return -1;
    const std::string k(key);
    if (k == "n") return h->n;
    if (k == "n_obs") return h->n_obs;
    if (k == "n_terms") return h->n_terms;
    if (k == "n_r") return h->n_r;
    if (k == "n_cells") return h->n_cells;
    if (k == "ell_width") return h->W;
    if (k == "ell_width_nodal") return h->Wn;
    if (k == "sm_count") return h->sm_count;
    if (k == "pcg_threads") return h->last_T;
    if (k == "pcg_rows_per_thread") return h->last_R;
    if (k == "pcg_ctas_per_sm") return h->last_occ;
    if (k == "pcg_smem_bytes") return (int64_t)h->last_smem;
    if (k == "pcg_ell_width_compiled") return h->last_WT;
    if (k == "pcg_reg_slots") return h->last_WR;
    if (k == "rom_chunk") return h->rom_chunk;
    if (k == "pcg_path") return h->last_path;
    if (k == "fom_solver") return h->last_solver;          // solver of the last forward solve: 1 PCG, 2 direct
    if (k == "frontal_kernel") return h->last_fkernel;     // 1 = D1, 2 = D2 observables mode, 3 = D2 solve mode
    if (k == "frontal_threads") return h->last_fthreads;
    if (k == "frontal_bsub_ctas_per_sm") return h->last_focc_b;
    if (k == "frontal_lanes") return h->fr_aff.ok ? h->fr_aff.streams.lanes : -1;
    if (k == "frontal_ring_rows") return h->fr_aff.ok ? h->fr_aff.streams.lr_rows : -1;
    if (k == "frontal_ctas_per_sm") return h->last_focc;
    if (k == "frontal_smem_bytes") return (int64_t)h->last_fsmem;
    if (k == "frontal_ok") return h->fr_aff.ok ? 1 : 0;
    if (k == "frontal_nodal_ok") return h->fr_nod.ok ? 1 : 0;
    if (k == "frontal_slots") return h->fr_aff.ok ? h->fr_aff.host.nslots : -1;
    if (k == "frontal_slots_lane") return h->fr_aff.ok ? h->fr_aff.host1.nslots : -1;
    if (k == "frontal_cmax") return h->fr_aff.ok ? h->fr_aff.host.cmax : -1;
    if (k == "frontal_nnz_factor") return h->fr_aff.ok ? h->fr_aff.host.nnzL : -1;
    if (k == "frontal_pair_updates") return h->fr_aff.ok ? (int64_t)h->fr_aff.host.pair_updates : -1;
    if (k == "nodal_coef_mode") return h->coef_mode;
    if (k == "pcg_precision") return h->precision;
    if (k == "stream_tile") return h->last_tile;
    if (k == "stream_ell_width") return h->s_We;
    if (k == "stream_ld") return h->s_ldr;
    if (k == "stream_bandwidth") return h->s_bandwidth;
    if (k == "stream_ring") return h->last_ring;
    if (k.rfind("stream_prof_", 0) == 0 && h->d_sprof.p) {  // stream_prof_0..3: P1, P2, P3 clocks, iterations (CTA 0)
        const int slot = k.back() - '0';
        if (slot < 0 || slot > 3) return -1;
        unsigned long long v[4];
        DeviceGuard guard(h->device);
        if (cudaStreamSynchronize(h->stream) != cudaSuccess ||
            cudaMemcpy(v, h->d_sprof.p, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess)
            return -1;
        return (int64_t)v[slot];
    }
    return -1;
}

extern "C" int tfin_set_int(tfin_handle_t h, const char* key, int64_t value) {
    if (!h || !key) return fail(TFIN_E_ARG, "tfin_set_int: bad argument");
    const std::string k(key);
    if (k == "pcg_rows_per_thread") {
        h->pcg_R = (int)value;
        return 0;
    }
    if (k == "rom_chunk") {
        h->rom_chunk = value;
        return 0;
    }
    if (k == "pcg_reg_slots") {
        h->pcg_WR = (int)value;
        return 0;
    }
    if (k == "pcg_path") {
        h->pcg_path = (int)value;
        return 0;
    }
    if (k == "fom_solver") {
        if (value < 0 || value > 2) return fail(TFIN_E_ARG, "fom_solver must be 0 (auto), 1 (PCG) or 2 (direct)");
        h->fom_solver = (int)value;
        return 0;
    }
    if (k == "frontal_kernel") {
        if (value < 0 || value > 2) return fail(TFIN_E_ARG, "frontal_kernel must be 0 (auto), 1 (D1) or 2 (D2)");
        h->frontal_kernel = (int)value;
        return 0;
    }
    if (k == "frontal_threads") {
        h->frontal_threads = (int)value;
        return 0;
    }
    if (k == "frontal_lanes" || k == "frontal_ring_rows") {
        if (k == "frontal_ring_rows") {
            if (value < 0 || value > 4096) return fail(TFIN_E_ARG, "frontal_ring_rows must be 0 (auto) .. 4096");
            h->frontal_ring_rows = (int)value;
        } else {
            if (value != 0 && (value < 4 || value > 32)) return fail(TFIN_E_ARG, "frontal_lanes must be 0 (auto) or 4..32");
            h->frontal_lanes = (int)value;
        }
        for (FrontalSet* fs : {&h->fr_aff, &h->fr_nod})   // repack the D1 streams for the new row width
            if (fs->ok) {
                DeviceGuard guard(h->device);
                if (int e = fs->upload(h->stream, h->n_obs, h->frontal_cfg())) return e;
                TFIN_CUDA(cudaStreamSynchronize(h->stream));
            }
        return 0;
    }
    if (k == "frontal_split") {
        h->frontal_split = value != 0;
        return 0;
    }
    if (k == "frontal_mode") {
        if (value < -1 || value > 1) return fail(TFIN_E_ARG, "frontal_mode must be -1 (auto), 0 (observables) or 1 (solve)");
        h->frontal_mode = (int)value;
        return 0;
    }
    if (k == "host_chunk") {
        if (value < 0) return fail(TFIN_E_ARG, "host_chunk must be >= 0");
        h->host_chunk = value;
        return 0;
    }
    if (k == "pcg_precision") {
        if (value != 64 && value != 32) return fail(TFIN_E_ARG, "pcg_precision must be 64 or 32");
        h->precision = (int)value;
        return 0;
    }
    if (k == "nodal_coef_mode") {
        if (value != 0 && value != 1) return fail(TFIN_E_ARG, "nodal_coef_mode must be 0 (k) or 1 (exp(k))");
        h->coef_mode = (int)value;
        return 0;
    }
    if (k == "stream_tile") {
        h->stream_tile = (int)value;
        return 0;
    }
    if (k == "stream_ring") {
        h->stream_ring = (int)value;
        return 0;
    }
    if (k == "stream_pad_smem") {
        h->stream_pad_smem = (int)value;
        return 0;
    }
    if (k == "stream_prof") {
        h->stream_prof = (int)value;
        return 0;
    }
    return fail(TFIN_E_ARG, "tfin_set_int: unknown key '%s'", key);
}

// ------------------------------------------------------------------------------------------------ diagnostics
// Host-only view of the symbolic phase of the direct solver (no device needed): tests interpret the program on the CPU.
struct tfin_frontal_program {
    FrontalProgram P;
};

extern "C" int tfin_frontal_analyze(int32_t n, int32_t nnz, const int32_t* row_ptr, const int32_t* col_idx, int32_t n_terms,
                                    const double* vals, const double* rhs, int32_t n_obs, const int32_t* obs_ptr,
                                    const int32_t* obs_idx, const double* obs_val, void** out) {
    return tfin_frontal_analyze_ex(n, nnz, row_ptr, col_idx, n_terms, vals, rhs, n_obs, obs_ptr, obs_idx, obs_val, 1, out);
}

extern "C" int tfin_frontal_analyze_ex(int32_t n, int32_t nnz, const int32_t* row_ptr, const int32_t* col_idx, int32_t n_terms,
                                       const double* vals, const double* rhs, int32_t n_obs, const int32_t* obs_ptr,
                                       const int32_t* obs_idx, const double* obs_val, int32_t lookahead, void** out) {
    if (n <= 0 || nnz <= 0 || !row_ptr || !col_idx || !vals || !rhs || !out || n_terms < 1)
        return fail(TFIN_E_ARG, "tfin_frontal_analyze: bad argument");
    auto* fp = new tfin_frontal_program();
    auto terms = [&](int e, std::vector<FrontalTermEntry>& o) {
        for (int t = 0; t < n_terms; ++t) {
            const double v = vals[(size_t)t * nnz + e];
            if (v != 0.0) o.push_back(FrontalTermEntry{t, v});
        }
    };
    const std::string why = frontal_build(n, row_ptr, col_idx, rhs, terms, &fp->P, lookahead != 0);
    if (!why.empty()) {
        delete fp;
        return fail(TFIN_E_STATE, "tfin_frontal_analyze: %s", why.c_str());
    }
    if (n_obs > 0 && obs_ptr && obs_idx && obs_val) frontal_set_obs(fp->P, n_obs, obs_ptr, obs_idx, obs_val);
    *out = fp;
    return 0;
}

extern "C" int64_t tfin_frontal_array(void* prog, const char* name, void* dst, int64_t dst_bytes) {
    if (!prog || !name) return -1;
    const FrontalProgram& P = static_cast<tfin_frontal_program*>(prog)->P;
    const std::string k(name);
    auto give = [&](const void* src, size_t bytes) -> int64_t {
        if (dst && (int64_t)bytes <= dst_bytes) std::memcpy(dst, src, bytes);
        return (int64_t)bytes;
    };
    if (k == "n") return P.n;
    if (k == "nslots") return P.nslots;
    if (k == "cmax") return P.cmax;
    if (k == "nnzL") return P.nnzL;
    if (k == "pair_updates") return (int64_t)P.pair_updates;
    if (k == "perm") return give(P.perm.data(), P.perm.size() * 4);
    if (k == "piv_slot") return give(P.piv_slot.data(), P.piv_slot.size() * 2);
    if (k == "col_ptr") return give(P.col_ptr.data(), P.col_ptr.size() * 4);
    if (k == "col_slot") return give(P.col_slot.data(), P.col_slot.size() * 2);
    if (k == "rhs") return give(P.rhs.data(), P.rhs.size() * 8);
    if (k == "asm_ptr") return give(P.asm_ptr.data(), P.asm_ptr.size() * 4);
    if (k == "asm_addr") return give(P.asm_addr.data(), P.asm_addr.size() * 4);
    if (k == "asm_eptr") return give(P.asm_eptr.data(), P.asm_eptr.size() * 4);
    if (k == "ent_term") return give(P.ent_term.data(), P.ent_term.size() * 4);
    if (k == "ent_coef") return give(P.ent_coef.data(), P.ent_coef.size() * 8);
    if (k == "obs_ptr") return give(P.obs_ptr.data(), P.obs_ptr.size() * 4);
    if (k == "obs_row") return give(P.obs_row.data(), P.obs_row.size() * 4);
    if (k == "obs_val") return give(P.obs_val.data(), P.obs_val.size() * 8);
    return -1;
}

extern "C" void tfin_frontal_free(void* prog) { delete static_cast<tfin_frontal_program*>(prog); }

// ------------------------------------------------------------------------------------------------ micro-benchmarks
// Shared-memory read bandwidth of the device (all SMs): the roofline denominator of the on-chip kernels (K1/K2 PCG, D1 /
// D2 frontal solver), which never touch HBM in their inner loops.  Conflict-free 8-byte loads, 16 per iteration.
__global__ void __launch_bounds__(1024) smem_bandwidth_kernel(int iters, double* sink) {
    extern __shared__ __align__(16) double sm_bw[];
    const int t = threadIdx.x, T = blockDim.x;
    for (int i = t; i < 16 * T; i += T) sm_bw[i] = 1.0 + i;
    __syncthreads();
    // 8-byte loads, the access width of the solver kernels (a warp reads one 256-byte row = 2 wavefronts); 16 distinct rows
    // per iteration through ld.volatile (ptxas hoists / merges plain ld.shared even inside asm volatile: an earlier version
    // executed a quarter of its loads, ncu smsp__inst_executed_op_shared_ld), so that nothing is merged, hoisted or dropped
    const unsigned base = smem_u32(sm_bw + t);
    double acc0 = 0.0, acc1 = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; u += 2) {
            double v0, v1;
            asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v0) : "r"(base + 8u * T * u));
            asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v1) : "r"(base + 8u * T * (u + 1)));
            acc0 += v0;
            acc1 += v1;
        }
    }
    if (acc0 + acc1 == -1.0) sink[blockIdx.x] = acc0;   // never true: keeps the loads alive
}

extern "C" int tfin_smem_bandwidth(tfin_handle_t h, double* gbs_out) {
    CHECK_HANDLE(h);
    if (!gbs_out) return fail(TFIN_E_ARG, "tfin_smem_bandwidth: gbs_out is NULL");
    const int T = 1024, iters = 20000;
    const size_t smem = (size_t)16 * T * sizeof(double);
    TFIN_CUDA(cudaFuncSetAttribute(smem_bandwidth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    TFIN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, smem_bandwidth_kernel, T, smem));
    const int grid = h->sm_count * std::max(occ, 1);
    if (int e = h->d_relres.reserve((size_t)grid)) return e;
    cudaEvent_t e0, e1;
    TFIN_CUDA(cudaEventCreate(&e0));
    TFIN_CUDA(cudaEventCreate(&e1));
    smem_bandwidth_kernel<<<grid, T, smem, h->stream>>>(200, h->d_relres.p);   // warm-up
    TFIN_CUDA(cudaEventRecord(e0, h->stream));
    smem_bandwidth_kernel<<<grid, T, smem, h->stream>>>(iters, h->d_relres.p);
    TFIN_CUDA(cudaEventRecord(e1, h->stream));
    TFIN_CUDA(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    TFIN_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    h->launches += 2;
    *gbs_out = (double)grid * T * iters * 16.0 * 8.0 / (ms * 1e-3) / 1e9;
    return 0;
}
