// Symbolic phase of the batched sparse-direct (frontal Cholesky) solver -- host only, no CUDA.
//
// The reference solves every sample with a sparse DIRECT factorisation (dolfin `solve` -> PETSc LU,
// fom/forward_solve.py:286, rom/averaged_affine_ROM.py:256).  All samples of a batch share ONE sparsity pattern, so the
// whole symbolic analysis (ordering, elimination tree, fill, storage addresses) is done once per operator here and
// compiled into a flat "program"; the kernels of frontal.cuh interpret it per sample with the numeric values formed
// in-kernel (affine terms or per-sample cell coefficients).
//
//   ordering   reverse breadth-first order from the root set (the dofs carrying the right-hand side), i.e. the
//              elimination runs from the fin tips towards the root, then an elimination-tree POSTORDER so that every
//              branch (sub-fin) is finished before the trunk continues: the active front stays as narrow as the strip
//              it sweeps (5 nodes in a sub-fin, 14-22 in the post at m = 3).
//   storage    every node of the active front owns a SLOT from its first appearance to its elimination; the Schur
//              complement of the front lives in a dense packed triangle indexed by slot, tri(hi, lo) = hi(hi+1)/2 + lo.
//              Entries of free slots are exactly zero (the pivot column is zeroed when it is consumed).
//   program    per pivot j (elimination order): its slot, the slots of its column structure SORTED BY SLOT (so that
//              position a >= b implies slot_a >= slot_b and the update address is tri(slot_a) + slot_b without min/max),
//              the right-hand side, the observation weights, and the assembly list of column j: target address +
//              (term, coefficient) pairs, value = sum coef * cvec[term] with cvec = [1, theta...] or [1, cell coefs...].
//              Column j + 1 is assembled during step j (its slots are allocated one step early), which lets the
//              multi-thread kernel overlap assembly with the gather of column j without a race.
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <functional>
#include <queue>
#include <string>
#include <vector>

namespace tfin {

struct FrontalTermEntry {
    int term;
    double coef;
};

struct FrontalProgram {
    int n = 0, nslots = 0, cmax = 0;
    long long nnzL = 0;
    double pair_updates = 0;             // sum_j c_j (c_j + 1) / 2  (FMAs of the factorisation)
    std::vector<int> perm;               // [n]     elimination position -> caller's dof
    std::vector<uint16_t> piv_slot;      // [n]
    std::vector<int> col_ptr;            // [n+1]
    std::vector<uint16_t> col_slot;      // [nnzL]  ascending within a column
    std::vector<double> rhs;             // [n]     right-hand side in elimination order
    std::vector<int> asm_ptr;            // [n+1]   -> assembly positions of column j
    std::vector<uint32_t> asm_addr;      // [npos]  packed-triangle address of the target entry
    std::vector<int> asm_eptr;           // [npos+1] -> entries
    std::vector<int> ent_term;           // [nent]
    std::vector<double> ent_coef;        // [nent]
    std::vector<int> obs_ptr;            // [n+1]   observation weights of pivot j
    std::vector<int> obs_row;
    std::vector<double> obs_val;
};

inline uint32_t frontal_tri(uint32_t a, uint32_t b) {
    const uint32_t hi = a > b ? a : b, lo = a > b ? b : a;
    return hi * (hi + 1) / 2 + lo;
}

// Reverse BFS order from `roots` (multi-source; other components start at their lowest-numbered node).
// order[pos] = node; far nodes first, roots last.  Neighbours are visited by ascending degree (Cuthill-McKee).
inline std::vector<int> frontal_rbfs(int n, const int32_t* rp, const int32_t* ci, const std::vector<int>& roots) {
    std::vector<int> order;
    order.reserve(n);
    std::vector<char> seen(n, 0);
    std::vector<int> nb;
    size_t head = 0;
    int next_seed = 0;
    auto push_roots = [&](const std::vector<int>& rs) {
        for (int r : rs)
            if (!seen[r]) {
                seen[r] = 1;
                order.push_back(r);
            }
    };
    push_roots(roots);
    while ((int)order.size() < n) {
        if (head == order.size()) {
            while (seen[next_seed]) ++next_seed;
            push_roots({next_seed});
        }
        while (head < order.size()) {
            const int u = order[head++];
            nb.clear();
            for (int j = rp[u]; j < rp[u + 1]; ++j) {
                const int v = ci[j];
                if (!seen[v]) {
                    seen[v] = 1;
                    nb.push_back(v);
                }
            }
            std::sort(nb.begin(), nb.end(), [&](int a, int b) {
                const int da = rp[a + 1] - rp[a], db = rp[b + 1] - rp[b];
                return da != db ? da < db : a < b;
            });
            order.insert(order.end(), nb.begin(), nb.end());
        }
    }
    std::reverse(order.begin(), order.end());
    return order;
}

// Elimination tree (Liu, path compression) of the pattern permuted by `order`.
inline std::vector<int> frontal_etree(int n, const int32_t* rp, const int32_t* ci, const std::vector<int>& order,
                                      const std::vector<int>& inv) {
    std::vector<int> parent(n, -1), anc(n, -1);
    for (int j = 0; j < n; ++j) {
        const int old = order[j];
        for (int e = rp[old]; e < rp[old + 1]; ++e) {
            int i = inv[ci[e]];
            while (i != -1 && i < j) {
                const int nxt = anc[i];
                anc[i] = j;
                if (nxt == -1) parent[i] = j;
                i = nxt;
            }
        }
    }
    return parent;
}

// Postorder of the elimination forest; children of a node are visited by DESCENDING subtree size when big_first (the
// trunk before the branches), ascending otherwise.  Returns post[k] = node visited k-th.
inline std::vector<int> frontal_postorder(int n, const std::vector<int>& parent, bool big_first) {
    std::vector<int> size(n, 1), head(n, -1), next(n, -1), roots;
    for (int j = 0; j < n; ++j)
        if (parent[j] >= 0) size[parent[j]] += size[j];
    std::vector<std::vector<int>> children(n);
    for (int j = 0; j < n; ++j) {
        if (parent[j] >= 0) children[parent[j]].push_back(j);
        else roots.push_back(j);
    }
    auto cmp = [&](int a, int b) {
        if (size[a] != size[b]) return big_first ? size[a] > size[b] : size[a] < size[b];
        return a < b;
    };
    std::vector<int> post;
    post.reserve(n);
    std::vector<std::pair<int, int>> stack;
    std::sort(roots.begin(), roots.end(), cmp);
    for (int r : roots) {
        stack.push_back({r, 0});
        while (!stack.empty()) {
            auto& top = stack.back();
            const int v = top.first;
            if (top.second == 0) std::sort(children[v].begin(), children[v].end(), cmp);
            if (top.second < (int)children[v].size()) {
                const int c = children[v][top.second++];
                stack.push_back({c, 0});
            } else {
                post.push_back(v);
                stack.pop_back();
            }
        }
    }
    return post;
}

// Build the program.  `entry_terms(e, out)` appends the (term, coef) pairs of CSR entry e (row i, column ci[e]); it is
// called for entries of the LOWER triangle in elimination order only (and the diagonal).  Returns an empty string on
// success, else the reason the mesh is not supported.
inline std::string frontal_build(int n, const int32_t* rp, const int32_t* ci, const double* rhs,
                                 const std::function<void(int, std::vector<FrontalTermEntry>&)>& entry_terms,
                                 FrontalProgram* out) {
    std::vector<int> roots;
    for (int i = 0; i < n; ++i)
        if (rhs[i] != 0.0) roots.push_back(i);
    const std::vector<int> base = frontal_rbfs(n, rp, ci, roots);

    struct Candidate {
        std::vector<int> order, inv;
        std::vector<int> sptr;       // [n+1] column structures (elimination indices, ascending)
        std::vector<int> sidx;
        std::vector<int> slot;       // slot of node j during its life
        std::vector<int> alloc_at;   // step at which node j gets its slot
        int nslots = 0;
    };
    auto analyse = [&](bool big_first, Candidate& c) {
        std::vector<int> inv(n);
        for (int i = 0; i < n; ++i) inv[base[i]] = i;
        std::vector<int> parent = frontal_etree(n, rp, ci, base, inv);
        const std::vector<int> post = frontal_postorder(n, parent, big_first);
        c.order.resize(n);
        for (int k = 0; k < n; ++k) c.order[k] = base[post[k]];
        c.inv.resize(n);
        for (int i = 0; i < n; ++i) c.inv[c.order[i]] = i;
        parent = frontal_etree(n, rp, ci, c.order, c.inv);
        // column structures: struct(j) = higher neighbours of j  U  (struct(child) \ {j}) over the children of j
        std::vector<std::vector<int>> children(n);
        for (int j = 0; j < n; ++j)
            if (parent[j] >= 0) children[parent[j]].push_back(j);
        c.sptr.assign(n + 1, 0);
        c.sidx.clear();
        std::vector<int> mark(n, -1);
        for (int j = 0; j < n; ++j) {
            const size_t begin = c.sidx.size();
            mark[j] = j;
            const int old = c.order[j];
            for (int e = rp[old]; e < rp[old + 1]; ++e) {
                const int i = c.inv[ci[e]];
                if (i > j && mark[i] != j) {
                    mark[i] = j;
                    c.sidx.push_back(i);
                }
            }
            for (int ch : children[j])
                for (int q = c.sptr[ch]; q < c.sptr[ch + 1]; ++q) {
                    const int i = c.sidx[q];
                    if (mark[i] != j) {
                        mark[i] = j;
                        c.sidx.push_back(i);
                    }
                }
            std::sort(c.sidx.begin() + begin, c.sidx.end());
            c.sptr[j + 1] = (int)c.sidx.size();
        }
        // slots: a node is live from its first appearance (in a column structure, or one step before its own
        // elimination together with its higher neighbours, for the early assembly) to the end of its own step
        c.slot.assign(n, -1);
        c.alloc_at.assign(n, -1);
        std::priority_queue<int, std::vector<int>, std::greater<int>> free_slots;
        int top = 0;
        auto ensure = [&](int v, int step) {
            if (c.slot[v] >= 0) return;
            if (!free_slots.empty()) {
                c.slot[v] = free_slots.top();
                free_slots.pop();
            } else {
                c.slot[v] = top++;
            }
            c.alloc_at[v] = step;
        };
        auto ensure_column = [&](int j, int step) {
            ensure(j, step);
            const int old = c.order[j];
            for (int e = rp[old]; e < rp[old + 1]; ++e) {
                const int i = c.inv[ci[e]];
                if (i > j) ensure(i, step);
            }
        };
        ensure_column(0, -1);
        for (int j = 0; j < n; ++j) {
            for (int q = c.sptr[j]; q < c.sptr[j + 1]; ++q) ensure(c.sidx[q], j);
            if (j + 1 < n) ensure_column(j + 1, j);
            free_slots.push(c.slot[j]);
        }
        c.nslots = top;
    };
    Candidate a, b;
    analyse(true, a);
    analyse(false, b);
    Candidate& c = b.nslots < a.nslots ? b : a;
    if (c.nslots > 65535) return "frontal solver: more than 65535 live front nodes";

    FrontalProgram& P = *out;
    P = FrontalProgram();
    P.n = n;
    P.nslots = c.nslots;
    P.perm = c.order;
    P.piv_slot.resize(n);
    P.col_ptr.assign(n + 1, 0);
    P.col_slot.resize(c.sidx.size());
    P.rhs.resize(n);
    P.nnzL = (long long)c.sidx.size();
    std::vector<int> tmp;
    for (int j = 0; j < n; ++j) {
        P.piv_slot[j] = (uint16_t)c.slot[j];
        P.rhs[j] = rhs[c.order[j]];
        tmp.clear();
        for (int q = c.sptr[j]; q < c.sptr[j + 1]; ++q) tmp.push_back(c.slot[c.sidx[q]]);
        std::sort(tmp.begin(), tmp.end());
        for (size_t q = 0; q < tmp.size(); ++q) P.col_slot[c.sptr[j] + q] = (uint16_t)tmp[q];
        P.col_ptr[j + 1] = c.sptr[j + 1];
        const int cj = c.sptr[j + 1] - c.sptr[j];
        P.cmax = std::max(P.cmax, cj);
        P.pair_updates += 0.5 * cj * (cj + 1.0);
    }
    // assembly lists: diagonal first, then the higher neighbours of the column
    P.asm_ptr.assign(n + 1, 0);
    P.asm_eptr.push_back(0);
    std::vector<FrontalTermEntry> terms;
    for (int j = 0; j < n; ++j) {
        const int old = c.order[j];
        for (int pass = 0; pass < 2; ++pass)
            for (int e = rp[old]; e < rp[old + 1]; ++e) {
                const int i = c.inv[ci[e]];
                if ((pass == 0) != (i == j) || i < j) continue;
                terms.clear();
                entry_terms(e, terms);
                if (terms.empty() && i != j) continue;
                P.asm_addr.push_back(frontal_tri((uint32_t)c.slot[i], (uint32_t)c.slot[j]));
                for (const FrontalTermEntry& t : terms) {
                    P.ent_term.push_back(t.term);
                    P.ent_coef.push_back(t.coef);
                }
                P.asm_eptr.push_back((int)P.ent_term.size());
            }
        P.asm_ptr[j + 1] = (int)P.asm_addr.size();
    }
    P.obs_ptr.assign(n + 1, 0);
    return std::string();
}

// Observation operator B_obs (CSR over the caller's dofs) -> weights per pivot.
inline void frontal_set_obs(FrontalProgram& P, int n_obs, const int32_t* obs_ptr, const int32_t* obs_idx,
                            const double* obs_val) {
    const int n = P.n;
    std::vector<int> inv(n);
    for (int i = 0; i < n; ++i) inv[P.perm[i]] = i;
    std::vector<std::vector<std::pair<int, double>>> per(n);
    for (int o = 0; o < n_obs; ++o)
        for (int q = obs_ptr[o]; q < obs_ptr[o + 1]; ++q) per[inv[obs_idx[q]]].push_back({o, obs_val[q]});
    P.obs_ptr.assign(n + 1, 0);
    P.obs_row.clear();
    P.obs_val.clear();
    for (int j = 0; j < n; ++j) {
        for (auto& pr : per[j]) {
            P.obs_row.push_back(pr.first);
            P.obs_val.push_back(pr.second);
        }
        P.obs_ptr[j + 1] = (int)P.obs_row.size();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Packed instruction streams.  The kernels never chase the program's CSR arrays (every dependent global load would cost
// a full L2 round trip per pivot with one or two resident warps): the program is flattened into one byte stream per
// direction, read strictly sequentially through a small shared-memory ring that cp.async keeps filled ahead of use.
//
// forward record (prologue j = -1 with c = 0, then j = 0 .. n-1), 16-byte aligned, little endian:
//    +0 u32 c | +4 u32 pivot slot | +8 u32 npos | +12 u32 nent | +16 u32 nobs | +20 u32 record bytes | +24 f64 rhs_j
//    +32 c x u16 slots (padded to 8) | npos x {u32 address, u32 entries} | nent x f64 coef | nent x u32 term (padded to 8)
//    | nobs x f64 weight | nobs x u32 row (padded to 8) | pad to 16
//    (npos / nent describe the assembly of column j + 1, which step j performs; nobs the observation weights of pivot j)
// backward record (j = n-1 .. 0):
//    +0 u32 c | +4 u32 pivot slot | +8 u32 nobs | +12 u32 dof (perm_j) | +16 u32 record bytes | +20 u32 c of the NEXT record
//    +24 f64 rhs_j | +32 slots | weights | rows | pad to 16
struct FrontalStreams {
    std::vector<unsigned char> fwd, bwd;
    int max_record = 0;   // largest record of either stream (bytes)
    int ring_bytes = 0;   // power of two >= 2 * max_record + 512
};

inline void frontal_pack_streams(const FrontalProgram& P, FrontalStreams* out) {
    FrontalStreams& S = *out;
    S = FrontalStreams();
    auto put32 = [](std::vector<unsigned char>& v, uint32_t x) {
        for (int k = 0; k < 4; ++k) v.push_back((unsigned char)(x >> (8 * k)));
    };
    auto put16 = [](std::vector<unsigned char>& v, uint16_t x) {
        v.push_back((unsigned char)(x & 255));
        v.push_back((unsigned char)(x >> 8));
    };
    auto put64 = [](std::vector<unsigned char>& v, double d) {
        unsigned char b[8];
        std::memcpy(b, &d, 8);
        v.insert(v.end(), b, b + 8);
    };
    auto pad = [](std::vector<unsigned char>& v, size_t a) {
        while (v.size() % a) v.push_back(0);
    };
    auto set32 = [](std::vector<unsigned char>& v, size_t at, uint32_t x) {
        for (int k = 0; k < 4; ++k) v[at + k] = (unsigned char)(x >> (8 * k));
    };
    const int n = P.n;
    for (int j = -1; j < n; ++j) {
        const size_t start = S.fwd.size();
        const int c = j >= 0 ? P.col_ptr[j + 1] - P.col_ptr[j] : 0;
        const int q0 = j + 1 < n ? P.asm_ptr[j + 1] : 0, q1 = j + 1 < n ? P.asm_ptr[j + 2] : 0;
        const int e0 = q1 > q0 ? P.asm_eptr[q0] : 0, e1 = q1 > q0 ? P.asm_eptr[q1] : 0;
        const int o0 = j >= 0 ? P.obs_ptr[j] : 0, o1 = j >= 0 ? P.obs_ptr[j + 1] : 0;
        put32(S.fwd, (uint32_t)c);
        put32(S.fwd, j >= 0 ? P.piv_slot[j] : 0u);
        put32(S.fwd, (uint32_t)(q1 - q0));
        put32(S.fwd, (uint32_t)(e1 - e0));
        put32(S.fwd, (uint32_t)(o1 - o0));
        put32(S.fwd, 0u);
        put64(S.fwd, j >= 0 ? P.rhs[j] : 0.0);
        for (int a = 0; a < c; ++a) put16(S.fwd, P.col_slot[P.col_ptr[j] + a]);
        pad(S.fwd, 8);
        for (int q = q0; q < q1; ++q) {
            put32(S.fwd, P.asm_addr[q]);
            put32(S.fwd, (uint32_t)(P.asm_eptr[q + 1] - P.asm_eptr[q]));
        }
        for (int e = e0; e < e1; ++e) put64(S.fwd, P.ent_coef[e]);
        for (int e = e0; e < e1; ++e) put32(S.fwd, (uint32_t)P.ent_term[e]);
        pad(S.fwd, 8);
        for (int o = o0; o < o1; ++o) put64(S.fwd, P.obs_val[o]);
        for (int o = o0; o < o1; ++o) put32(S.fwd, (uint32_t)P.obs_row[o]);
        pad(S.fwd, 16);
        const size_t len = S.fwd.size() - start;
        set32(S.fwd, start + 20, (uint32_t)len);
        S.max_record = std::max(S.max_record, (int)len);
    }
    for (int j = n - 1; j >= 0; --j) {
        const size_t start = S.bwd.size();
        const int c = P.col_ptr[j + 1] - P.col_ptr[j];
        const int cn = j > 0 ? P.col_ptr[j] - P.col_ptr[j - 1] : 0;
        const int o0 = P.obs_ptr[j], o1 = P.obs_ptr[j + 1];
        put32(S.bwd, (uint32_t)c);
        put32(S.bwd, P.piv_slot[j]);
        put32(S.bwd, (uint32_t)(o1 - o0));
        put32(S.bwd, (uint32_t)P.perm[j]);
        put32(S.bwd, 0u);
        put32(S.bwd, (uint32_t)cn);
        put64(S.bwd, P.rhs[j]);
        for (int a = 0; a < c; ++a) put16(S.bwd, P.col_slot[P.col_ptr[j] + a]);
        pad(S.bwd, 8);
        for (int o = o0; o < o1; ++o) put64(S.bwd, P.obs_val[o]);
        for (int o = o0; o < o1; ++o) put32(S.bwd, (uint32_t)P.obs_row[o]);
        pad(S.bwd, 16);
        const size_t len = S.bwd.size() - start;
        set32(S.bwd, start + 16, (uint32_t)len);
        S.max_record = std::max(S.max_record, (int)len);
    }
    int ring = 1024;
    while (ring < 2 * S.max_record + 512) ring *= 2;
    S.ring_bytes = ring;
    // the ring loader fetches whole 512-byte chunks up to one ring ahead of the reader: zero padding behind the streams
    S.fwd.resize((S.fwd.size() + 511) / 512 * 512 + (size_t)ring + 512, 0);
    S.bwd.resize((S.bwd.size() + 511) / 512 * 512 + (size_t)ring + 512, 0);
}

}  // namespace tfin
