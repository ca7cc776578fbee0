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
                                 FrontalProgram* out, bool lookahead = true) {
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
            // lookahead: column j + 1 gets its slots while pivot j still holds its own (the sample-per-CTA kernel assembles
            // it concurrently with the update of step j).  Without: the pivot's slot is recycled first -- legal for a
            // kernel that assembles column j + 1 after it has gathered (and zeroed) column j, and one slot smaller.
            if (lookahead) {
                if (j + 1 < n) ensure_column(j + 1, j);
                free_slots.push(c.slot[j]);
            } else {
                free_slots.push(c.slot[j]);
                if (j + 1 < n) ensure_column(j + 1, j);
            }
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
// A record never straddles the wrap point of the ring (the previous record is padded instead), so a kernel addresses the
// fields of a record from ONE base pointer.
//
// D2 (sample per CTA) streams
//   forward record (prologue j = -1 with c = 0, then j = 0 .. n-1), 16-byte aligned, little endian:
//      +0 u32 c | +4 u32 pivot slot | +8 u32 npos | +12 u32 nent | +16 u32 nobs | +20 u32 record bytes | +24 f64 rhs_j
//      +32 c x u16 slots (padded to 8) | npos x {u32 address, u32 entries | first entry << 8} | nent x f64 coef | nent x u32 term (padded
//      to 8) | nobs x f64 weight | nobs x u32 row (padded to 8) | pad to 16
//      (npos / nent describe the assembly of column j + 1, which step j performs; nobs the observation weights of pivot j)
//   backward record (a prologue with c = 0, then j = n-1 .. 0):
//      +0 u32 c | +4 u32 pivot slot | +8 u32 nobs | +12 u32 dof (perm_j) | +16 u32 record bytes | +20 u32 c of the NEXT
//      record | +24 f64 rhs_j | +32 pad | +40 slots | weights | rows | pad to 16
// D1 (sample per thread) streams: everything the kernel would otherwise compute per use is stored pre-scaled to BYTE offsets
// of its [entry][lane] shared-memory layout (one row = 8 * lanes bytes; written 256 below for lanes = 32), in 16-byte
// groups that one LDS.128 fetches
//   forward record:
//      +0 u32 c | +4 u32 256 p | +8 u32 256 (tri(p) + p) | +12 u32 npos | +16 u32 nent | +20 u32 record bytes | +24 f64 rhs_j
//      +32 u32 gather[c4] (256 x address of entry (s_a, p)) | u32 row[c4] (256 tri(s_a)) | u32 col[c4] (256 s_a), c4 = c
//      rounded up to 4 | npos x {u32 256 address, u32 entries} padded to 16 | nent x f64 coef padded to 16 | nent x u32
//      256 term padded to 16
//   backward record (prologue, then j = n-1 .. 0):
//      +0 u32 c | +4 u32 256 p | +8 u32 nobs | +12 u32 dof | +16 u32 record bytes | +20 u32 npf | kw << 16 | +24 f64 rhs_j
//      +32 u32 256 x first ring row of this pivot's factor block | +36 u32 workspace row of that block | +48 npf x {u32 256 ring row, u32 rows, u32
//      source row, pad} (blocks to request now) | u32 col[c4] | nobs x f64 weight padded to 16 | nobs x u32 256 row pad 16
struct FrontalStreams {
    std::vector<unsigned char> fwd, bwd;     // D2
    std::vector<unsigned char> fwd1, bwd1;   // D1 (empty if the front is too wide for it)
    std::vector<unsigned char> fsub1;        // D1: forward substitution with a new right-hand side (adjoint solves)
    int max_record = 0;   // largest record of any stream (bytes)
    int ring_bytes = 0;   // power of two
    int ring_fwd1 = 0;    // D1 factor kernel alone: ring of its forward stream (power of two <= ring_bytes)
    int lr_rows = 0;      // rows of D1's factor-row ring the backward schedule was simulated for
    int lanes = 32;       // samples per warp the D1 streams were scaled for (row = 8 * lanes bytes)
};

namespace frontal_detail {
struct ByteStream {
    std::vector<unsigned char> v;
    std::vector<size_t> starts;   // record starts
    void put32(uint32_t x) {
        for (int k = 0; k < 4; ++k) v.push_back((unsigned char)(x >> (8 * k)));
    }
    void put16(uint16_t x) {
        v.push_back((unsigned char)(x & 255));
        v.push_back((unsigned char)(x >> 8));
    }
    void put64(double d) {
        unsigned char b[8];
        std::memcpy(b, &d, 8);
        v.insert(v.end(), b, b + 8);
    }
    void pad(size_t a) {
        while (v.size() % a) v.push_back(0);
    }
    void set32(size_t at, uint32_t x) {
        for (int k = 0; k < 4; ++k) v[at + k] = (unsigned char)(x >> (8 * k));
    }
    void begin() { starts.push_back(v.size()); }
};
// Re-lay the records so that none straddles a multiple of `ring` (pad the previous record, whose length field sits at
// `len_off`), then append the zero padding the ring loader may prefetch beyond the last record.
inline std::vector<unsigned char> finish_stream(const ByteStream& in, size_t len_off, int ring, int tail = 0) {
    std::vector<unsigned char> out;
    size_t prev = (size_t)-1;
    for (size_t r = 0; r < in.starts.size(); ++r) {
        const size_t b = in.starts[r], e = r + 1 < in.starts.size() ? in.starts[r + 1] : in.v.size();
        const size_t len = e - b;
        if (out.size() % ring + len > (size_t)ring && prev != (size_t)-1) {
            const size_t target = (out.size() / ring + 1) * ring;
            out.resize(target, 0);
            uint32_t plen = (uint32_t)(target - prev);
            for (int k = 0; k < 4; ++k) out[prev + len_off + k] = (unsigned char)(plen >> (8 * k));
        }
        prev = out.size();
        out.insert(out.end(), in.v.begin() + b, in.v.begin() + e);
        uint32_t l32 = (uint32_t)len;
        for (int k = 0; k < 4; ++k) out[prev + len_off + k] = (unsigned char)(l32 >> (8 * k));
    }
    out.resize((out.size() + 511) / 512 * 512 + (size_t)std::max(ring, tail) + 512, 0);
    return out;
}
}  // namespace frontal_detail

// P1: the program the sample-per-thread streams are packed from (built without lookahead: fewer slots), default P.
inline void frontal_pack_streams(const FrontalProgram& P, int lr_rows, int dmax, int lanes, FrontalStreams* out,
                                 const FrontalProgram* P1 = nullptr) {
    using frontal_detail::ByteStream;
    FrontalStreams& S = *out;
    S = FrontalStreams();
    const int n = P.n;
    const FrontalProgram& PD2 = P;
    const bool lane_ok = lr_rows >= P.cmax + 2 && P.cmax <= 32;
    S.lr_rows = lane_ok ? lr_rows : 0;
    S.lanes = lanes;
    auto tri = [](uint32_t s) { return s * (s + 1) / 2; };
    auto max_len = [](const ByteStream& b) {
        size_t m = 0;
        for (size_t r = 0; r < b.starts.size(); ++r)
            m = std::max(m, (r + 1 < b.starts.size() ? b.starts[r + 1] : b.v.size()) - b.starts[r]);
        return (int)m;
    };
    // ------------------------------------------------------------------ D2 forward / backward
    ByteStream f2, b2;
    for (int j = -1; j < n; ++j) {
        f2.begin();
        const int c = j >= 0 ? P.col_ptr[j + 1] - P.col_ptr[j] : 0;
        const int q0 = j + 1 < n ? P.asm_ptr[j + 1] : 0, q1 = j + 1 < n ? P.asm_ptr[j + 2] : 0;
        const int e0 = q1 > q0 ? P.asm_eptr[q0] : 0, e1 = q1 > q0 ? P.asm_eptr[q1] : 0;
        const int o0 = j >= 0 ? P.obs_ptr[j] : 0, o1 = j >= 0 ? P.obs_ptr[j + 1] : 0;
        f2.put32((uint32_t)c);
        f2.put32(j >= 0 ? P.piv_slot[j] : 0u);
        f2.put32((uint32_t)(q1 - q0));
        f2.put32((uint32_t)(e1 - e0));
        f2.put32((uint32_t)(o1 - o0));
        f2.put32(0u);
        f2.put64(j >= 0 ? P.rhs[j] : 0.0);
        for (int a = 0; a < c; ++a) f2.put16(P.col_slot[P.col_ptr[j] + a]);
        f2.pad(8);
        for (int q = q0; q < q1; ++q) {
            f2.put32(P.asm_addr[q]);
            f2.put32((uint32_t)(P.asm_eptr[q + 1] - P.asm_eptr[q]) | ((uint32_t)(P.asm_eptr[q] - e0) << 8));   // count | first entry << 8
        }
        for (int e = e0; e < e1; ++e) f2.put64(P.ent_coef[e]);
        for (int e = e0; e < e1; ++e) f2.put32((uint32_t)P.ent_term[e]);
        f2.pad(8);
        for (int o = o0; o < o1; ++o) f2.put64(P.obs_val[o]);
        for (int o = o0; o < o1; ++o) f2.put32((uint32_t)P.obs_row[o]);
        f2.pad(16);
    }
    for (int j = n; j >= 0; --j) {   // j == n: prologue record
        b2.begin();
        const bool pro = j == n;
        const int c = pro ? 0 : P.col_ptr[j + 1] - P.col_ptr[j];
        const int cn = (!pro && j > 0) ? P.col_ptr[j] - P.col_ptr[j - 1] : (pro ? P.col_ptr[n] - P.col_ptr[n - 1] : 0);
        const int o0 = pro ? 0 : P.obs_ptr[j], o1 = pro ? 0 : P.obs_ptr[j + 1];
        b2.put32((uint32_t)c);
        b2.put32(pro ? 0u : P.piv_slot[j]);
        b2.put32((uint32_t)(o1 - o0));
        b2.put32(pro ? 0u : (uint32_t)P.perm[j]);
        b2.put32(0u);
        b2.put32((uint32_t)cn);
        b2.put64(pro ? 0.0 : P.rhs[j]);
        b2.put64(0.0);
        for (int a = 0; a < c; ++a) b2.put16(P.col_slot[P.col_ptr[j] + a]);
        b2.pad(8);
        for (int o = o0; o < o1; ++o) b2.put64(P.obs_val[o]);
        for (int o = o0; o < o1; ++o) b2.put32((uint32_t)P.obs_row[o]);
        b2.pad(16);
    }
    // ------------------------------------------------------------------ D1 forward / backward
    ByteStream f1, b1, s1;
    const uint32_t rb = 8u * (uint32_t)lanes;   // bytes of one [row][lane] row: D1 keeps `lanes` samples per warp
    if (lane_ok) {
        const FrontalProgram& P = P1 ? *P1 : PD2;   // shadows the D2 program for the rest of this block
        for (int j = -1; j < n; ++j) {
            f1.begin();
            const int c = j >= 0 ? P.col_ptr[j + 1] - P.col_ptr[j] : 0, c4 = (c + 3) & ~3;
            const uint32_t p = j >= 0 ? P.piv_slot[j] : 0u;
            const int q0 = j + 1 < n ? P.asm_ptr[j + 1] : 0, q1 = j + 1 < n ? P.asm_ptr[j + 2] : 0;
            const int e0 = q1 > q0 ? P.asm_eptr[q0] : 0, e1 = q1 > q0 ? P.asm_eptr[q1] : 0;
            f1.put32((uint32_t)c);
            f1.put32(rb * p);
            f1.put32(rb * (tri(p) + p));
            f1.put32((uint32_t)(q1 - q0));
            f1.put32((uint32_t)(e1 - e0));
            f1.put32(0u);
            f1.put64(j >= 0 ? P.rhs[j] : 0.0);
            for (int a = 0; a < c4; ++a) {
                const uint32_t sa = a < c ? P.col_slot[P.col_ptr[j] + a] : 0u;
                f1.put32(a < c ? rb * frontal_tri(sa, p) : 0u);
            }
            for (int a = 0; a < c4; ++a) f1.put32(a < c ? rb * tri(P.col_slot[P.col_ptr[j] + a]) : 0u);
            for (int a = 0; a < c4; ++a) f1.put32(a < c ? rb * P.col_slot[P.col_ptr[j] + a] : 0u);
            for (int q = q0; q < q1; ++q) {
                f1.put32(rb * P.asm_addr[q]);
                f1.put32((uint32_t)(P.asm_eptr[q + 1] - P.asm_eptr[q]));
            }
            f1.pad(16);
            for (int e = e0; e < e1; ++e) f1.put64(P.ent_coef[e]);
            f1.pad(16);
            for (int e = e0; e < e1; ++e) f1.put32(rb * (uint32_t)P.ent_term[e]);
            f1.pad(16);
        }
        // Substitution streams (backward: pivots n-1 .. 0; forward: 0 .. n-1, for adjoint right-hand sides).  The factor
        // blocks come back from HBM through a block ring in shared memory: simulate it so that every record says which
        // blocks to request (up to dmax pivots ahead, as many as fit) and how many of the youngest copy groups may still
        // be pending when it is consumed.  Step t handles pivot piv(t); its block [1/L_jj, y_j, column] has c + 2 rows and
        // is contiguous in the ring (the tail is skipped when it does not fit).  One copy group is committed per step
        // (group 0 = prologue).
        auto pack_sub = [&](bool backward, ByteStream& out) {
            struct Blk {
                int row, rows;
            };
            struct Req {
                uint32_t row, rows, src;
            };
            std::vector<Blk> blk(n);
            std::vector<std::vector<Req>> reqs(n + 1);   // reqs[0] = prologue, reqs[t + 1] = step t
            std::vector<int> grp(n, -1);
            std::vector<uint32_t> kw(n, 0);
            auto piv = [&](int t) { return backward ? n - 1 - t : t; };
            auto rows_of = [&](int t) { const int jj = piv(t); return P.col_ptr[jj + 1] - P.col_ptr[jj] + 2; };
            auto src_of = [&](int t) { const int jj = piv(t); return (uint32_t)(P.col_ptr[jj] + 2 * jj); };
            int f = 0, head = 0, oldest = 0;   // frontier step, next free ring row, oldest live step
            auto try_alloc = [&](int t, int g) -> bool {
                const int rows = rows_of(t);
                int cand = head + rows <= lr_rows ? head : 0;
                for (int pass = 0; pass < 2; ++pass) {
                    bool clash = false;
                    for (int u = oldest; u < t; ++u)
                        if (cand < blk[u].row + blk[u].rows && blk[u].row < cand + rows) clash = true;
                    if (!clash) {
                        blk[t] = Blk{cand, rows};
                        head = cand + rows;
                        grp[t] = g;
                        reqs[g].push_back(Req{(uint32_t)cand, (uint32_t)rows, src_of(t)});
                        return true;
                    }
                    if (cand == 0) break;
                    cand = 0;   // also try the start of the ring
                }
                return false;
            };
            while (f < n && f < dmax && try_alloc(f, 0)) ++f;
            for (int t = 0; t < n; ++t) {
                oldest = t;
                while (f < n && f <= t + dmax && try_alloc(f, t + 1)) ++f;
                kw[t] = (uint32_t)(t + 1 - grp[t]);   // grp[t] >= 0: once the ring has drained a block of <= lr_rows rows fits
            }
            for (int t = -1; t < n; ++t) {   // t == -1: prologue record (initial requests only)
                out.begin();
                const bool pro = t < 0;
                const int j = pro ? 0 : piv(t);
                const int c = pro ? 0 : P.col_ptr[j + 1] - P.col_ptr[j], c4 = (c + 3) & ~3;
                const int o0 = pro ? 0 : P.obs_ptr[j], o1 = pro ? 0 : P.obs_ptr[j + 1];
                const std::vector<Req>& rq = reqs[t + 1];
                out.put32((uint32_t)c);
                out.put32(pro ? 0u : rb * P.piv_slot[j]);
                out.put32((uint32_t)(o1 - o0));
                out.put32(pro ? 0u : (uint32_t)P.perm[j]);
                out.put32(0u);
                out.put32((uint32_t)rq.size() | ((pro ? 0u : kw[t]) << 16));
                out.put64(pro ? 0.0 : P.rhs[j]);
                out.put32(pro ? 0u : rb * (uint32_t)blk[t].row);
                out.put32(pro ? 0u : src_of(t));   // workspace row of this pivot's block (the forward pass rewrites y_j)
                out.put64(0.0);
                for (const Req& r : rq) {
                    out.put32(rb * r.row);
                    out.put32(r.rows);
                    out.put32(r.src);
                    out.put32(0u);
                }
                for (int a = 0; a < c4; ++a) out.put32(a < c ? rb * P.col_slot[P.col_ptr[j] + a] : 0u);
                for (int o = o0; o < o1; ++o) out.put64(P.obs_val[o]);
                out.pad(16);
                for (int o = o0; o < o1; ++o) out.put32(rb * (uint32_t)P.obs_row[o]);
                out.pad(16);
            }
        };
        pack_sub(true, b1);
        pack_sub(false, s1);
    }
    S.max_record = std::max(std::max(max_len(f2), max_len(b2)), std::max(max_len(f1), std::max(max_len(b1), max_len(s1))));
    // ring: the reader needs the current record complete while the loader runs up to one ring ahead in 512-byte chunks;
    // the records consumed under a wait that leaves dmax + 1 copy groups pending must be older than those groups
    int ring = 2048;
    while (ring < 3 * S.max_record + 1024 || ring < (dmax + 4) * std::max(max_len(b1), max_len(b2)) + 2048) ring *= 2;
    S.ring_bytes = ring;
    S.fwd = frontal_detail::finish_stream(f2, 20, ring);
    S.bwd = frontal_detail::finish_stream(b2, 16, ring);
    if (lane_ok) {
        // the factor kernel of the split launch keeps nothing but the front and this ring in shared memory: size its ring
        // for its own records (a multiple-of-ring_fwd1 boundary is never straddled, so the larger ring reads it as well)
        int rf = 2048;
        while (rf < 3 * max_len(f1) + 1024) rf *= 2;
        S.ring_fwd1 = std::min(rf, ring);
        S.fwd1 = frontal_detail::finish_stream(f1, 20, S.ring_fwd1, ring);
        S.bwd1 = frontal_detail::finish_stream(b1, 16, ring);
        S.fsub1 = frontal_detail::finish_stream(s1, 16, ring);
    }
}

}  // namespace tfin
