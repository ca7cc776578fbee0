"""Prototype (numpy, CPU) of the frontal symbolic analysis that csrc/frontal_host.h implements in C++:
ordering (reverse BFS from the root set -> elimination tree -> postorder), column structures, slot allocation and
the per-pivot "program"; plus an interpreter of that program.  Used to size fronts / fill for the meshes of the
bench and to check the algorithm against the oracle before writing the CUDA interpreter.  Not product code."""
import sys
import time

import numpy as np


def bfs_order(n, rp, ci, roots):
    deg = np.diff(rp)
    seen = np.zeros(n, bool)
    order = []
    comp_roots = list(roots)
    ptr = 0
    while len(order) < n:
        if not comp_roots:
            comp_roots = [int(np.flatnonzero(~seen)[0])]
        q = []
        for r in comp_roots:
            if not seen[r]:
                seen[r] = True
                q.append(r)
        comp_roots = []
        order.extend(q)
        while ptr < len(order):
            u = order[ptr]
            ptr += 1
            nb = [v for v in ci[rp[u]:rp[u + 1]] if not seen[v]]
            nb.sort(key=lambda v: (deg[v], v))
            for v in nb:
                seen[v] = True
            order.extend(nb)
    return np.array(order[::-1])          # far nodes first, roots last


def etree(n, lower_adj):
    """lower_adj[j] = neighbours i < j (in the permuted numbering).  Liu's algorithm with path compression."""
    parent = -np.ones(n, np.int64)
    anc = -np.ones(n, np.int64)
    for j in range(n):
        for i in lower_adj[j]:
            while i != -1 and i < j:
                nxt = anc[i]
                anc[i] = j
                if nxt == -1:
                    parent[i] = j
                i = nxt
    return parent


def postorder(n, parent, key):
    """Postorder of the forest; children visited in ascending `key` (so the LAST child is the one eliminated just
    before the parent)."""
    children = [[] for _ in range(n)]
    roots = []
    for j in range(n):
        (children[parent[j]] if parent[j] >= 0 else roots).append(j)
    out = []
    for r in sorted(roots, key=lambda v: key[v]):
        stack = [(r, 0)]
        while stack:
            v, idx = stack.pop()
            ch = children[v]
            if idx == 0:
                ch.sort(key=lambda c: key[c])
            if idx < len(ch):
                stack.append((v, idx + 1))
                stack.append((ch[idx], 0))
            else:
                out.append(v)
    return np.array(out)


def analyse(n, rp, ci, roots, child_key="size"):
    t0 = time.time()
    order = bfs_order(n, rp, ci, roots)           # order[new] = old
    for _ in range(2):
        inv = np.empty(n, np.int64)
        inv[order] = np.arange(n)
        lower = [[] for _ in range(n)]
        for new in range(n):
            old = order[new]
            for v in ci[rp[old]:rp[old + 1]]:
                w = inv[v]
                if w < new:
                    lower[new].append(w)
        parent = etree(n, lower)
        if _ == 1:
            break
        # subtree sizes -> children with the larger subtree first or last
        size = np.ones(n, np.int64)
        for j in range(n):
            if parent[j] >= 0:
                size[parent[j]] += size[j]
        key = size if child_key == "size" else -size
        po = postorder(n, parent, key)
        order = order[po]
    # column structures (ascending elimination order)
    higher = [[] for _ in range(n)]
    for new in range(n):
        for w in lower[new]:
            higher[w].append(new)
    struct = [None] * n
    children = [[] for _ in range(n)]
    for j in range(n):
        if parent[j] >= 0:
            children[parent[j]].append(j)
    for j in range(n):
        s = set(higher[j])
        for c in children[j]:
            s |= struct[c]
        s.discard(j)
        struct[j] = s
    # order inside a column: A-neighbours first, then fill-only
    cols = []
    for j in range(n):
        a = sorted(higher[j])
        f = sorted(struct[j] - set(a))
        cols.append((a, f))
    # slot allocation
    slot = -np.ones(n, np.int64)
    free = []
    nslots = 0
    live = 0
    fmax = 0

    def alloc(v):
        nonlocal nslots, live, fmax
        if free:
            free.sort(reverse=True)
            slot[v] = free.pop()
        else:
            slot[v] = nslots
            nslots += 1
        live += 1
        fmax = max(fmax, live)

    piv_slot = np.empty(n, np.int64)
    col_slots = []
    for j in range(n):
        if slot[j] < 0:
            alloc(j)
        a, f = cols[j]
        for v in a + f:
            if slot[v] < 0:
                alloc(v)
        piv_slot[j] = slot[j]
        col_slots.append([slot[v] for v in a + f])
        free.append(slot[j])
        live -= 1
    cnt = np.array([len(a) + len(f) for a, f in cols])
    stats = dict(n=n, nnzL=int(cnt.sum()), cmax=int(cnt.max()), nslots=int(nslots), fmax=int(fmax),
                 flops_factor=int((cnt * (cnt + 1) // 2).sum()), t=time.time() - t0)
    return dict(order=order, inv=inv, cols=cols, piv_slot=piv_slot, col_slots=col_slots, nslots=nslots, stats=stats,
                parent=parent)


def tri(a, b):
    hi, lo = (a, b) if a >= b else (b, a)
    return hi * (hi + 1) // 2 + lo


def run_program(sym, A, b):
    """Interpret the program for one sample.  A: scipy CSR (full symmetric), b: rhs.  Returns w."""
    n = len(b)
    order, inv = sym["order"], sym["inv"]
    ns = sym["nslots"]
    F = np.zeros(ns * (ns + 1) // 2)
    yv = np.zeros(ns)
    Lcols, rinv, y = [], np.zeros(n), np.zeros(n)
    Ad = A.tocsr()
    for j in range(n):
        old = order[j]
        a, f = sym["cols"][j]
        p = sym["piv_slot"][j]
        slots = sym["col_slots"][j]
        d = F[tri(p, p)] + Ad[old, old]
        F[tri(p, p)] = 0.0
        col = np.empty(len(slots))
        for q, s in enumerate(slots):
            col[q] = F[tri(s, p)]
            F[tri(s, p)] = 0.0
            if q < len(a):
                col[q] += Ad[order[a[q]], old]
        assert d > 0
        ri = 1.0 / np.sqrt(d)
        l = col * ri
        yp = (yv[p] + b[old]) * ri
        yv[p] = 0.0
        for qa, sa in enumerate(slots):
            yv[sa] -= l[qa] * yp
            for qb in range(qa + 1):
                F[tri(sa, slots[qb])] -= l[qa] * l[qb]
        Lcols.append(l)
        rinv[j] = ri
        y[j] = yp
    assert np.all(F == 0) or np.abs(F).max() < 1e-300, "front not empty at the end"
    wv = np.zeros(ns)
    w = np.zeros(n)
    for j in range(n - 1, -1, -1):
        slots = sym["col_slots"][j]
        acc = y[j]
        for q, s in enumerate(slots):
            acc -= Lcols[j][q] * wv[s]
        wj = acc * rinv[j]
        wv[sym["piv_slot"][j]] = wj
        w[order[j]] = wj
    return w


if __name__ == "__main__":
    sys.path.insert(0, ".")
    from bayesianinferencedl_b200 import get_space
    from bayesianinferencedl_b200.assembly import build_operators
    import scipy.sparse.linalg as spla

    ms = [int(a) for a in sys.argv[1:]] or [1, 3]
    for m in ms:
        if m > 0:
            V = get_space(40, m=m)
        else:
            from tests.meshes import unstructured_fin
            from bayesianinferencedl_b200.fom.thermal_fin import FinSpace
            V = FinSpace.from_mesh(*unstructured_fin(h=0.125 if m == 0 else 1.0 / (-m)))
        ops = build_operators(V)
        roots = np.flatnonzero(ops.rhs != 0)
        for ck in ("size", "-size"):
            sym = analyse(ops.n, ops.row_ptr, ops.col_idx, roots, ck)
            print(m, ck, sym["stats"], flush=True)
        if ops.n < 3000:
            theta = np.random.default_rng(0).uniform(0.1, 3.5, 9)
            A = ops.csr(ops.affine_values(theta))
            w = run_program(sym, A, ops.rhs)
            ref = spla.splu(A.tocsc()).solve(ops.rhs)
            print("  max rel err vs splu:", np.abs(w - ref).max() / np.abs(ref).max())


def entry_alloc(sym):
    """Entry-level allocation: every structural entry (i,k) of the active submatrix gets an address for its lifetime
    (first update .. gather at pivot k).  Returns max live entries and per-pivot address lists."""
    import heapq
    n = len(sym["cols"])
    structs = [a + f for a, f in sym["cols"]]
    structs = [sorted(s) for s in structs]
    addr = [dict() for _ in range(n)]      # addr[k][i] for i in struct(k) or i == k
    free, top, live, peak = [], 0, 0, 0
    n_fresh = 0
    for j in range(n):
        # gather column j: frees (j,j) and (i,j)
        for i in [j] + structs[j]:
            if i in addr[j]:
                heapq.heappush(free, addr[j][i])
                live -= 1
        s = structs[j]
        for qa, ia in enumerate(s):
            for ib in s[:qa + 1]:
                d = addr[ib]
                if ia not in d:
                    if free:
                        d[ia] = heapq.heappop(free)
                    else:
                        d[ia] = top
                        top += 1
                    live += 1
                    n_fresh += 1
                    peak = max(peak, live)
    return dict(max_addr=top, peak_live=peak, fresh=n_fresh)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[-1] == "entries":
    pass
