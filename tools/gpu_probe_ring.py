"""Small, time-bounded probe of the ring-mode streaming kernel (m = 8 mesh, slot reuse exercised)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bayesianinferencedl_b200 import get_space, _cabi
from bayesianinferencedl_b200.assembly import build_operators
m = int(os.environ.get("PROBE_M", 8))
V = get_space(40, m=m); ops = build_operators(V)
h = _cabi.TfinHandle(0)
h.set_operator(ops.row_ptr, ops.col_idx, ops.vals, ops.rhs); h.set_observation(*ops.obs_csr())
print("n", ops.n, "bandwidth", h.get_int("stream_bandwidth"), "ld", h.get_int("stream_ld"), flush=True)
theta = np.random.default_rng(2).uniform(0.1, 10.0, (12, 9))
ref = None
cfgs = [tuple(int(x) for x in c.split(",")) for c in os.environ.get("PROBE_CFG", "8,0;4,1;8,1").split(";")]
for tile, ring in cfgs:
    h.set_int("stream_tile", tile); h.set_int("stream_ring", ring)
    t0 = time.time(); out = h.fom_affine(theta, maxit=int(os.environ.get("PROBE_MAXIT", 20000)))
    print("tile", tile, "ring", h.get_int("stream_ring"), "iters", out["iters"][:4], "status", out["status"][:4], "qoi0", out["qoi"][0, :3], "%.2fs" % (time.time() - t0), flush=True)
    if ref is None: ref = out["qoi"]
    else: print("   max rel diff vs direct", np.abs(out["qoi"] - ref).max() / np.abs(ref).max(), flush=True)
