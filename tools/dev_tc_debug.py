"""Developer probe for the tcgen05 matcher: small cases vs the oracle with error breakdown."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
from so_dso_place_recognition_b200 import api, synth

ctx = api.default_context(0)
rng = np.random.default_rng(0)
m, n = int(sys.argv[1]) if len(sys.argv) > 1 else 8, int(sys.argv[2]) if len(sys.argv) > 2 else 300
xyz, inten, off = synth.make_scan_set(max(m, n), 1024, planted_loops=True)
sig = O.sc_generate(xyz, inten, off, nthreads=8)
q, db = sig[:m], sig[:n]
rp, ri = O.sc_match_numpy(q, db)
ctx.set_match_algo(api.SODSO_ALGO_TC)
dp, di = api.processSC(q, db)
print("kernel", ctx.last_kernel_name, ctx.last_kernel_ms, "ms")
for name, a, b in (("d_p", dp, rp), ("d_i", di, ri)):
    e = np.abs(a - b)
    print(name, "max err", e.max(), "mean err", e.mean(), "nan", np.isnan(a).sum(), "gpu range", np.nanmin(a), np.nanmax(a),
          "ref range", b.min(), b.max())
    print("  per-query max err", np.round(e.max(axis=1)[:8], 7))
    print("  sample gpu", np.round(a[0, :6], 6), "ref", np.round(b[0, :6], 6))
# variant diagnosis: distance using only forward shifts / only shift 0
a = q[:, :1200] / np.linalg.norm(q[:, :1200], axis=1, keepdims=True)
b = db[:, :1200] / np.linalg.norm(db[:, :1200], axis=1, keepdims=True)
d0 = (1 - a @ b.T) / 2
print("if only shift 0:   max|gpu - d0|", np.abs(dp - d0).max())
