"""Throughput of the Chebyshev graph filter (gsi_cheby_filter_host) on an item graph of MovieLens shape: dense-ish random graph
with nv vertices, kernel time of the supersteps (CUDA events) against the HBM roofline: 12 bytes per edge and superstep
(8-byte normalised weight + 4-byte column; the vertex vectors stay in L2).   python scripts/probe_cheby.py [nv] [density] [ncoef]"""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
from collaborative_filtering_b200.api import Context

nv = int(sys.argv[1]) if len(sys.argv) > 1 else 10681
density = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
ncoef = int(sys.argv[3]) if len(sys.argv) > 3 else 20
rng = np.random.default_rng(31413)
A = rng.random((nv, nv)) < density
A = np.triu(A, 1); A = A | A.T
col = np.nonzero(A)[1].astype(np.int32)
off = np.concatenate([[0], np.cumsum(A.sum(1))]).astype(np.int64)
U = np.triu(rng.uniform(0.11, 1.0, (nv, nv)), 1); U = U + U.T
w = U[A]
x = rng.normal(3.5, 1.0, nv)
coef = rng.normal(0, 1, ncoef)
ctx = Context(0)
ctx.cheby_filter(off, col, w, x, coef)                      # warm-up (allocations)
ctx.timing_enable(True); ctx.timing_reset()
t0 = time.time(); y = ctx.cheby_filter(off, col, w, x, coef); wall = time.time() - t0
tm = ctx.timing()["cheby"]
nnz = int(off[-1])
steps = ncoef - 1
bytes_alg = nnz * 12.0 * steps + nnz * (8 + 4 + 8 + 8)     # supersteps + degree (8) + normalise (4 + 8 + 8)
print(json.dumps({"nv": nv, "edges": nnz, "ncoef": ncoef, "kernel_ms": tm["ms"], "launches": tm["launches"], "wall_s": round(wall, 4),
                  "edge_updates_per_s": nnz * steps / (tm["ms"] * 1e-3), "algorithmic_GBps": bytes_alg / (tm["ms"] * 1e-3) / 1e9,
                  "frac_of_hbm_6549.8": bytes_alg / (tm["ms"] * 1e-3) / 1e9 / 6549.8, "finite": bool(np.isfinite(y).all())}))
ctx.close()
