"""One matrix through gsi_debug_eigh (ncu target): python scripts/prof_one.py n team"""
import sys
import numpy as np
sys.path.insert(0, ".")
from collaborative_filtering_b200.api import Context
n, team = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(1)
W = np.triu((rng.random((n, n)) < 0.9) * (0.5 + 0.5 * rng.random((n, n))), 1); W = W + W.T
deg = W.sum(1); s = np.sqrt(1.0 / deg)
A = (np.diag(deg) - W) * s[:, None] * s[None, :]
ctx = Context(0)
r = ctx.debug_eigh(A, thr=1.02, team=team)
print("k", r["k"], "lam0", r["lam"][:3])
ctx.close()
