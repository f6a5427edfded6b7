import sys, time, numpy as np
sys.path.insert(0, ".")
from collaborative_filtering_b200 import datasets as D
from collaborative_filtering_b200.api import Context
N_ITEMS = 17770
w = D.make_weights(N_ITEMS, density=0.9)
ctx = Context(0); ctx.set_workspace_limit(64 << 30); ctx.set_weights(w)
rng = np.random.default_rng(5)
for n in [int(x) for x in sys.argv[1:]]:
    items = np.sort(rng.choice(N_ITEMS, size=n, replace=False) + 1).astype(np.int32)
    off = np.array([0, n], dtype=np.int64)
    t0 = time.time(); recs = ctx.precompute(off, items); dt = time.time() - t0
    lam = recs.lam_of(0); U = recs.vec_of(0)
    # residual of a few kept eigenpairs against the Laplacian built on the host
    Wl = w[np.ix_(items, items)]; d = Wl.sum(1); d[d == 0] = 1; s = np.sqrt(1.0 / d)
    L = (np.diag(d) - Wl) * s[:, None] * s[None, :]
    Ls = np.tril(L) + np.tril(L, -1).T
    idx = [0, len(lam) // 2, len(lam) - 1]
    res = max(np.abs(Ls @ U[:, i] - lam[i] * U[:, i]).max() for i in idx)
    print("n=%d k=%d  %.2fs  residual %.2e  orth %.2e" % (n, recs.k[0], dt, res, abs(U[:, idx].T @ U[:, idx] - np.eye(3)).max()), flush=True)
ctx.close()
