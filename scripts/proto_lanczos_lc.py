"""numpy prototype of the Lanczos fast path of gsi_local_calc_host (kern_lc.cuh: lc_lanczos_kernel): convergence of the
smallest Ritz value of G = P[unrated, unrated], P = L L^T, from the constant start vector, on the ML-100K-shaped fold the GPU
probe uses (scripts/probe_local_calc.py; knn2-style cosine weights computed densely here).  Prints, per sampled (movie, user)
pair, the three smallest eigenvalues of G, the number of Lanczos steps until the residual estimate drops below 1e-13 (checked
every 8 steps) and the error against numpy.linalg.eigvalsh.  Result recorded in profiles/r01w_local_calc.md: 8-24 steps,
error <= 1.6e-15.  CPU only; imports the oracle for the Laplacian restatement (a script, not product code).

    python scripts/proto_lanczos_lc.py [shape] [n_movies]
"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from collaborative_filtering_b200 import datasets as D
from oracle import gsi_oracle as O
shape = sys.argv[1] if len(sys.argv) > 1 else "ml-100k"
n_sample = int(sys.argv[2]) if len(sys.argv) > 2 else 12
r = D.make_ratings(shape); folds = D.fold_split(r, 5)
val_idx = np.sort(folds[0]); trn_idx = np.sort(np.concatenate(folds[1:]))
_, v_off, v_items, v_rat = D.subset(r, val_idx)
_, t_off, t_items, t_rat = D.subset(r, trn_idx)
I = r.n_items + 1
R = np.zeros((len(trn_idx), I), dtype=np.float32)
for u in range(len(trn_idx)):
    R[u, t_items[t_off[u]:t_off[u+1]]] = t_rat[t_off[u]:t_off[u+1]]
B = (R > 0).astype(np.float32)
num = R.T @ R; den1 = (R*R).T @ B; den2 = den1.T; cnt = B.T @ B
with np.errstate(all="ignore"):
    W = np.where(cnt > 5, num / (np.sqrt(den1) * np.sqrt(den2)), 0).astype(np.float32)
np.fill_diagonal(W, 0)
W = np.where(W > 0.01, W, 0)
Wd = W.astype(np.float64)
keep = Wd > 0.1
Wt = np.where(keep, Wd, 0.0)
print("edges", keep.sum())
# test ratings by movie
test = {}
for ui in range(len(val_idx)):
    for t in range(v_off[ui], v_off[ui+1]):
        test.setdefault(int(v_items[t]), {})[ui] = float(v_rat[t])
rng = np.random.default_rng(0)
movies = [m for m in test if keep[m].sum() + 1 >= 3]
sample = rng.choice(movies, size=n_sample, replace=False)
def lanczos_min(G, tol=1e-13, maxit=600, seed=1):
    n = G.shape[0]
    rs = np.random.default_rng(seed)
    q = np.ones(n) / np.sqrt(n)
    qp = np.zeros(n); beta = 0.0
    al, be = [], []
    theta_prev = None
    for k in range(maxit):
        w = G @ q - beta * qp
        a = q @ w; w -= a * q
        al.append(a)
        beta = np.linalg.norm(w)
        if (k % 8 == 7) or beta < 1e-300:
            T = np.diag(al) + np.diag(be, 1) + np.diag(be, -1)
            ev, S = np.linalg.eigh(T)
            theta = ev[0]; resid = abs(beta * S[-1, 0])
            if resid < tol or beta < 1e-300:
                return theta, k + 1, resid
        be.append(beta)
        qp, q = q, w / beta
    return theta, maxit, resid
its = []
for m in sample:
    nodes = [m] + [int(j) for j in np.nonzero(keep[m])[0] if j != m]
    n = len(nodes)
    ww = Wt[np.ix_(nodes, nodes)].copy(); np.fill_diagonal(ww, 0)
    ww[0, :] = Wt[m, nodes]; ww[:, 0] = Wt[m, nodes]; ww[0, 0] = 0
    _, _, L = O.normalized_laplacian(ww)
    P = L @ L.T
    pos = {v: i for i, v in enumerate(nodes)}
    for ui in list(test[m])[:4]:
        rated = np.zeros(n, bool)
        for t in range(v_off[ui], v_off[ui+1]):
            j = pos.get(int(v_items[t]), -1)
            if j > 0: rated[j] = True
        if rated.sum() == 0: continue
        G = P[np.ix_(~rated, ~rated)]
        ev = np.linalg.eigvalsh(G)
        th, k, res = lanczos_min(G)
        its.append(k)
        print("n=%4d nunr=%4d kk=%3d lam1=%.3e lam2=%.3e lam3=%.3e | lanczos its=%3d err=%.2e resid=%.1e" % (n, (~rated).sum(), rated.sum(), ev[0], ev[1], ev[2], k, abs(th-ev[0]), res))
print("iterations: mean %.0f max %d" % (np.mean(its), max(its)))
