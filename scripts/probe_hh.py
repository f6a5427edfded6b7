"""GPU probe of the Householder / D&C / back-transform stages through gsi_debug_eigh."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from collaborative_filtering_b200.api import Context


def laplacian_like(n, density, rng):
    W = np.triu((rng.random((n, n)) < density) * (0.5 + 0.5 * rng.random((n, n))), 1)
    W = W + W.T
    deg = W.sum(1); deg[deg == 0] = 1.0
    s = np.sqrt(1.0 / deg)
    return (np.diag(deg) - W) * s[:, None] * s[None, :]


def check(ctx, n, dens, team, thr=1e30):
    rng = np.random.default_rng(n * 7 + team)
    A = laplacian_like(n, dens, rng)
    t0 = time.time()
    r = ctx.debug_eigh(A, thr=thr, team=team)
    dt = time.time() - t0
    ref = np.linalg.eigvalsh(A)
    T = np.diag(r["d"]) + np.diag(r["e"], 1) + np.diag(r["e"], -1)
    e_t = np.abs(np.linalg.eigvalsh(T) - ref).max()
    e_l = np.abs(np.sort(r["lam"]) - ref).max()
    srt = bool(np.all(np.diff(r["lam"]) >= 0))
    k = r["k"]; U = r["u"]; lam = r["lam"][:k]
    res = np.abs(A @ U - U * lam[None, :]).max() if k else -1
    orth = np.abs(U.T @ U - np.eye(k)).max() if k else -1
    kref = max(2, int((ref <= np.float32(thr)).sum()))
    print("n=%5d dens=%.2f team=%3d  |T-spec|=%.1e |dlam|=%.1e sorted=%s k=%d (ref %d) resid=%.1e orth=%.1e  %.3fs"
          % (n, dens, team, e_t, e_l, srt, k, kref, res, orth, dt), flush=True)
    return max(e_t, e_l, res, orth)


def main():
    ctx = Context(0)
    worst = 0.0
    cases = [(33, 0.9, 1), (64, 0.9, 1), (65, 0.9, 2), (130, 0.5, 1), (200, 0.9, 1), (200, 0.9, 4), (257, 0.1, 3), (500, 0.9, 1),
             (777, 0.5, 8), (1000, 0.9, 1), (1000, 0.9, 16), (1500, 0.9, 2), (2100, 0.9, 148)]
    if len(sys.argv) > 1:
        cases = [tuple(float(x) if "." in x else int(x) for x in a.split(",")) for a in sys.argv[1:]]
    for n, dens, team in cases:
        worst = max(worst, check(ctx, n, dens, team))
        worst = max(worst, check(ctx, n, dens, team, thr=1.05))
    print("worst", worst)
    ctx.close()


if __name__ == "__main__":
    main()
