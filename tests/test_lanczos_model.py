"""A numpy model of the decision logic of lc_lanczos_kernel (collaborative_filtering_b200/csrc/kern_lc.cuh): same start
vector, same masked recurrence without reorthogonalisation, Ritz value checked every 4 steps from step 8, done when it
moves by <= 1e-14 max(1, theta) or the Krylov space is exhausted, given up (exact path) after 96 steps.  The model is run
over a zoo of local graphs -- random sparse and dense, stars (lambda = 1 with high multiplicity), cliques, paths, rings,
two communities -- and every pair it declares converged must agree with numpy.linalg.eigvalsh to 1e-9 in w_lim: this
is the property the GPU fast path relies on when it skips the exact tridiagonalisation.  CPU only."""
import numpy as np
import pytest

from oracle import gsi_oracle as O

LZ_MAX = 96


def lanczos_model(P, rated):
    n = P.shape[0]
    mask = ~rated
    n_unr = int(mask.sum())
    q = np.where(mask, 1.0 / np.sqrt(n_unr), 0.0)
    qp = np.zeros(n)
    alpha, beta = [], []
    beta_prev, theta, theta_old = 0.0, 0.0, 1e300
    for k in range(min(LZ_MAX, n_unr)):
        w = np.where(mask, P @ q, 0.0) - beta_prev * qp
        a = float(q @ w)
        w = w - a * q
        b = float(np.sqrt(w @ w))
        alpha.append(a)
        beta.append(b)
        exhausted = (k == n_unr - 1) or not (b > 1e-14 * (abs(a) + beta_prev))
        if exhausted or (k + 1 >= 8 and (k + 1) % 4 == 0):
            T = np.diag(alpha) + np.diag(beta[:-1], 1) + np.diag(beta[:-1], -1)
            theta = float(np.linalg.eigvalsh(T)[0])
            if exhausted or abs(theta_old - theta) <= 1e-14 * max(1.0, abs(theta)):
                return theta, k + 1, True
            theta_old = theta
        qp, q, beta_prev = q, w / b, b
    return theta, min(LZ_MAX, n_unr), False


def _graphs():
    rng = np.random.default_rng(11)
    out = []
    for n, dens in ((12, 0.3), (40, 0.15), (40, 0.6), (90, 0.05), (90, 0.9), (200, 0.3)):
        a = np.triu((rng.random((n, n)) < dens) * rng.uniform(0.11, 1.0, (n, n)), 1)
        out.append(("random n=%d d=%.2f" % (n, dens), a + a.T))
    star = np.zeros((30, 30))
    out.append(("star", star))
    out.append(("clique", np.full((25, 25), 0.8) - 0.8 * np.eye(25)))
    path = np.zeros((50, 50))
    for i in range(1, 49):
        path[i, i + 1] = path[i + 1, i] = 0.5
    out.append(("path", path))
    ring = path.copy()
    ring[1, 49] = ring[49, 1] = 0.5
    out.append(("ring", ring))
    two = np.zeros((60, 60))
    two[1:30, 1:30] = 0.9
    two[30:, 30:] = 0.9
    two[29, 30] = two[30, 29] = 0.2
    np.fill_diagonal(two, 0.0)
    out.append(("two communities", two))
    res = []
    for name, ww in out:
        ww = ww.copy()
        ww[0, 1:] = ww[1:, 0] = rng.uniform(0.11, 1.0, ww.shape[0] - 1)      # node 0 = the movie: linked to every neighbour
        ww[0, 0] = 0.0
        res.append((name, ww))
    return res


@pytest.mark.parametrize("name,ww", _graphs(), ids=[g[0] for g in _graphs()])
def test_converged_means_correct(name, ww):
    rng = np.random.default_rng(5)
    n = ww.shape[0]
    _, _, ll2 = O.normalized_laplacian(ww)
    P = ll2 @ ll2.T
    given_up = 0
    steps = []
    for frac in (0.05, 0.2, 0.5, 0.8, 0.95):
        for _ in range(6):
            rated = rng.random(n) < frac
            rated[0] = False
            if not rated.any():
                rated[1 + rng.integers(n - 1)] = True
            exact = float(np.linalg.eigvalsh(P[np.ix_(~rated, ~rated)])[0])
            theta, k, ok = lanczos_model(P, rated)
            steps.append(k)
            if not ok:
                given_up += 1
                continue
            assert abs(np.sqrt(max(theta, 0.0)) - np.sqrt(max(exact, 0.0))) <= 1e-9, (name, frac, k, theta, exact)
    assert given_up <= 6, "the fast path gives up on %d of 30 pairs of '%s'" % (given_up, name)
