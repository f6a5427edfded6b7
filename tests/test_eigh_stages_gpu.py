"""Stage-wise GPU parity of the large-n eigensolver (Householder tridiagonalisation -> divide & conquer
-> back-transform) through the C ABI test hook gsi_debug_eigh, against the oracle's eigensolve
(oracle.gsi_oracle.eig_lower == LAPACK syevd on the lower triangle, the restatement of Eigen's
SelfAdjointEigenSolver, precompute_local.cpp:231).

Tolerances (fp64): the tridiagonal T must be orthogonally similar to A -> its spectrum matches to 1e-12;
eigenvalues 1e-12; residual and orthonormality of the kept vectors 1e-11; the kept count is exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from collaborative_filtering_b200.api import Context
    c = Context(0)
    yield c
    c.close()


def _laplacian(n, density, seed):
    from oracle import gsi_oracle as O
    rng = np.random.default_rng(seed)
    w = np.triu((rng.random((n, n)) < density) * (0.5 + 0.5 * rng.random((n, n))), 1)
    w = w + w.T
    _, _, ll2 = O.normalized_laplacian(w)          # P4+P5, precompute_local.cpp:196-222
    return np.tril(ll2) + np.tril(ll2, -1).T


def _householder_q(v, tau):
    n = v.shape[0]
    q = np.eye(n)
    for j in range(n - 2, -1, -1):          # Q = H_0 H_1 ... H_{n-2}
        vj = v[:, j].copy()
        vj[:j + 1] = 0.0
        vj[j + 1] = 1.0
        q -= tau[j] * np.outer(vj, vj @ q)
    return q


@pytest.mark.parametrize("n,density,team", [(33, 0.9, 0), (97, 0.3, 0), (128, 0.9, 0), (160, 0.9, 0),      # team 0: the default routing (n <= 160: CTA-resident tridiagonalisation)
                                            (33, 0.9, 1), (64, 0.5, 1), (65, 0.9, 2), (161, 0.9, 1), (200, 0.1, 4),
                                            (257, 0.9, 3), (500, 0.9, 1), (777, 0.5, 8), (1000, 0.9, 16),
                                            (1100, 0.9, 2), (2100, 0.9, 148)])
def test_stages_against_oracle(ctx, n, density, team):
    from oracle import gsi_oracle as O
    a = _laplacian(n, density, 31413 + n)
    lam_ref, _ = O.eig_lower(a)
    thr = float(np.float32(np.median(lam_ref) + 0.013))
    r = ctx.debug_eigh(a, thr=thr, team=team)
    # stage 1: T similar to A, reflectors reproduce it
    t = np.diag(r["d"]) + np.diag(r["e"], 1) + np.diag(r["e"], -1)
    assert np.abs(np.linalg.eigvalsh(t) - lam_ref).max() < 1e-12
    if n <= 300:
        q = _householder_q(r["v"], r["tau"])
        assert np.abs(q.T @ a @ q - t).max() < 1e-12
        assert np.abs(q.T @ q - np.eye(n)).max() < 1e-12
    # stage 2: eigenvalues ascending, equal to the oracle's
    assert np.all(np.diff(r["lam"]) >= 0)
    assert np.abs(r["lam"] - lam_ref).max() < 1e-12
    # cutoff and stage 3
    k_ref = max(2, int((lam_ref <= np.float32(thr)).sum()))
    assert r["k"] == k_ref
    u = r["u"]
    assert np.abs(a @ u - u * r["lam"][:k_ref]).max() < 1e-11
    assert np.abs(u.T @ u - np.eye(k_ref)).max() < 1e-11


def test_isolated_items_all_deflate(ctx):
    """No edges at all: L = I, every pole deflates in every merge (lambda = 1, n times)."""
    n = 300
    r = ctx.debug_eigh(np.eye(n), thr=2.0, team=1)
    assert np.array_equal(r["lam"], np.ones(n)) and r["k"] == n
    assert np.abs(r["u"].T @ r["u"] - np.eye(n)).max() < 1e-14


def test_complete_graph_closed_form(ctx):
    """K_n with unit weights: lambda = {0, n/(n-1) x (n-1)} (SURVEY.md 8c closed form) -- a fully
    degenerate cluster, the Givens deflation path of the merge."""
    n = 130
    a = (np.eye(n) * n - np.ones((n, n))) / (n - 1.0)
    r = ctx.debug_eigh(a, thr=0.5, team=2)
    ref = np.r_[0.0, np.full(n - 1, n / (n - 1.0))]
    assert np.abs(r["lam"] - ref).max() < 1e-12
    assert r["k"] == 2
    u = r["u"]
    assert np.abs(np.abs(u[:, 0]) - 1 / np.sqrt(n)).max() < 1e-12
    assert np.abs(a @ u - u * r["lam"][:2]).max() < 1e-12 and np.abs(u.T @ u - np.eye(2)).max() < 1e-12


def test_team_size_does_not_change_results(ctx):
    """Fixed reduction orders: the same matrix gives bit-identical T for repeated runs of one team size."""
    a = _laplacian(400, 0.9, 7)
    r1 = ctx.debug_eigh(a, team=4)
    r2 = ctx.debug_eigh(a, team=4)
    assert np.array_equal(r1["d"], r2["d"]) and np.array_equal(r1["e"], r2["e"]) and np.array_equal(r1["lam"], r2["lam"])
    assert np.array_equal(r1["u"], r2["u"])
