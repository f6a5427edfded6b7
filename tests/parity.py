"""Shared parity checks: GPU records vs the CPU oracle (tests only).

Tolerances (stated, north_star): integer structure (n, k, ids) exact; sig_min bit-exact (the GPU
reproduces the reference's float accumulation order); eigenvalues |d lam| <= 1e-10; eigenvectors:
residual / orthonormality <= 1e-9, sign convention exact, per-vector agreement for eigenvalues
isolated by a gap > 1e-5, and projector agreement on the kept subspace when the cut does not fall
inside a cluster."""
import numpy as np

from oracle import gsi_oracle as O

LAM_TOL = 1e-10
VEC_TOL = 1e-9


def check_user(items, weights, sig, k, lam, vec, tag=""):
    rec = O.precompute_user(0, items, weights, keep_ll2=True)
    n = len(items)
    assert np.array_equal(rec.sigs_min, sig), "%s sig_min not bit-exact (max diff %g)" % (tag, np.abs(rec.sigs_min - sig).max())
    lam_full, u_full = O.eig_lower(rec.ll2)
    k_or = len(rec.lam)
    if k != k_or:
        # legitimate only when an eigenvalue ties with the threshold
        _, thr = O.sig_min_rows(rec.ll2)
        near = np.abs(lam_full - float(thr)).min()
        assert near < 1e-9, "%s k mismatch: gpu %d oracle %d (nearest lam-thr %g)" % (tag, k, k_or, near)
    assert lam.shape == (k,) and vec.shape == (n, k)
    kk = min(k, n)
    assert np.abs(lam[:kk] - lam_full[:kk]).max() <= LAM_TOL, "%s lam diff %g" % (tag, np.abs(lam[:kk] - lam_full[:kk]).max())
    if k > n:                                               # B5 defined behaviour
        assert np.all(lam[n:] == 0) and np.all(vec[:, n:] == 0)
    a = np.tril(rec.ll2) + np.tril(rec.ll2, -1).T
    v = vec[:, :kk]
    assert np.abs(a @ v - v * lam[:kk]).max() <= VEC_TOL, "%s residual %g" % (tag, np.abs(a @ v - v * lam[:kk]).max())
    assert np.abs(v.T @ v - np.eye(kk)).max() <= VEC_TOL, "%s orthonormality" % tag
    # sign convention: entry of largest magnitude positive
    idx = np.argmax(np.abs(v), axis=0)
    assert np.all(v[idx, np.arange(kk)] > 0), "%s sign convention" % tag
    # isolated eigenvalues: same vector (same sign convention)
    gaps = np.full(n, np.inf)
    if n > 1:
        d = np.diff(lam_full)
        gaps[:-1] = np.minimum(gaps[:-1], d)
        gaps[1:] = np.minimum(gaps[1:], d)
    for j in range(kk):
        if gaps[j] > 1e-5:
            top = np.sort(np.abs(u_full[:, j]))[-2:] if n > 1 else np.array([0.0, 1.0])
            if abs(top[1] - top[0]) < 1e-6:
                continue                                    # the sign convention itself is a tie
            assert np.abs(v[:, j] - u_full[:, j]).max() <= 1e-8, \
                "%s vector %d differs by %g (gap %g)" % (tag, j, np.abs(v[:, j] - u_full[:, j]).max(), gaps[j])
    if kk < n and lam_full[kk] - lam_full[kk - 1] > 1e-5:
        p1 = v @ v.T
        p2 = u_full[:, :kk] @ u_full[:, :kk].T
        assert np.abs(p1 - p2).max() <= 1e-8, "%s kept-subspace projector %g" % (tag, np.abs(p1 - p2).max())


def check_records(recs, weights, users=None, tag=""):
    nu = len(recs.offsets) - 1
    for u in (range(nu) if users is None else users):
        items = recs.items[recs.offsets[u]: recs.offsets[u + 1]]
        check_user(items, weights, recs.sig_of(u), int(recs.k[u]), recs.lam_of(u), recs.vec_of(u), "%s user %d n=%d" % (tag, u, len(items)))
