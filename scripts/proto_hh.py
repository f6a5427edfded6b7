"""numpy prototype of the large-n eigensolver (kern_trd / kern_dc / kern_bt), written phase by phase
the way the CUDA kernels are organised, so that the math (panel-deferred Householder updates, the
deflation scan, the safeguarded secular solver, the Loewner z-hat, the WY back-transform) can be
checked on the CPU before it is spent GPU time on.  Development aid only: nothing imports it.

    python scripts/proto_hh.py            # random Laplacian-like matrices, compares with eigh
"""
import sys

import numpy as np

EPS = np.finfo(np.float64).eps / 2  # LAPACK dlamch('E') = 2^-53


# ----------------------------------------------------------------------------------------
# stage 1: blocked Householder tridiagonalisation (lower), panel-deferred rank-2k updates
# ----------------------------------------------------------------------------------------
def trd_blocked(A, nb=8):
    """A symmetric (full storage; only tiles on/below the diagonal are read like the kernel).
    Returns d, e, V (n x n, column j holds reflector j with v[j+1]=1, zeros above), tau."""
    A = A.copy()
    n = A.shape[0]
    d = np.zeros(n)
    e = np.zeros(max(n - 1, 0))
    tau = np.zeros(max(n - 1, 0))
    VV = np.zeros((n, n))
    j0 = 0
    while j0 < n - 1:
        pw = min(nb, n - 1 - j0)
        V = np.zeros((n, pw))
        W = np.zeros((n, pw))
        # acol for the first column of the panel: the matrix is up to date
        acol = A[:, j0].copy()
        for jj in range(pw):
            j = j0 + jj
            d[j] = acol[j]
            # --- barrier 1: norm of acol[j+1:]
            x = acol[j + 1:]
            alpha = x[0]
            xnorm2 = float(np.dot(x[1:], x[1:]))
            if xnorm2 == 0.0:
                tj = 0.0
                beta = alpha
                scale = 0.0
            else:
                beta = -np.copysign(np.sqrt(alpha * alpha + xnorm2), alpha)
                tj = (beta - alpha) / beta
                scale = 1.0 / (alpha - beta)
            v = np.zeros(n)
            v[j + 1] = 1.0
            v[j + 2:] = x[1:] * scale
            e[j] = beta
            tau[j] = tj
            V[:, jj] = v
            VV[:, j] = v
            # --- phase B: symv with the panel-start matrix + small dot products
            Av = np.tril(A, 0) @ v + np.tril(A, -1).T @ v      # lower tiles only, both directions
            Wtv = W[:, :jj].T @ v
            Vtv = V[:, :jj].T @ v
            vAv = float(v @ Av)
            # --- barrier 2; phase C
            p = Av - V[:, :jj] @ Wtv - W[:, :jj] @ Vtv
            pv = vAv - 2.0 * float(Vtv @ Wtv)
            alpha2 = -0.5 * tj * tj * pv
            w = tj * p + alpha2 * v
            w[:j + 1] = 0.0
            W[:, jj] = w
            # phase A of the next column
            if j + 1 < n:
                acol = A[:, j + 1] - V[:, :jj + 1] @ W[j + 1, :jj + 1] - W[:, :jj + 1] @ V[j + 1, :jj + 1]
        # trailing update (lower tiles incl. full diagonal tiles)
        jn = j0 + pw
        A[jn:, jn:] -= V[jn:, :] @ W[jn:, :].T + W[jn:, :] @ V[jn:, :].T
        j0 = jn
        if j0 == n - 1:
            d[n - 1] = A[n - 1, n - 1]
    if n == 1:
        d[0] = A[0, 0]
    return d, e, VV, tau


# ----------------------------------------------------------------------------------------
# stage 2: divide and conquer on the tridiagonal
# ----------------------------------------------------------------------------------------
def leaf_bounds(n, leaf=32):
    """Split [0,n) into 2^L nearly equal leaves of size <= leaf; returns list of levels, level 0 =
    leaves [(off, size)], level l = merges [(off, n1, n2)]."""
    L = 0
    while (n + (1 << L) - 1) >> L > leaf:
        L += 1
    nl = 1 << L
    cuts = [(n * i) // nl for i in range(nl + 1)]
    levels = [[(cuts[i], cuts[i + 1] - cuts[i]) for i in range(nl)]]
    step = 1
    for l in range(1, L + 1):
        merges = []
        for i in range(0, nl, 2 * step):
            merges.append((cuts[i], cuts[i + step] - cuts[i], cuts[i + 2 * step] - cuts[i + step]))
        levels.append(merges)
        step *= 2
    return levels


def secular_roots(d, z2, rho, maxit=80):
    """All k roots of 1/rho + sum z2_i / (d_i - lam) = 0, d ascending & distinct, z2 > 0, rho > 0.
    Vectorised over the roots (one GPU thread per root).  Returns (orig index, tau): lam_j =
    d[orig_j] + tau_j, so that d_i - lam_j is formed as (d_i - d[orig_j]) - tau_j."""
    k = len(d)
    rhoinv = 1.0 / rho
    j = np.arange(k)
    last = j == k - 1
    dj = d
    dn = np.where(last, d[-1] + rho * z2.sum(), np.append(d[1:], 0.0))   # upper end of the interval
    gap = dn - dj
    idx = np.arange(k)
    left = idx[None, :] <= j[:, None]       # poles i <= j  (psi part)

    def evaluate(orig_d, tau):
        delta = (d[None, :] - orig_d[:, None]) - tau[:, None]            # d_i - lam_j
        t = z2[None, :] / delta
        psi = np.where(left, t, 0.0).sum(1)
        phi = np.where(left, 0.0, t).sum(1)
        t2 = t / delta
        dpsi = np.where(left, t2, 0.0).sum(1)
        dphi = np.where(left, 0.0, t2).sum(1)
        return psi, phi, dpsi, dphi

    # choose the origin: evaluate at the midpoint with origin d_j
    mid = 0.5 * gap
    psi, phi, _, _ = evaluate(dj, mid)
    wmid = rhoinv + psi + phi
    use_left = (wmid >= 0.0) | last                       # root in the left half -> origin d_j
    orig = np.where(use_left, j, np.minimum(j + 1, k - 1))
    orig_d = d[orig]
    # bracket in tau (relative to the origin)
    lo = np.where(use_left, 0.0, -mid)
    hi = np.where(use_left, mid, 0.0)
    hi = np.where(last, np.where(wmid >= 0.0, mid, gap), hi)
    lo = np.where(last & (wmid < 0.0), mid, lo)
    tau = np.where(use_left, np.where(last & (wmid < 0.0), 0.75 * gap, 0.5 * mid), -0.5 * mid)
    # the two poles used by the interpolation, relative to the origin
    d1 = dj - orig_d                                       # <= 0
    d2 = np.where(last, np.inf, dn - orig_d)
    done = np.zeros(k, bool)
    for it in range(maxit):
        psi, phi, dpsi, dphi = evaluate(orig_d, tau)
        w = rhoinv + psi + phi
        erretm = 8.0 * (np.abs(psi) + np.abs(phi)) + rhoinv + np.abs(tau) * (dpsi + dphi)
        conv = np.abs(w) <= EPS * erretm
        done |= conv
        if done.all():
            break
        lo = np.where(~done & (w < 0), np.maximum(lo, tau), lo)
        hi = np.where(~done & (w >= 0), np.minimum(hi, tau), hi)
        D1 = d1 - tau
        D2 = d2 - tau
        with np.errstate(all="ignore"):
            # interior: c eta^2 - a eta + b = 0
            c = w - D1 * dpsi - D2 * dphi
            a = (D1 + D2) * w - D1 * D2 * (dpsi + dphi)
            b = D1 * D2 * w
            disc = np.sqrt(np.abs(a * a - 4.0 * b * c))
            eta_i = np.where(a <= 0, (a - disc) / (2.0 * c), 2.0 * b / (a + disc))
            eta_i = np.where(c == 0, b / a, eta_i)
            # last root: one pole (psi ~ s + p/(d1 - x)), phi = 0
            c1 = rhoinv + psi - dpsi * D1
            eta_l = D1 + dpsi * D1 * D1 / c1
            eta = np.where(last, eta_l, eta_i)
            newton = -w / (dpsi + dphi)
            eta = np.where(~np.isfinite(eta) | (w * eta >= 0), newton, eta)
            cand = tau + eta
            bad = ~np.isfinite(cand) | (cand <= lo) | (cand >= hi)
            cand = np.where(bad, 0.5 * (lo + hi), cand)
        stuck = (hi - lo) <= 4.0 * EPS * np.maximum(np.abs(lo), np.abs(hi))
        done |= stuck
        tau = np.where(done, tau, cand)
    return orig, tau, it + 1


def merge(d1v, Q1, d2v, Q2, beta):
    """One D&C merge.  Q1 (n1 x n1), Q2 (n2 x n2): eigenvectors of the two halves (whose coupling
    diagonal entries were reduced by |beta|).  Returns (lam ascending, Q (m x m))."""
    n1, n2 = len(d1v), len(d2v)
    m = n1 + n2
    rho = 2.0 * abs(beta)
    sgn = 1.0 if beta >= 0 else -1.0
    z = np.concatenate([Q1[-1, :], sgn * Q2[0, :]]) / np.sqrt(2.0)
    d = np.concatenate([d1v, d2v])
    Q = np.zeros((m, m))
    Q[:n1, :n1] = Q1
    Q[n1:, n1:] = Q2
    perm = np.argsort(d, kind="stable")
    d = d[perm].copy()
    z = z[perm].copy()
    Q = Q[:, perm].copy()
    tol = 8.0 * EPS * max(np.abs(d).max(), np.abs(z).max())
    defl = np.zeros(m, bool)
    nrot = 0
    if rho * np.abs(z).max() <= tol:
        defl[:] = True
    else:
        pj = -1
        for i in range(m):
            if rho * abs(z[i]) <= tol:
                defl[i] = True
                continue
            if pj >= 0:
                s = z[pj]
                c = z[i]
                tt = np.hypot(c, s)
                t = d[i] - d[pj]
                c /= tt
                s = -s / tt
                if abs(t * c * s) <= tol:
                    z[i] = tt
                    z[pj] = 0.0
                    x = Q[:, pj].copy()
                    y = Q[:, i].copy()
                    Q[:, pj] = c * x + s * y
                    Q[:, i] = c * y - s * x
                    t2 = d[pj] * c * c + d[i] * s * s
                    d[i] = d[pj] * s * s + d[i] * c * c
                    d[pj] = t2
                    defl[pj] = True
                    nrot += 1
            pj = i
    nd = np.flatnonzero(~defl)
    k = len(nd)
    lam_all = d.copy()
    Qn = Q.copy()
    stats = dict(k=k, m=m, nrot=nrot, iters=0)
    if k > 0:
        dk = d[nd]
        zk = z[nd]
        if k == 1:
            lam_k = dk + rho * zk * zk
            U = np.ones((1, 1))
        else:
            orig, tau, iters = secular_roots(dk, zk * zk, rho)
            stats["iters"] = iters
            # delta[i, j] = d_i - lam_j
            delta = (dk[:, None] - dk[orig][None, :]) - tau[None, :]
            lam_k = dk[orig] + tau
            # Loewner: zhat_i^2 = prod_j (lam_j - d_i) / prod_{j != i} (d_j - d_i)   (/rho, dropped: normalised)
            num = -delta                                     # lam_j - d_i
            den = dk[None, :] - dk[:, None]                  # d_j - d_i
            np.fill_diagonal(den, 1.0)
            ratio = num / den
            zhat2 = np.prod(ratio, axis=1)
            zhat = np.sqrt(np.abs(zhat2)) * np.sign(zk)
            U = zhat[:, None] / delta
            U /= np.linalg.norm(U, axis=0)[None, :]
        lam_all[nd] = lam_k
        Qn[:, nd] = Q[:, nd] @ U
    order = np.argsort(lam_all, kind="stable")
    return lam_all[order], Qn[:, order], stats


def dc_solve(d, e, leaf=32, verbose=False):
    n = len(d)
    d = d.copy()
    levels = leaf_bounds(n, leaf)
    # tear: subtract |beta| at every cut of every level
    for lev in levels[1:]:
        for (off, n1, n2) in lev:
            b = abs(e[off + n1 - 1])
            d[off + n1 - 1] -= b
            d[off + n1] -= b
    lam = np.zeros(n)
    Q = np.zeros((n, n))
    for (off, sz) in levels[0]:
        T = np.diag(d[off:off + sz]) + np.diag(e[off:off + sz - 1], 1) + np.diag(e[off:off + sz - 1], -1)
        w, v = np.linalg.eigh(T)
        lam[off:off + sz] = w
        Q[off:off + sz, off:off + sz] = v
    for li, lev in enumerate(levels[1:]):
        for (off, n1, n2) in lev:
            m = n1 + n2
            l2, Q2, st = merge(lam[off:off + n1], Q[off:off + n1, off:off + n1], lam[off + n1:off + m],
                               Q[off + n1:off + m, off + n1:off + m], e[off + n1 - 1])
            lam[off:off + m] = l2
            Q[off:off + m, off:off + m] = Q2
            if verbose:
                print("  level %d merge off=%d m=%d k=%d rot=%d iters=%d" % (li + 1, off, m, st["k"], st["nrot"], st["iters"]))
    return lam, Q


# ----------------------------------------------------------------------------------------
# stage 3: back-transform  U = H_0 H_1 ... H_{n-2} Z   in WY panels
# ----------------------------------------------------------------------------------------
def form_T(V, tau):
    """Forward columnwise T of the compact WY form: H_0 ... H_{b-1} = I - V T V^T."""
    b = V.shape[1]
    T = np.zeros((b, b))
    G = V.T @ V
    for j in range(b):
        T[j, j] = tau[j]
        if j:
            T[:j, j] = -tau[j] * (T[:j, :j] @ G[:j, j])
    return T


def back_transform(VV, tau, Z, nbt=8):
    n = VV.shape[0]
    Z = Z.copy()
    nref = n - 1
    starts = list(range(0, nref, nbt))
    for j0 in reversed(starts):
        b = min(nbt, nref - j0)
        V = VV[:, j0:j0 + b]
        T = form_T(V, tau[j0:j0 + b])
        X = V.T @ Z
        Z -= V @ (T @ X)
    return Z


def eig_hh(A, nb=8, leaf=32, nbt=8, verbose=False):
    d, e, VV, tau = trd_blocked(A, nb)
    lam, Q = dc_solve(d, e, leaf, verbose)
    U = back_transform(VV, tau, Q, nbt)
    return lam, U, (d, e)


def laplacian_like(n, density, rng):
    W = np.triu((rng.random((n, n)) < density) * (0.5 + 0.5 * rng.random((n, n))), 1)
    W = W + W.T
    deg = W.sum(1)
    deg[deg == 0] = 1.0
    s = np.sqrt(1.0 / deg)
    return (np.diag(deg) - W) * s[:, None] * s[None, :]


def main():
    rng = np.random.default_rng(31413)
    worst = 0.0
    for n, dens in [(1, 0.9), (2, 0.9), (3, 0.9), (33, 0.9), (64, 0.9), (65, 0.1), (200, 0.9), (257, 0.05), (300, 0.0), (500, 0.9), (777, 0.5)]:
        A = laplacian_like(n, dens, rng)
        if dens == 0.0:
            A = np.eye(n)          # isolated items: lam = 1 (all deflated)
        lam, U, (d, e) = eig_hh(A, nb=8 if n < 100 else 32, leaf=8 if n < 100 else 32, nbt=16, verbose="-v" in sys.argv)
        ref = np.linalg.eigvalsh(A)
        T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        err_t = np.abs(np.linalg.eigvalsh(T) - ref).max()
        err_l = np.abs(lam - ref).max()
        res = np.abs(A @ U - U * lam[None, :]).max()
        orth = np.abs(U.T @ U - np.eye(n)).max()
        print("n=%4d dens=%.2f  |T-spec|=%.1e |dlam|=%.1e resid=%.1e orth=%.1e" % (n, dens, err_t, err_l, res, orth))
        worst = max(worst, err_l, res, orth)
    # clustered / glued Wilkinson-like tridiagonals straight into D&C
    for name, (d, e) in {
        "wilkinson21x5": (np.tile(np.abs(np.arange(-10, 11)).astype(float), 5), np.concatenate([np.r_[np.ones(20), 1e-8]] * 5)[:-1]),
        "const": (np.ones(300), np.full(299, 0.5)),
        "tiny-e": (np.linspace(0, 1, 200), np.full(199, 1e-12)),
        "random": (rng.standard_normal(400), rng.standard_normal(399)),
        "negative-e": (rng.standard_normal(130), -np.abs(rng.standard_normal(129))),
    }.items():
        n = len(d)
        lam, Q = dc_solve(d, e, 16, verbose="-v" in sys.argv)
        T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        ref = np.linalg.eigvalsh(T)
        err_l = np.abs(lam - ref).max()
        res = np.abs(T @ Q - Q * lam[None, :]).max()
        orth = np.abs(Q.T @ Q - np.eye(n)).max()
        print("%-14s n=%4d |dlam|=%.1e resid=%.1e orth=%.1e" % (name, n, err_l, res, orth))
        worst = max(worst, err_l / max(1, np.abs(ref).max()), res / max(1, np.abs(ref).max()), orth)
    print("worst", worst)
    assert worst < 1e-11


if __name__ == "__main__":
    main()
