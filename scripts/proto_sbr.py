"""numpy prototype of the TWO-STAGE tridiagonalisation planned for the biggest users: what
`trd_kernel` would be replaced by for n > ~2048.  Development aid only: nothing imports it.

    stage 1  sy2sb   dense -> band of half-width b: per panel of b columns a QR of the block below the band (compact WY: V, T),
                     then the two-sided update  A22 <- H^T A22 H,  H = I - V T V^T,  as  W = A22 V T;  X = W - 1/2 V (T^T V^T W);
                     A22 -= X V^T + V X^T  -- one symm-like and one syr2k-like pass over the trailing matrix PER PANEL
                     (the one-stage kernel streams the trailing matrix once PER COLUMN), all of it level-3.
    stage 2  sb2st   band -> tridiagonal by bulge chasing: column by column, a length-b Householder reflector annihilates
                     the column below the sub-diagonal and the bulge it creates is chased down the band in blocks of b.
    back     Q = Q1 Q2 applied to the tridiagonal's eigenvectors: Q2 (the bulge-chasing reflectors, grouped per sweep) first,
             then Q1 (the panel WY factors), both as level-3 products over the kept columns.

    python scripts/proto_sbr.py [n] [b]      # checks against numpy.linalg.eigh and prints the work / traffic model

The model printed at the end uses the ML-10M shard of bench.py (sum n^3 = 1.23e12 for n > 44, 1.689 TB algorithmic bytes in the
one-stage kernel) to project what each stage costs on a B200 (6.55 TB/s, 36 TF/s FP64 tensor)."""
import sys

import numpy as np


def house(x):
    """LAPACK dlarfg: H x = beta e1, H = I - tau v v^T, v[0] = 1."""
    alpha = x[0]
    s = float(np.dot(x[1:], x[1:]))
    v = x.copy()
    v[0] = 1.0
    if s == 0.0:
        return v, 0.0, alpha
    beta = -np.copysign(np.sqrt(alpha * alpha + s), alpha)
    tau = (beta - alpha) / beta
    v[1:] = x[1:] / (alpha - beta)
    return v, tau, beta


def panel_qr(B):
    """Householder QR of a tall block, compact WY: Q = I - V T V^T (V unit lower trapezoidal), returns V, T, R."""
    m, b = B.shape
    R = B.copy()
    V = np.zeros((m, b))
    T = np.zeros((b, b))
    for j in range(min(b, m)):
        v, tau, beta = house(R[j:, j])
        V[j:, j] = v
        R[j:, j:] -= tau * np.outer(v, v @ R[j:, j:])
        R[j, j] = beta
        R[j + 1:, j] = 0.0
        T[j, j] = tau
        if j:
            T[:j, j] = -tau * (T[:j, :j] @ (V[:, :j].T @ V[:, j]))
    return V, T, R


def sy2sb(A, b):
    """Stage 1.  Returns the band matrix (full storage, symmetric, zero outside |i - j| <= b) and the panel factors
    [(row offset, V, T)] with Q1 = H_0 H_1 ... ."""
    A = A.copy()
    n = A.shape[0]
    panels = []
    for j in range(0, n - b - 1, b):
        r0 = j + b                                           # first row below the band for columns j .. j+b-1
        V, T, R = panel_qr(A[r0:, j:j + b])
        A[r0:, j:j + b] = R
        A[j:j + b, r0:] = R.T
        A22 = A[r0:, r0:]
        W = A22 @ V @ T                                      # "symm": one pass over the trailing matrix
        X = W - 0.5 * V @ (T.T @ (V.T @ W))
        A[r0:, r0:] = A22 - X @ V.T - V @ X.T                # "syr2k": a second pass, read + write
        panels.append((r0, V, T))
    return A, panels


def sb2st(Bd, b):
    """Stage 2: bulge chasing on the band (full storage for clarity).  Returns d, e and the reflectors
    [(first row, v, tau)] in application order, Q2 = G_0 G_1 ... ."""
    A = Bd.copy()
    n = A.shape[0]
    refl = []
    for j in range(n - 2):
        # annihilate A[j+2 : j+b+1, j]; then chase the bulge in steps of b
        col, r0 = j, j + 1
        while r0 < n - 1:
            r1 = min(r0 + b, n)
            x = A[r0:r1, col].copy()
            if np.dot(x[1:], x[1:]) == 0.0:
                break
            v, tau, beta = house(x)
            refl.append((r0, v, tau))
            # two-sided application on the rows / columns r0 .. r1-1 (touches a (2b) x b window of the band + the bulge)
            lo, hi = max(col, r0 - b), min(n, r1 + b)
            A[r0:r1, lo:hi] -= tau * np.outer(v, v @ A[r0:r1, lo:hi])
            A[lo:hi, r0:r1] -= tau * np.outer(A[lo:hi, r0:r1] @ v, v)
            col, r0 = r0, r0 + b                             # the bulge sits below the band in column r0 now
    d = np.diag(A).copy()
    e = np.diag(A, -1).copy()
    return d, e, refl


def apply_q(Z, panels, refl):
    """U = Q1 Q2 Z: Q2's reflectors in reverse order first, then the panels' WY factors in reverse order."""
    U = Z.copy()
    for r0, v, tau in reversed(refl):
        U[r0:r0 + len(v)] -= tau * np.outer(v, v @ U[r0:r0 + len(v)])
    for r0, V, T in reversed(panels):
        U[r0:] -= V @ (T @ (V.T @ U[r0:]))
    return U


def laplacian(n, seed):
    rng = np.random.default_rng(seed)
    W = np.triu((rng.random((n, n)) < 0.9) * (0.5 + 0.5 * rng.random((n, n))), 1)
    W = W + W.T
    d = W.sum(1)
    s = np.sqrt(1.0 / d)
    return (np.diag(d) - W) * s[:, None] * s[None, :]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    A = laplacian(n, 7)
    A = (A + A.T) / 2
    band, panels = sy2sb(A, b)
    off = np.abs(np.subtract.outer(np.arange(n), np.arange(n))) > b
    print("stage 1: |outside band| max %.2e, spectrum diff %.2e" % (np.abs(band[off]).max(), np.abs(np.linalg.eigvalsh(band) - np.linalg.eigvalsh(A)).max()))
    d, e, refl = sb2st(band, b)
    Tm = np.diag(d) + np.diag(e, -1) + np.diag(e, 1)
    lam, Z = np.linalg.eigh(Tm)
    lam_ref = np.linalg.eigvalsh(A)
    print("stage 2: %d reflectors of length <= %d, spectrum diff %.2e" % (len(refl), b, np.abs(lam - lam_ref).max()))
    U = apply_q(Z, panels, refl)
    print("back-transform: residual %.2e, orthonormality %.2e" % (np.abs(A @ U - U * lam).max(), np.abs(U.T @ U - np.eye(n)).max()))

    # ---- work / traffic model on the ML-10M shard of bench.py (users with n > 2048 take the two-stage path) ----
    big = np.array([7359, 4650, 3818, 3560, 3293, 3084, 2983, 2756, 2660, 2550, 2432, 2414, 2300, 2200, 2150, 2100, 2060], dtype=np.float64)
    B = 64.0
    n3, n2 = (big ** 3).sum(), (big ** 2).sum()
    one_stage_bytes = 1.375 * n3
    s1_flop = (4.0 / 3.0) * n3                               # symm + syr2k with k = 64
    s1_bytes = n3 / (3.0 * B) * 12.0                         # per panel: lower half of A22 read twice + written once (12 m^2 bytes), summed over the panels
    s2_flop = 6.0 * B * n2
    bt2_flop = (2.0 * (big ** 2) * 0.9 * big).sum()          # second back-transform: 2 n^2 k, k ~ 0.9 n for the big users
    print("\nmodel, users with n > 2048 of the bench shard (sum n^3 = %.3g, %.0f %% of the shard's 1.23e12):" % (n3, 100 * n3 / 1.23e12))
    print("  one-stage trd_kernel        : %.2f TB streamed          -> %.0f ms at the measured 3.5 TB/s, %.0f ms at 6.55 TB/s" %
          (one_stage_bytes / 1e12, one_stage_bytes / 3.5e12 * 1e3, one_stage_bytes / 6.55e12 * 1e3))
    print("  stage 1 (dense -> band 64)  : %.2f TF, %.3f TB            -> %.0f ms at 60 %% of 36 TF/s (tensor bound: %.0f flop/byte)" %
          (s1_flop / 1e12, s1_bytes / 1e12, s1_flop / (0.6 * 36e12) * 1e3, s1_flop / s1_bytes))
    print("  stage 2 (bulge chasing)     : %.3f TF, latency bound: %.0f sweeps x n/b tasks, pipelined 3 tasks apart" % (s2_flop / 1e12, big.max()))
    print("  extra back-transform (Q2)   : %.2f TF                      -> %.0f ms at 50 %% of 36 TF/s (bt_apply today: 49 %%)" %
          (bt2_flop / 1e12, bt2_flop / (0.5 * 36e12) * 1e3))


if __name__ == "__main__":
    main()
