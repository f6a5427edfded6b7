"""Checker for ONE out_eigen_ record at sizes where a full eigh with vectors is too slow to repeat many
times (test infrastructure -- used by tests/ and by bench.py's post-timed parity block, never by the
product).  Follows compute_eigens (precompute_local_threads.cpp:100-213 == precompute_local.cpp:185-261):
the Laplacian, sig_min and the cutoff are restated by oracle/gsi_oracle.py; the eigenvalues come from
LAPACK (eigvalsh, lower triangle = what SelfAdjointEigenSolver reads, precompute_local.cpp:231); the
eigenvectors are checked by what defines them (residual, orthonormality) instead of against a second set
of vectors, which is solver-dependent inside clusters.

Tolerances (the same as tests/parity.py): sig_min bit-exact, k exact (a difference is legitimate only when
an eigenvalue lies within 1e-9 of the float threshold), |d lambda| <= 1e-10, residual and orthonormality
<= 1e-9, sign convention (largest |entry| of a column positive) exact."""
from __future__ import annotations

import numpy as np

from . import gsi_oracle as O

LAM_TOL = 1e-10
VEC_TOL = 1e-9


def check_record(items, weights, sig, k, lam, vec) -> dict:
    """Returns the measured deviations of one record; ``ok`` is the conjunction of the bars above."""
    items = np.asarray(items, dtype=np.int64)
    n = len(items)
    ww = O.gather_ww(items, weights)
    _, _, ll2 = O.normalized_laplacian(ww)
    sigs_min, thr = O.sig_min_rows(ll2)
    lam_full = np.linalg.eigvalsh(ll2, UPLO="L")
    k_or = O.cutoff(lam_full, thr)
    kk = min(int(k), n)
    lam = np.asarray(lam, dtype=np.float64)
    vec = np.asarray(vec, dtype=np.float64).reshape(n, int(k))
    sig_exact = bool(np.array_equal(sigs_min, np.asarray(sig)))
    k_tie = float(np.abs(lam_full - float(thr)).min())
    k_ok = (int(k) == k_or) or k_tie < 1e-9
    dlam = float(np.abs(lam[:kk] - lam_full[:kk]).max()) if kk else 0.0
    a = np.tril(ll2) + np.tril(ll2, -1).T
    v = vec[:, :kk]
    resid = float(np.abs(a @ v - v * lam[:kk]).max()) if kk else 0.0
    orth = float(np.abs(v.T @ v - np.eye(kk)).max()) if kk else 0.0
    idx = np.argmax(np.abs(v), axis=0)
    sign_ok = bool(np.all(v[idx, np.arange(kk)] > 0)) if kk else True
    ok = sig_exact and k_ok and dlam <= LAM_TOL and resid <= VEC_TOL and orth <= VEC_TOL and sign_ok
    return {"n": n, "k": int(k), "k_oracle": int(k_or), "sig_min_bit_exact": sig_exact, "dlam": dlam, "residual": resid,
            "orthonormality": orth, "sign_convention": sign_ok, "nearest_lambda_to_threshold": k_tie, "ok": bool(ok)}


def summarise(rows: list) -> dict:
    """One JSON-able block over several records (bench.py prints it as "parity")."""
    if not rows:
        return {"users_checked": 0, "ok": None}
    return {
        "users_checked": len(rows), "n_checked": sorted((r["n"] for r in rows), reverse=True)[:8],
        "ok": bool(all(r["ok"] for r in rows)),
        "sig_min_bit_exact": bool(all(r["sig_min_bit_exact"] for r in rows)),
        "k_exact": bool(all(r["k"] == r["k_oracle"] for r in rows)),
        "max_dlam": max(r["dlam"] for r in rows), "max_residual": max(r["residual"] for r in rows),
        "max_orthonormality": max(r["orthonormality"] for r in rows),
        "sign_convention": bool(all(r["sign_convention"] for r in rows)),
        "bars": {"dlam": LAM_TOL, "residual": VEC_TOL, "orthonormality": VEC_TOL},
        "checker": "oracle/light_check.py: Laplacian / sig_min / cutoff restated from precompute_local.cpp:185-261, eigenvalues "
                   "by LAPACK eigvalsh(lower), eigenvectors by residual and orthonormality",
    }
