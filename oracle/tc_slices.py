"""TEST INFRASTRUCTURE (oracle): numpy restatement of the INT8-slice scheme of csrc/tc_gemm.cu, integer for integer.

The product path never imports this file.  It exists so that the arithmetic of the tensor-core engine -- exponent per row /
column, one rounding to a 7 S-bit integer, balanced base-128 digits, exact INT32 accumulation of the slice pairs with
t + u = g, Horner combination in FP64, the two scale factors -- can be checked on the CPU against a plain FP64 product and
against the stated error bound, independently of the GPU (tests/test_tc_slices.py); the GPU kernel is then checked against
numpy's product in tests/test_tc_gemm_gpu.py.  The GEMMs it stands in for are the merge products of the divide and conquer
that replaces the QR-iteration half of Eigen's SelfAdjointEigenSolver (precompute_local.cpp:231).
"""
import numpy as np

EXP_NONE = -100000


def line_exponents(x: np.ndarray, axis: int) -> np.ndarray:
    """e with |x| 2^-e < 1/2 along `axis` (tc_elem_exp / tc_line_exp: frexp exponent + 1, 0 for an all-zero line)."""
    amax = np.abs(x).max(axis=axis)
    _, e = np.frexp(amax)
    return np.where(amax > 0, e + 1, 0).astype(np.int64)


def digits(x: np.ndarray, e: np.ndarray, s: int) -> np.ndarray:
    """tc_digits: X = rint(x 2^(7S - e)) clamped to +-(2^(7S-1) - 1), then S balanced base-128 digits, most significant first.
    Returns int64 array of shape (S,) + x.shape with entries in [-65, 65]."""
    big = np.rint(np.ldexp(x, (7 * s - e).astype(np.int64))).astype(np.int64) if s <= 8 else None
    lim = (1 << (7 * s - 1)) - 1
    big = np.clip(big, -lim, lim)
    d = np.zeros((s,) + x.shape, dtype=np.int64)
    for t in range(s - 1, 0, -1):
        r = ((big + 64) & 127) - 64
        d[t] = r
        big = (big - r) >> 7
    d[0] = big
    return d


def gemm(a: np.ndarray, b: np.ndarray, s: int = 8) -> np.ndarray:
    """C = A @ B the way tc_gemm_kernel<S> computes it."""
    m, k = a.shape
    k2, n = b.shape
    assert k == k2
    ea = line_exponents(a, 1)                    # rows of A
    eb = line_exponents(b, 0)                    # columns of B
    da = digits(a, ea[:, None], s)               # (S, m, k)
    db = digits(b, eb[None, :], s)               # (S, k, n)
    acc = np.zeros((m, n))
    for g in range(s - 1, -1, -1):               # Horner: acc = acc / 128 + G_g
        gg = np.zeros((m, n), dtype=np.int64)
        for t in range(g + 1):
            gg += da[t] @ db[g - t]              # exact: |digits| <= 65, K (g + 1) 65^2 < 2^31 for K < 63,000 (the INT32 accumulator)
        assert np.abs(gg).max() < 2 ** 31
        acc = acc * 0.0078125 + gg.astype(np.float64)
    rs = np.ldexp(1.0, (ea - 7 - 7 * s).astype(np.int64))
    cs = np.ldexp(1.0, (eb - 7 + 7 * s).astype(np.int64))
    return acc * rs[:, None] * cs[None, :]


def bound(a: np.ndarray, b: np.ndarray, s: int) -> np.ndarray:
    """The bound tests/test_tc_gemm_gpu.py states: (S + 2) 2^(-7S + 2) K max_k|A_ik| max_k|B_kj|."""
    k = a.shape[1]
    return (s + 2) * 2.0 ** (-7 * s + 2) * k * np.abs(a).max(1)[:, None] * np.abs(b).max(0)[None, :]
