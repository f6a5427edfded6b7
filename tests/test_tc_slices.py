"""CPU checks of the INT8-slice arithmetic behind the tensor-core GEMM (oracle/tc_slices.py restates csrc/tc_gemm.cu integer for
integer): digit ranges, exact reconstruction, the INT32 accumulator bound, the stated error bound against numpy's FP64 product for
S = 6, 7, 8 on badly scaled operands, exactness on small integers, and FP64-level accuracy on orthogonal blocks (the operands of the
divide-and-conquer merges).  The GPU kernel meets the same bound in tests/test_tc_gemm_gpu.py."""
import numpy as np
import pytest

from oracle import tc_slices as T


@pytest.mark.parametrize("s", [6, 7, 8])
def test_digits_reconstruct_the_rounded_value(s):
    rng = np.random.default_rng(s)
    x = rng.standard_normal((40, 50)) * np.exp(rng.uniform(-8, 8, size=(40, 1)))
    x[3] = 0.0
    x[5, 7] = 1.0
    e = T.line_exponents(x, 1)
    assert e[3] == 0
    assert np.all(np.abs(x) * np.ldexp(1.0, -e)[:, None] < 0.5)
    d = T.digits(x, e[:, None], s)
    assert d.min() >= -65 and d.max() <= 65 and np.abs(d[1:]).max() <= 64
    big = sum(d[t].astype(object) * 128 ** (s - 1 - t) for t in range(s))
    ref = np.rint(np.ldexp(x, (7 * s - e)[:, None])).astype(np.int64)
    lim = (1 << (7 * s - 1)) - 1
    assert np.array_equal(np.array(big, dtype=np.int64), np.clip(ref, -lim, lim))
    # one rounding: |x - X 2^(e - 7S)| <= 2^(e - 7S - 1) (+ the clamp's single unit at |x| 2^-e = 1/2)
    back = np.ldexp(np.array(big, dtype=np.float64), (e - 7 * s)[:, None])
    assert np.all(np.abs(back - x) <= np.ldexp(1.0, (e - 7 * s))[:, None])


@pytest.mark.parametrize("s", [6, 7, 8])
@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (17, 9, 33), (64, 48, 300)])
def test_error_bound(s, m, n, k):
    rng = np.random.default_rng(100 * s + m + n + k)
    a = rng.standard_normal((m, k)) * np.exp(rng.uniform(-6, 6, size=(m, 1)))
    b = rng.standard_normal((k, n)) * np.exp(rng.uniform(-6, 6, size=(1, n)))
    c = T.gemm(a, b, s)
    assert np.all(np.abs(c - a @ b) <= T.bound(a, b, s) + 1e-300)


def test_small_integers_are_exact():
    rng = np.random.default_rng(3)
    a = rng.integers(-50, 51, size=(30, 70)).astype(np.float64)
    b = rng.integers(-50, 51, size=(70, 20)).astype(np.float64)
    assert np.array_equal(T.gemm(a, b, 8), a @ b)


def test_orthogonal_blocks_reach_fp64():
    rng = np.random.default_rng(7)
    q, _ = np.linalg.qr(rng.standard_normal((300, 300)))
    z, _ = np.linalg.qr(rng.standard_normal((300, 120)))
    a = q[:200]
    ref = (a.astype(np.longdouble) @ z.astype(np.longdouble)).astype(np.float64)
    err8 = np.abs(T.gemm(a, z, 8) - ref).max()
    err_np = np.abs(a @ z - ref).max()
    assert err8 <= 4e-16 and err8 <= 4 * err_np + 1e-17
    assert np.abs(T.gemm(a, z, 7) - ref).max() <= 1e-13
    assert np.abs(T.gemm(a, z, 6) - ref).max() <= 1e-11


def test_int32_accumulator_headroom():
    """Worst case of one accumulator: (g + 1) K 65^2 with g + 1 <= 8 -- the limit on K the engine's callers stay below."""
    assert 8 * 60000 * 65 * 65 < 2 ** 31
