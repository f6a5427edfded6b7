"""GPU parity of the FP64-equivalent tensor-core GEMM (csrc/tc_gemm.cu: INT8 slices on tcgen05.mma.kind::i8, INT32
accumulators in tensor memory) against numpy's FP64 product -- the checker of the engine behind the divide-and-conquer
merges (the QR-iteration half of Eigen's SelfAdjointEigenSolver, precompute_local.cpp:231).  Through the C ABI test hook
gsi_debug_tc_gemm.

Tolerance (stated, not fitted): every operand entry is rounded once to 7 S bits below its row / column exponent and slice
pairs below 128^-S are dropped, everything else is exact integer arithmetic, so
    |C - A B|_ij  <=  (S + 2) 2^(-7 S + 2) K max_k|A_ik| max_k|B_kj|
(a bound; the observed error is ~sqrt(K) smaller).  For S = 8 that is 2.2e-16 K: the FP64 rounding level of a K-long sum."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from collaborative_filtering_b200.api import Context
    c = Context(0)
    yield c
    c.close()


def _bound(a, b, s):
    k = a.shape[1]
    return (s + 2) * 2.0 ** (-7 * s + 2) * k * np.abs(a).max(1)[:, None] * np.abs(b).max(0)[None, :]


@pytest.mark.parametrize("m,n,k", [(128, 64, 64), (1, 1, 1), (130, 70, 33), (300, 200, 500), (1000, 700, 2048), (257, 129, 4100)])
@pytest.mark.parametrize("s", [8, 7, 6])
def test_random_matrices(ctx, m, n, k, s):
    rng = np.random.default_rng(100 * m + n + k + s)
    a = rng.standard_normal((m, k)) * np.exp(rng.uniform(-6, 6, size=(m, 1)))      # rows of very different scale
    b = rng.standard_normal((k, n)) * np.exp(rng.uniform(-6, 6, size=(1, n)))
    c, _, _ = ctx.debug_tc_gemm(a, b, slices=s)
    ref = a @ b
    err = np.abs(c - ref)
    assert np.all(err <= _bound(a, b, s) + 1e-300), float((err / (_bound(a, b, s) + 1e-300)).max())


def test_orthogonal_blocks_reach_fp64(ctx):
    """The operands of the D&C merges: a block of an orthogonal matrix times normalised vectors.  S = 8 must be at the level of
    the FP64 product itself (compared in extended precision through a compensated reference)."""
    rng = np.random.default_rng(7)
    q, _ = np.linalg.qr(rng.standard_normal((1500, 1500)))
    s, _ = np.linalg.qr(rng.standard_normal((1500, 900)))
    a = q[:1200, :]
    c, _, _ = ctx.debug_tc_gemm(a, s, slices=8)
    ref = (a.astype(np.longdouble) @ s.astype(np.longdouble)).astype(np.float64)
    err_tc = np.abs(c - ref).max()
    err_np = np.abs(a @ s - ref).max()
    assert err_tc <= 4e-15 and err_tc <= 8 * err_np + 1e-16, (err_tc, err_np)


def test_exact_on_small_integers(ctx):
    """Integer matrices whose entries fit the top digits are reproduced exactly (the slices are error free)."""
    rng = np.random.default_rng(3)
    a = rng.integers(-50, 51, size=(200, 300)).astype(np.float64)
    b = rng.integers(-50, 51, size=(300, 100)).astype(np.float64)
    c, _, _ = ctx.debug_tc_gemm(a, b, slices=8)
    assert np.array_equal(c, a @ b)


def test_zero_rows_and_identity(ctx):
    a = np.zeros((140, 70)); a[5, 3] = 1.0; a[139, 69] = -2.5
    b = np.eye(70)
    c, _, _ = ctx.debug_tc_gemm(a, b, slices=8)
    assert np.array_equal(c, a)


@pytest.mark.parametrize("n,sbr", [(700, False), (1100, False), (1537, False), (1100, True)])
def test_back_transform_on_the_engine(ctx, monkeypatch, n, sbr):
    """Back-transformation with aggregated 512-reflector panels on the tcgen05 engine (kern_bt_tc.cuh: G = V^T V, T from G,
    VT = V T, then X = V^T Z and Z -= VT X per super-panel) instead of the 64-reflector DMMA kernel: eigenpairs of a normalised
    Laplacian through gsi_debug_eigh with the switch-over lowered (GSI_BT_TC_MIN), one-stage and two-stage reflector layouts,
    sizes with a short last super-panel.  Same bars as the stage tests: eigenvalues 1e-12, residual / orthonormality 1e-11."""
    from oracle import gsi_oracle as O
    monkeypatch.setenv("GSI_BT_TC_MIN", "600")
    if sbr:
        monkeypatch.setenv("GSI_SBR_MIN", "600")
    rng = np.random.default_rng(5 + n)
    w = np.triu((rng.random((n, n)) < 0.8) * (0.5 + 0.5 * rng.random((n, n))), 1)
    w = w + w.T
    _, _, ll2 = O.normalized_laplacian(w)
    a = np.tril(ll2) + np.tril(ll2, -1).T
    lam_ref, _ = O.eig_lower(a)
    thr = float(np.float32(np.median(lam_ref) + 0.013))
    r = ctx.debug_eigh(a, thr=thr, team=0)
    k_ref = max(2, int((lam_ref <= np.float32(thr)).sum()))
    assert r["k"] == k_ref and np.abs(r["lam"] - lam_ref).max() < 1e-12
    u = r["u"]
    assert np.abs(a @ u - u * r["lam"][:k_ref]).max() < 1e-11
    assert np.abs(u.T @ u - np.eye(k_ref)).max() < 1e-11
    monkeypatch.setenv("GSI_BT_TC_MIN", "0")                         # the DMMA kernel on the same matrix: same vectors up to rounding
    r2 = ctx.debug_eigh(a, thr=thr, team=0)
    assert np.abs(np.abs(r2["u"]) - np.abs(u)).max() < 1e-9
