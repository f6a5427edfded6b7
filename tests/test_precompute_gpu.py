"""GPU parity tests of the precompute path (call through the C ABI via ctypes)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from collaborative_filtering_b200.api import Context
    c = Context(0)
    yield c
    c.close()


def _records_match_golden(recs, z):
    from tests.parity import LAM_TOL
    nu = len(z["k"])
    assert np.array_equal(recs.sig_min, z["sig_min"])             # bit-exact
    assert np.array_equal(recs.k, z["k"])
    lo = np.concatenate([[0], np.cumsum(z["k"])])
    for u in range(nu):
        lam_g = z["lam"][lo[u]:lo[u + 1]]
        assert np.abs(recs.lam_of(u) - lam_g).max() <= LAM_TOL


def test_golden_precompute_rand(ctx, golden_dir):
    from tests.parity import check_records
    z = np.load(os.path.join(golden_dir, "precompute_rand", "oracle.npz"))
    ctx.set_weights(z["weights"])
    recs = ctx.precompute(z["offsets"], z["items"])
    _records_match_golden(recs, z)
    check_records(recs, z["weights"], tag="precompute_rand")


@pytest.mark.parametrize("case", ["tiny_int", "tiny_half"])
def test_golden_pipeline_cases(ctx, golden_dir, case):
    from tests.parity import check_records
    z = np.load(os.path.join(golden_dir, case, "oracle.npz"))
    ctx.set_weights(z["weights"])
    recs = ctx.precompute(z["offsets"], z["items"])
    _records_match_golden(recs, z)
    check_records(recs, z["weights"], tag=case)


def test_weights_from_edges_equals_dense(ctx, golden_dir):
    z = np.load(os.path.join(golden_dir, "tiny_int", "oracle.npz"))
    w = z["weights"]
    a, b = np.nonzero(w)
    rows = ctx.set_weights_edges(a, b, w[a, b])
    assert rows == w.shape[0]
    r1 = ctx.precompute(z["offsets"], z["items"])
    ctx.set_weights(w)
    r2 = ctx.precompute(z["offsets"], z["items"])
    assert np.array_equal(r1.sig_min, r2.sig_min) and np.array_equal(r1.k, r2.k)
    assert np.array_equal(r1.lam, r2.lam) and np.array_equal(r1.vec, r2.vec)       # deterministic


def test_ml100k_shape_sample_vs_oracle(ctx):
    """Seeded ML-100K shaped users against the oracle: every size bucket of the
    CTA-resident solver and the Householder path (n > 160)."""
    from collaborative_filtering_b200 import datasets as D
    from tests.parity import check_records
    r = D.make_ratings("ml-100k")
    w = D.make_weights(r.n_items)
    ctx.set_weights(w)
    recs = ctx.precompute(r.offsets, r.items)
    deg = r.degrees()
    order = np.argsort(deg)
    pick = sorted(set(order[:: max(1, len(order) // 40)].tolist() + order[-4:].tolist()))
    check_records(recs, w, users=pick, tag="ml-100k")
    # size-independent properties on ALL users: k in [2, max(n,2)], lam ascending in [0,2], unit norm
    for u in range(r.n_users):
        lam, vec, n, k = recs.lam_of(u), recs.vec_of(u), int(deg[u]), int(recs.k[u])
        kk = min(k, n)
        assert 2 <= k <= max(n, 2)
        assert np.all(np.diff(lam[:kk]) >= -1e-12) and lam[0] > -1e-10 and lam[kk - 1] < 2 + 1e-10
        assert np.abs((vec * vec).sum(0)[:kk] - 1).max() < 1e-10
        assert abs(lam[0]) < 1e-10                             # every graph has the null vector


def test_device_api_matches_host_api(ctx):
    import torch
    from collaborative_filtering_b200 import datasets as D
    from collaborative_filtering_b200.api import GsiError, upper_bounds
    r = D.make_ratings("ml-100k", n_users=200)
    w = D.make_weights(r.n_items)
    dev = torch.device("cuda:0")
    d_w = torch.from_numpy(w).to(dev)
    ctx.set_weights(d_w)
    host = ctx.precompute(r.offsets, r.items)
    lam_cap, vec_cap = upper_bounds(r.offsets)
    d_items = torch.from_numpy(r.items).to(dev)
    d_sig = torch.zeros(r.nnz, dtype=torch.float64, device=dev)
    d_k = torch.zeros(r.n_users, dtype=torch.int32, device=dev)
    d_lo = torch.zeros(r.n_users, dtype=torch.int64, device=dev)
    d_vo = torch.zeros(r.n_users, dtype=torch.int64, device=dev)
    d_lam = torch.zeros(lam_cap, dtype=torch.float64, device=dev)
    d_vec = torch.zeros(vec_cap, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    used = ctx.precompute_device(r.offsets, d_items, d_sig, d_k, d_lo, d_vo, d_lam, d_vec)
    assert used == (len(host.lam), len(host.vec))
    assert np.array_equal(d_sig.cpu().numpy(), host.sig_min)
    assert np.array_equal(d_k.cpu().numpy(), host.k)
    lo, vo, lam, vec = d_lo.cpu().numpy(), d_vo.cpu().numpy(), d_lam.cpu().numpy(), d_vec.cpu().numpy()
    for u in range(r.n_users):
        n, k = host.n(u), int(host.k[u])
        assert np.array_equal(lam[lo[u]:lo[u] + k], host.lam_of(u))
        assert np.array_equal(vec[vo[u]:vo[u] + n * k].reshape(n, k), host.vec_of(u))
    small = torch.zeros(10, dtype=torch.float64, device=dev)    # capacity error path
    with pytest.raises(GsiError) as e:
        ctx.precompute_device(r.offsets, d_items, d_sig, d_k, d_lo, d_vo, d_lam, small)
    assert e.value.code == 5


def test_error_paths(ctx):
    from collaborative_filtering_b200.api import Context, GsiError
    c2 = Context(0)
    with pytest.raises(GsiError) as e:                         # no weights yet
        c2.precompute(np.array([0, 2]), np.array([1, 2]))
    assert e.value.code == 4
    c2.set_weights(np.zeros((4, 4)))
    with pytest.raises(GsiError) as e:                         # duplicate / unsorted ids
        c2.precompute(np.array([0, 2]), np.array([2, 2]))
    assert e.value.code == 1
    with pytest.raises(GsiError):                              # empty user
        c2.precompute(np.array([0, 0, 1]), np.array([1]))
    c2.close()


def test_large_path_vs_oracle(ctx):
    """Householder / divide & conquer path at sizes the oracle still finishes in seconds (n = 161 .. 700: shared-memory and
    persistent tridiagonalisation kernels): mixed sizes in one call, ids beyond the table and an isolated item."""
    from tests.parity import check_records
    rng = np.random.default_rng(31413)
    n_items = 900
    w = np.round(1.0 - 0.5 * rng.random((n_items + 1, n_items + 1)), 6)
    w = np.where(rng.random(w.shape) < 0.8, w, 0.0)
    w = np.triu(w, 1)
    w = w + w.T
    w[0] = 0
    w[:, 0] = 0
    w[5] = 0
    w[:, 5] = 0
    sizes = [161, 176, 177, 255, 256, 257, 300, 511, 513, 700]
    lists = [np.sort(rng.choice(np.arange(1, n_items + 40), n, replace=False)).astype(np.int32) for n in sizes]
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    items = np.concatenate(lists)
    ctx.set_weights(w)
    recs = ctx.precompute(offsets, items)
    check_records(recs, w, tag="large")


def test_block_jacobi_route_vs_oracle():
    """The first large path (kern_bj.cuh, one-sided block Jacobi) is still reachable behind GSI_LARGE=bj (comparison runs) and
    for n > 9,216 when the two-stage path is switched off (GSI_SBR_MIN=0): it must stay correct.  Same records as the
    Householder path to the parity bars (k exact, eigenvalues 1e-10, residual / orthonormality 1e-9, sig_min bit-exact)."""
    from tests.parity import check_records
    rng = np.random.default_rng(77)
    n_items = 500
    w = np.round(1.0 - 0.5 * rng.random((n_items + 1, n_items + 1)), 6)
    w = np.where(rng.random(w.shape) < 0.8, w, 0.0)
    w = np.triu(w, 1)
    w = w + w.T
    w[0] = 0
    w[:, 0] = 0
    sizes = [161, 200, 257, 300, 64, 12]
    lists = [np.sort(rng.choice(np.arange(1, n_items + 1), n, replace=False)).astype(np.int32) for n in sizes]
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    items = np.concatenate(lists)
    c = _fresh_context({"GSI_LARGE": "bj"})
    try:
        c.set_weights(w)
        recs = c.precompute(offsets, items)
        check_records(recs, w, tag="bj")
    finally:
        c.close()


def _fresh_context(env):
    """A context created under temporary environment settings (read by gsi_create)."""
    from collaborative_filtering_b200.api import Context
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return Context(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_paths_agree_across_the_crossover(golden_dir):
    """Users with 81 <= n <= 160 can take either eigensolver (GSI_SMALL_MAX moves the crossover): CTA-resident Jacobi
    and Householder + divide & conquer must produce the same records (k exact, eigenvalues 1e-10, kept subspace 1e-8)."""
    from collaborative_filtering_b200 import datasets as D
    from tests.parity import check_records
    r = D.make_ratings("ml-100k")
    deg = np.diff(r.offsets)
    sel = np.nonzero((deg > 80) & (deg <= 160))[0][:40]
    assert len(sel) >= 20
    _, off, items, _ = D.subset(r, sel)
    w = D.make_weights(r.n_items)
    recs = {}
    for name, env in (("jacobi", {"GSI_SMALL_MAX": "160"}), ("householder", {"GSI_SMALL_MAX": "80"})):
        c = _fresh_context(env)
        try:
            assert c.small_max == int(env["GSI_SMALL_MAX"])
            c.set_weights(w)
            recs[name] = c.precompute(off, items)
        finally:
            c.close()
    a, b = recs["jacobi"], recs["householder"]
    assert np.array_equal(a.sig_min, b.sig_min) and np.array_equal(a.k, b.k)
    for u in range(len(sel)):
        assert np.abs(a.lam_of(u) - b.lam_of(u)).max() <= 1e-10
        pa, pb = a.vec_of(u) @ a.vec_of(u).T, b.vec_of(u) @ b.vec_of(u).T
        assert np.abs(pa - pb).max() <= 1e-8
    check_records(b, w, users=range(0, len(sel), 4), tag="householder 81..160")


def test_small_workspace_splits_into_chunks(ctx):
    """A 64 MiB workspace forces the Householder path through several chunks; the records must not depend on it."""
    from collaborative_filtering_b200 import datasets as D
    r = D.make_ratings("ml-100k")
    deg = np.diff(r.offsets)
    sel = np.nonzero((deg > 200) & (deg <= 420))[0][:24]
    _, off, items, _ = D.subset(r, sel)
    w = D.make_weights(r.n_items)
    ctx.set_weights(w)
    ctx.set_workspace_limit(8 << 30)
    one = ctx.precompute(off, items)
    ctx.set_workspace_limit(64 << 20)
    many = ctx.precompute(off, items)
    ctx.set_workspace_limit(8 << 30)
    assert np.array_equal(one.k, many.k) and np.array_equal(one.sig_min, many.sig_min)
    for u in range(len(sel)):
        assert np.array_equal(one.lam_of(u), many.lam_of(u))          # fixed reduction orders: bit-identical
        assert np.array_equal(one.vec_of(u), many.vec_of(u))


def test_host_path_survives_timing_reset_and_close(golden_dir):
    """The host path copies the eigenvector blocks on a second stream; resetting the timers between two calls and closing
    the context afterwards must leave that stream alone (regression: the reset destroyed it)."""
    from collaborative_filtering_b200 import datasets as D
    from collaborative_filtering_b200.api import Context
    r = D.make_ratings("ml-100k", n_users=60)
    w = D.make_weights(r.n_items)
    c = Context(0)
    try:
        c.set_weights(w)
        a = c.precompute(r.offsets, r.items)
        c.timing_enable(True)
        c.timing_reset()
        b = c.precompute(r.offsets, r.items)
        c.timing_reset()
        assert np.array_equal(a.vec, b.vec) and np.array_equal(a.lam, b.lam) and np.array_equal(a.sig_min, b.sig_min)
    finally:
        c.close()
