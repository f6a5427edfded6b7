"""Parity on the configuration bench.py measures: the ML-10M-shaped heavy tail goes
through gsi_precompute_host in ONE mixed batch that lands in every size class of the large path (users above
4,096 / 2,048 / 1,024 rated movies and the filler below), plus the CTA-resident small path -- the planner,
the team levels of the tridiagonalisation and the grouped back-transform are exactly the ones the bench
runs.  Checker: oracle/light_check.py (precompute_local.cpp:185-261)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_ITEMS = 10681          # ML-10M item universe (BASELINE.json configs[2])


@pytest.fixture(scope="module")
def weights():
    from collaborative_filtering_b200 import datasets as D
    return D.make_weights(N_ITEMS)


@pytest.fixture(scope="module")
def ctx(weights):
    from collaborative_filtering_b200.api import Context
    c = Context(0)
    c.set_workspace_limit(48 << 30)
    c.set_weights(weights)
    yield c
    c.close()


def _batch(sizes, seed):
    rng = np.random.default_rng(seed)
    offsets = np.zeros(len(sizes) + 1, dtype=np.int64)
    np.cumsum(sizes, out=offsets[1:])
    items = np.concatenate([np.sort(rng.choice(N_ITEMS, size=n, replace=False) + 1) for n in sizes]).astype(np.int32)
    return offsets, items


def _check(recs, weights, users):
    from oracle.light_check import check_record, summarise
    rows = []
    for u in users:
        it = recs.items[recs.offsets[u]: recs.offsets[u + 1]]
        r = check_record(it, weights, recs.sig_of(u), int(recs.k[u]), recs.lam_of(u), recs.vec_of(u))
        assert r["ok"], "user %d: %s" % (u, r)
        rows.append(r)
    return summarise(rows)


def test_mixed_batch_all_team_levels(ctx, weights):
    """n = 7,359 (the ML-10M maximum) and 4,200 -> teams of 37 CTAs; 2,300 -> 6; 1,100 -> 2; 600 and the filler
    -> single CTAs / trd_small; 20-item users -> eig_cta_kernel.  Every heavy user and a sample of the filler is
    checked: sig_min bit-exact, k exact, |d lam| <= 1e-10, residual / orthonormality <= 1e-9."""
    sizes = [7359, 4200, 2300, 1100, 600] + [200] * 30 + [20] * 200
    offsets, items = _batch(sizes, 7)
    recs = ctx.precompute(offsets, items)
    s = _check(recs, weights, [0, 1, 2, 3, 4, 5, 20, 34, 35, 120, 234])
    assert s["ok"] and s["sig_min_bit_exact"] and s["k_exact"]
    # size-independent properties on every record
    for u, n in enumerate(sizes):
        lam, vec, k = recs.lam_of(u), recs.vec_of(u), int(recs.k[u])
        assert 2 <= k <= n and np.all(np.diff(lam) >= -1e-12) and lam[0] > -1e-10 and lam[-1] < 2 + 1e-10
        assert np.abs((vec * vec).sum(0) - 1).max() < 1e-10


def test_user_above_9216(ctx, weights):
    """One user beyond the shared-memory vectors of the one-stage tridiagonalisation kernel (n > 9,216: the Netflix
    tail, precompute_local.cpp:231 solves any n with the same call) through whatever route the planner picks."""
    offsets, items = _batch([9300, 300, 25], 11)
    recs = ctx.precompute(offsets, items)
    s = _check(recs, weights, [0, 1, 2])
    assert s["ok"]


@pytest.mark.skipif(not __import__("os").environ.get("GSI_TEST_NETFLIX_MAX"), reason="minutes of LAPACK on the host: set GSI_TEST_NETFLIX_MAX=1")
def test_user_at_the_netflix_maximum():
    """BASELINE.json configs[3]: the heaviest Netflix-shaped user rates 17,653 of 17,770 movies.  One such user (plus filler)
    through gsi_precompute_host: sig_min bit-exact, k exact, |d lam| <= 1e-10, residual / orthonormality <= 1e-9."""
    from collaborative_filtering_b200 import datasets as D
    from collaborative_filtering_b200.api import Context
    n_items = 17770
    w = D.make_weights(n_items)
    c = Context(0)
    try:
        c.set_workspace_limit(64 << 30)
        c.set_weights(w)
        rng = np.random.default_rng(13)
        sizes = [17653, 400, 30]
        offsets = np.zeros(len(sizes) + 1, dtype=np.int64)
        np.cumsum(sizes, out=offsets[1:])
        items = np.concatenate([np.sort(rng.choice(n_items, size=n, replace=False) + 1) for n in sizes]).astype(np.int32)
        recs = c.precompute(offsets, items)
        s = _check(recs, w, [0, 1, 2])
        assert s["ok"] and s["sig_min_bit_exact"] and s["k_exact"], s
        print("netflix-max user:", s)
    finally:
        c.close()


def test_records_do_not_depend_on_the_batch(ctx, weights):
    """SURVEY.md section 7 test (h) on one device: the records of a user are bit-identical whether it is solved alone,
    in the whole batch, or in either half of a 2-way LPT shard (what a 2-GPU run computes per rank)."""
    from collaborative_filtering_b200 import shard as SH
    sizes = [1500, 1100, 700, 300, 300, 130, 90, 50, 33, 20, 8, 3, 2, 1]
    offsets, items = _batch(sizes, 3)
    whole = ctx.precompute(offsets, items)
    owners = SH.lpt_assign(np.diff(offsets), 2)
    for r in range(2):
        idx = np.nonzero(owners == r)[0]
        off = np.zeros(len(idx) + 1, dtype=np.int64)
        np.cumsum(np.diff(offsets)[idx], out=off[1:])
        it = np.concatenate([items[offsets[u]: offsets[u + 1]] for u in idx])
        part = ctx.precompute(off, it)
        for j, u in enumerate(idx):
            assert part.k[j] == whole.k[u]
            assert np.array_equal(part.sig_of(j), whole.sig_of(u))
            assert np.array_equal(part.lam_of(j), whole.lam_of(u)), "lam of user %d depends on the batch" % u
            assert np.array_equal(part.vec_of(j), whole.vec_of(u)), "vec of user %d depends on the batch" % u


def test_two_gpu_records_identical(weights):
    """Hardware version of the test above (needs 2 devices: `gpurun --gpus 2`): rank-sharded records on device 0 and
    device 1 equal the single-device records bit for bit."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from collaborative_filtering_b200 import shard as SH
    from collaborative_filtering_b200.api import Context
    sizes = [2500, 1500, 1100, 700, 300, 300, 130, 90, 50, 33, 20, 8, 3, 2, 1]
    offsets, items = _batch(sizes, 5)
    c0, c1 = Context(0), Context(1)
    try:
        c0.set_weights(weights)
        c1.set_weights(weights)
        whole = c0.precompute(offsets, items)
        owners = SH.lpt_assign(np.diff(offsets), 2)
        for r, c in enumerate((c0, c1)):
            idx = np.nonzero(owners == r)[0]
            off = np.zeros(len(idx) + 1, dtype=np.int64)
            np.cumsum(np.diff(offsets)[idx], out=off[1:])
            it = np.concatenate([items[offsets[u]: offsets[u + 1]] for u in idx])
            part = c.precompute(off, it)
            for j, u in enumerate(idx):
                assert part.k[j] == whole.k[u]
                assert np.array_equal(part.sig_of(j), whole.sig_of(u))
                assert np.array_equal(part.lam_of(j), whole.lam_of(u))
                assert np.array_equal(part.vec_of(j), whole.vec_of(u))
    finally:
        c0.close()
        c1.close()
