"""CPU tests of the oracle's restatement of the per-movie variant (local_calc.cpp:262-526):
golden reproducibility, the reference's column-0 fix-up, and the identities the exact cutoff satisfies
(w_lim is the smallest singular value of the unrated rows of L; it never exceeds the norm of any of those rows,
which is what lets the GPU path truncate the per-movie spectrum at the largest row norm)."""
import os

import numpy as np
import pytest

from oracle import gsi_oracle as O


def _random_case(seed, n_items=40, n_users=30, density=0.5):
    rng = np.random.default_rng(seed)
    fin = []
    for a in range(1, n_items + 1):
        for b in range(a + 1, n_items + 1):
            if rng.random() < density:
                w = float("%g" % np.float32(rng.uniform(0.05, 1.0)))
                fin.append((a, b, w))
                fin.append((b, a, w))
    test = {}
    for u in range(1, n_users + 1):
        its = rng.choice(np.arange(1, n_items + 1), size=int(rng.integers(3, 15)), replace=False)
        for m in its:
            test.setdefault(int(m), {})[O.UIMAX - u] = float(rng.integers(1, 6))
    return fin, test


@pytest.mark.parametrize("case", ["tiny_int", "tiny_half"])
def test_golden_reproducible(golden_dir, case):
    d = os.path.join(golden_dir, case)
    z = np.load(os.path.join(d, "local_calc.npz"))
    fin = O.parse_fin(open(os.path.join(d, "out_fin_1_of_1")).read())
    test_rat = O.parse_rat(open(os.path.join(d, "out_test_rat_1_of_1")).read())
    rows = O.local_calc(fin, test_rat)
    assert [r[0] for r in rows] == z["movie"].tolist() and [r[1] for r in rows] == z["user"].tolist()
    assert [r[3] for r in rows] == z["kk"].tolist() and [r[5] for r in rows] == z["status"].tolist()
    assert [r[6] for r in rows] == z["lim"].tolist()
    ok = (z["status"] == O.PRED_OK) & (z["gap"] > 1e-6)
    assert np.abs(np.array([r[4] for r in rows])[ok] - z["pred"][ok]).max() <= 1e-9
    assert np.abs(np.array([r[7] for r in rows])[ok] - z["w_lim"][ok]).max() <= 1e-12
    assert O.format_res(rows) == open(os.path.join(d, "out_res_local_calc")).read()


def test_local_graph_fixup_and_scope():
    # 1 -> {2, 3}; 2 -> {1, 3, 9}; 3 -> {2}; 9 is outside the local graph of 1; weights deliberately asymmetric
    fin = [(1, 2, 0.5), (1, 3, 0.7), (2, 1, 0.9), (2, 3, 0.3), (2, 9, 0.8), (3, 2, 0.4), (3, 1, 0.05), (9, 2, 0.8)]
    gw = O.item_graph_weights(fin)
    assert 1 not in gw[3]                                            # (float)0.05 > 0.1 is false: no edge
    nodes, ww = O.local_graph(1, gw)
    assert nodes == [1, 2, 3]
    f = lambda x: float(np.float32(x))
    expect = np.array([[0, f(0.5), f(0.7)], [f(0.5), 0, f(0.3)], [f(0.7), f(0.4), 0]])
    assert np.array_equal(ww, expect)                                # column 0 = row 0 = w(1 -> i), local_calc.cpp:331-333
    assert O.local_calc_movie(9, gw, {9: {5: 3.0}}) == []            # fewer than 3 nodes: no output (:271-272)


@pytest.mark.parametrize("seed,density", [(1, 0.5), (2, 0.15), (3, 0.9)])
def test_cutoff_identities(seed, density):
    fin, test = _random_case(seed, density=density)
    gw = O.item_graph_weights(fin)
    seen = 0
    for m in sorted(test):
        nodes, ww = O.local_graph(m, gw)
        if len(nodes) < 3:
            continue
        _, _, ll2 = O.normalized_laplacian(ww)
        row_norm = np.sqrt((ll2 * ll2).sum(axis=1))
        for (mm, u, err, kk, pred, status, lim, w_lim, gap) in O.local_calc_movie(m, gw, test):
            if kk == 0:
                assert status == O.PRED_EMPTY and np.isnan(pred) and np.isnan(err)
                continue
            rated = np.array([v != m and u in test.get(v, {}) for v in nodes])
            assert rated.sum() == kk
            sv = np.linalg.svd(ll2[~rated, :], compute_uv=False)
            assert abs(sv.min() - w_lim) <= 1e-7 * max(1.0, sv.min())
            assert w_lim <= row_norm[~rated].min() + 1e-12
            if lim > 2:
                assert lim <= kk and status != O.PRED_UNDERDETERMINED       # uniqueness set: the Gram has full rank
            seen += 1
    assert seen > 20


def test_asymmetric_table_is_tolerance_level():
    """The reference reads whichever direction of an edge pair its hash order visits (lower triangle of a
    non-symmetric ll2); a table whose directions differ at the 1e-3 level moves w_lim by about as much --
    with the bit-symmetric weights knn2 produces the order is immaterial."""
    fin, test = _random_case(5)
    rng = np.random.default_rng(99)
    fin_a = [(a, b, float("%g" % (w * (1.0 + 1e-3 * rng.standard_normal()))) if (a > b and w > 0.12) else w) for a, b, w in fin]
    r0 = O.local_calc(fin, test)
    r1 = O.local_calc(fin_a, test)
    assert [(r[0], r[1], r[3]) for r in r0] == [(r[0], r[1], r[3]) for r in r1]
    d = [abs(a[7] - b[7]) for a, b in zip(r0, r1) if a[3] > 0]
    assert max(d) < 5e-2


def _csr(test_rat):
    by_user = {}
    for m, d in test_rat.items():
        for u, r in d.items():
            by_user.setdefault(u, []).append((m, r))
    users = sorted(by_user)
    offsets, items, ratings = [0], [], []
    for u in users:
        for m, r in sorted(by_user[u]):
            items.append(m)
            ratings.append(r)
        offsets.append(len(items))
    return users, np.array(offsets, dtype=np.int64), np.array(items, dtype=np.int32), np.array(ratings, dtype=np.float64)


@pytest.mark.parametrize("which,honest,threads", [("tiny_int", True, 1), ("tiny_half", False, 3), ("random", True, 2)])
def test_cpp_restatement_agrees_with_numpy_oracle(golden_dir, which, honest, threads):
    """Two independent restatements of local_calc.cpp (numpy/LAPACK and the C++ Householder + QL solver pair with the
    reference's dense inverse and products) agree: emitted pairs, kk, status classes and lim exactly, w_lim to 1e-10,
    predictions to 1e-7 on well-posed pairs whose cutoff does not split a cluster."""
    from oracle import cpu_ref
    if which == "random":
        fin, test = _random_case(2, n_items=60, n_users=25, density=0.15)
    else:
        d = os.path.join(golden_dir, which)
        fin = O.parse_fin(open(os.path.join(d, "out_fin_1_of_1")).read())
        test = O.parse_rat(open(os.path.join(d, "out_test_rat_1_of_1")).read())
    weights = O.weights_from_fin(fin)
    users, offsets, items, ratings = _csr(test)
    out = cpu_ref.local_calc(weights, offsets, items, ratings, n_threads=threads, honest=honest)
    pos = {(int(items[t]), u): t for ui, u in enumerate(users) for t in range(offsets[ui], offsets[ui + 1])}
    rows = O.local_calc(fin, test)
    assert out["pairs"] == len(rows)
    emitted = {(r[0], r[1]) for r in rows}
    for key, t in pos.items():
        assert (out["status"][t] == 4) == (key not in emitted)
    n_cmp = 0
    for (m, u, err, kk, pred, status, lim, w_lim, gap) in rows:
        t = pos[(m, u)]
        assert out["kk"][t] == kk
        if kk == 0:
            assert out["status"][t] == 1 and np.isnan(out["pred"][t])
            continue
        assert abs(out["w_lim"][t] - w_lim) <= 1e-10, (m, u, out["w_lim"][t], w_lim)
        assert out["lim"][t] == lim
        assert (out["status"][t] == 2) == (status == O.PRED_UNDERDETERMINED)
        if status == O.PRED_OK and gap > 1e-6:
            assert abs(out["pred"][t] - pred) <= 1e-7, (m, u, out["pred"][t], pred)
            assert abs(float(out["err"][t]) - float(err)) <= 1e-5 * max(1.0, float(err))
            n_cmp += 1
    assert n_cmp > 50


@pytest.mark.parametrize("n,kk", [(5, 1), (9, 4), (20, 7), (33, 32)])
def test_closed_form_cutoff_on_a_clique(n, kk):
    """Known answer that pins the cutoff independently of any solver: a local graph that is a clique with equal weights
    has L = a I - b J (a = n/(n-1), b = 1/(n-1)), so L_h L_h^T = a^2 I - n/(n-1)^2 J on the unrated nodes and
    w_lim = sqrt(n kk) / (n - 1) for kk rated neighbours.  The spectrum of L is {0, a (n-1 times)}: the max(.,2) rule cuts
    through the degenerate cluster, which the oracle reports as gap ~ 0."""
    from oracle import cpu_ref
    movies = list(range(1, n + 1))
    fin = [(a, b, 0.5) for a in movies for b in movies if a != b]
    user = O.UIMAX - 1
    test = {1: {user: 4.0}}
    for m in movies[1:1 + kk]:
        test[m] = {user: 3.0}
    rows = O.local_calc_movie(1, O.item_graph_weights(fin), test)
    assert len(rows) == 1
    (_, _, _, kk_o, _, status, lim, w_lim, gap) = rows[0]
    assert kk_o == kk and lim == 2 and abs(gap) <= 1e-9
    assert abs(w_lim - np.sqrt(n * kk) / (n - 1)) <= 1e-12
    # the C++ restatement (Householder + QL) meets the same closed form
    offsets = np.array([0, kk + 1], dtype=np.int64)
    items = np.array(movies[:kk + 1], dtype=np.int32)
    ratings = np.array([4.0] + [3.0] * kk)
    out = cpu_ref.local_calc(O.weights_from_fin(fin), offsets, items, ratings)
    assert abs(out["w_lim"][0] - np.sqrt(n * kk) / (n - 1)) <= 1e-10 and out["kk"][0] == kk


def test_rating_zero_means_unrated_and_masks_drop_pairs():
    """local_calc.cpp:406-413 tests `usr_rat(i) == 0`: a stored rating of 0 is indistinguishable from "not rated" -- in both
    restatements; a pair mask (the --pct sample, :266) removes pairs without changing the others; movie ids beyond the weight
    table have no vertex data and no line."""
    from oracle import cpu_ref
    fin, test = _random_case(7, n_items=30, n_users=12, density=0.5)
    u0 = sorted({u for d in test.values() for u in d})[0]
    zeroed = [m for m in sorted(test) if u0 in test[m]][1:3]
    for m in zeroed:
        test[m][u0] = 0.0                                            # two of this user's ratings read as "not rated"
    weights = O.weights_from_fin(fin)
    users, offsets, items, ratings = _csr(test)
    rows = O.local_calc(fin, test)
    out = cpu_ref.local_calc(weights, offsets, items, ratings, honest=False)
    pos = {(int(items[t]), u): t for ui, u in enumerate(users) for t in range(offsets[ui], offsets[ui + 1])}
    for (m, u, err, kk, pred, status, lim, w_lim, gap) in rows:
        assert out["kk"][pos[(m, u)]] == kk
    n_u0 = sum(1 for m in test if u0 in test[m])
    for (m, u, err, kk, *_rest) in rows:
        if u == u0 and m not in zeroed:
            assert kk <= n_u0 - 1 - len(zeroed)                      # the zeroed movies never count as known
    mask = (items % 2 == 1).astype(np.uint8)
    part = cpu_ref.local_calc(weights, offsets, items, ratings, pair_mask=mask, honest=False)
    assert (part["status"][mask == 0] == 4).all()
    sel = mask == 1
    for k in ("kk", "status", "lim"):
        assert np.array_equal(part[k][sel], out[k][sel])
    assert np.array_equal(part["w_lim"][sel], out["w_lim"][sel], equal_nan=True)
    items2 = items.copy()
    items2[0] = weights.shape[0] + 5                                 # outside the table: no vertex, no line
    out2 = cpu_ref.local_calc(weights, offsets, items2, ratings, honest=False)
    assert out2["status"][0] == 4
