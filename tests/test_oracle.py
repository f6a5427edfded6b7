"""CPU tests of the oracle itself (no GPU).  The reference has no golden vectors for this path
(SURVEY.md 4, 8c), so the oracle is pinned by closed forms and identities, by agreement between
its two independent restatements (numpy/LAPACK and the C++ tred2/tql2 restatement), and by the
committed golden fixtures (regeneration must reproduce them)."""
import os

import numpy as np
import pytest

from oracle import cpu_ref as C
from oracle import gsi_oracle as O


def _rand_w(rng, n_items, density=0.6):
    w = np.round(1.0 - 0.5 * rng.random((n_items + 1, n_items + 1)), 6)
    w = np.where(rng.random(w.shape) < density, w, 0.0)
    w = np.triu(w, 1)
    w = w + w.T
    w[0] = 0
    w[:, 0] = 0
    return w


def test_complete_graph_closed_form():
    # K_n with unit weights: lambda = {0, n/(n-1) x (n-1)}; sig_min_i = sqrt(1 + 1/(n-1)) + 0.01
    for n in (2, 3, 7, 20):
        w = np.ones((n + 1, n + 1)) - np.eye(n + 1)
        w[0] = 0
        w[:, 0] = 0
        rec = O.precompute_user(5, np.arange(1, n + 1), w, keep_ll2=True)
        lam, _ = O.eig_lower(rec.ll2)
        assert abs(lam[0]) < 1e-12
        assert np.allclose(lam[1:], n / (n - 1.0), atol=1e-12)
        expect = np.float32(np.sqrt(np.float32(1 + 1.0 / (n - 1)))) + 0.01
        assert np.allclose(rec.sigs_min, expect, rtol=1e-6)


def test_isolated_items_and_out_of_table_ids():
    w = np.zeros((11, 11))
    rec = O.precompute_user(1, np.array([2, 5, 40]), w, keep_ll2=True)   # id 40 >= rows -> 0
    assert np.array_equal(rec.ll2, np.eye(3))                            # degree 0 -> 1
    assert np.allclose(rec.sigs_min, 1.01)
    assert len(rec.lam) == 3 and np.allclose(rec.lam, 1.0)              # all lam=1 <= 1.01 kept


def test_n1_defined_behaviour():
    rec = O.precompute_user(1, np.array([3]), np.zeros((5, 5)))
    assert len(rec.lam) == 2 and rec.vec.shape == (1, 2)
    assert rec.lam[0] == 1.0 and rec.lam[1] == 0.0 and rec.vec[0, 1] == 0.0


def test_sig_min_closed_form_and_identities():
    rng = np.random.default_rng(1)
    w = _rand_w(rng, 80)
    items = np.sort(rng.choice(np.arange(1, 81), 37, replace=False))
    rec = O.precompute_user(9, items, w, keep_ll2=True)
    ww = O.gather_ww(items, w)
    d = ww.sum(1)
    d[d == 0] = 1
    closed = np.sqrt(1 + ((ww ** 2) / np.outer(d, d)).sum(1)) + 0.01
    assert np.allclose(rec.sigs_min, closed, rtol=1e-6)
    lam, u = O.eig_lower(rec.ll2)
    sym = np.tril(rec.ll2) + np.tril(rec.ll2, -1).T
    assert np.abs(sym @ u - u * lam).max() < 1e-12
    assert np.abs(u.T @ u - np.eye(len(items))).max() < 1e-12
    assert lam.min() > -1e-12 and lam.max() < 2 + 1e-12
    assert abs(lam[0]) < 1e-12 and np.allclose(np.abs(u[:, 0]), np.sqrt(d / d.sum()), atol=1e-10)
    k = len(rec.lam)
    assert k >= 2 and (k == len(lam) or lam[k] > rec.sigs_min.max() - 1e-6)


def test_cpp_restatement_agrees_with_numpy_oracle():
    rng = np.random.default_rng(2)
    w = _rand_w(rng, 150, 0.5)
    sizes = [1, 2, 3, 9, 33, 64, 100]
    lists = [np.sort(rng.choice(np.arange(1, 170), n, replace=False)).astype(np.int32) for n in sizes]
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    items = np.concatenate(lists)
    for honest in (True, False):
        out = C.precompute(w, offsets, items, n_threads=3, honest=honest, format_text=True)
        for u, it in enumerate(lists):
            rec = O.precompute_user(u, it, w)
            assert np.array_equal(rec.sigs_min, out["sig_min"][offsets[u]:offsets[u + 1]])
            assert out["k"][u] == len(rec.lam)
            assert np.abs(rec.lam - out["lam"][u]).max() < 1e-11
            p1, p2 = rec.vec @ rec.vec.T, out["vec"][u] @ out["vec"][u].T
            assert np.abs(p1 - p2).max() < 1e-9
        assert out["text_bytes"] > 0


def test_out_eigen_roundtrip_and_bug_b1(golden_dir):
    path = os.path.join(golden_dir, "tiny_int", "out_eigen_")
    z = np.load(os.path.join(golden_dir, "tiny_int", "oracle.npz"))
    on = O.parse_out_eigen(path, bug_b1=True)
    off = O.parse_out_eigen(path, bug_b1=False)
    users = list(z["user"])
    assert sorted(on) == sorted(users)
    first = users[0]
    n0 = len(on[first]["items"])
    assert np.array_equal(on[first]["sigs_min"], off[first]["sigs_min"])
    second = users[1]
    # B1: the second record's sigs_min starts with the first record's values
    assert np.array_equal(on[second]["sigs_min"][:n0], on[first]["sigs_min"])
    assert len(off[second]["sigs_min"]) == len(off[second]["items"])
    # text carries 6 significant digits
    o = z["offsets"]
    assert np.allclose(off[first]["sigs_min"], z["sig_min"][o[0]:o[1]], rtol=1e-5)
    text = open(path).read()
    assert text.count("\n") == 3 * len(users) and " \n" in text


def test_format_g_matches_cxx_default_stream():
    assert O.fmt_g(1.01) == "1.01" and O.fmt_g(0.70710678) == "0.707107"
    assert O.fmt_g(-1.11022e-16) == "-1.11022e-16" and O.fmt_g(3.0) == "3"
    assert O.fmt_g(np.float32(0.1)) == "0.1" and O.fmt_g(float("nan")) == "nan"
    assert O.edge_kept(0.1) and not O.edge_kept(0.0999999)       # B8


def test_knn_chain_small():
    train = [(1, 1, 5), (1, 2, 3), (2, 1, 4), (2, 2, 2), (2, 3, 1), (3, 3, 5)]
    val = [(9, 2, 4), (9, 4, 2)]
    rat, test_rat, edg = O.knn1(train, val)
    assert edg[1] == [2, 3] and edg[2] == [1, 3, 4] and edg[4] == [2] and edg[3] == [1, 2]
    assert test_rat[2] == {O.UIMAX - 9: 4} and rat[4] == {} and test_rat[1] == {}
    text = O.format_rat(test_rat)
    assert "1 \n" in text and ("2 %d 4 \n" % (O.UIMAX - 9)) in text
    # cosine needs > 5 common raters
    ra = {u: float(1 + u % 5) for u in range(10)}
    rb = {u: float(1 + (u * 3) % 5) for u in range(4, 14)}
    w, cnt = O.knn2_weight(ra, rb)
    assert cnt == 6
    common = [u for u in ra if u in rb]
    a = np.array([ra[u] for u in common])
    b = np.array([rb[u] for u in common])
    assert abs(w - (a @ b) / np.sqrt((a @ a) * (b @ b))) < 1e-6
    assert O.knn2_weight({u: 1.0 for u in range(5)}, {u: 1.0 for u in range(5)}) == (0.0, 5)


def test_golden_is_reproducible(golden_dir, tmp_path):
    """Re-running the oracle on the committed inputs reproduces the committed outputs."""
    d = os.path.join(golden_dir, "tiny_int")
    train, val = O.read_rating_files(os.path.join(d, "movielens"))
    rat, test_rat, edg = O.knn1(train, val)
    assert O.format_edg(edg) == open(os.path.join(d, "out_edg_1_of_1")).read()
    assert O.format_rat(rat) == open(os.path.join(d, "out_rat_1_of_1")).read()
    fin_text = O.format_fin(O.knn2(rat, edg))
    assert fin_text == open(os.path.join(d, "out_fin_1_of_1")).read()
    fin = O.parse_fin(fin_text)
    recs = O.precompute_all(O.users_from_validate(val), O.weights_from_fin(fin))
    z = np.load(os.path.join(d, "oracle.npz"))
    assert np.array_equal(np.array([len(r.lam) for r in recs]), z["k"])
    assert np.array_equal(np.concatenate([r.sigs_min for r in recs]), z["sig_min"])
    assert np.allclose(np.concatenate([r.lam for r in recs]), z["lam"], atol=1e-12)
    ud = O.parse_out_eigen(os.path.join(d, "out_eigen_"), bug_b1=False)
    rows = O.local_calc_precomp(ud, O.item_graph(fin), O.parse_rat(open(os.path.join(d, "out_test_rat_1_of_1")).read()))
    assert O.format_res(rows) == open(os.path.join(d, "out_res_b1_off")).read()


def test_predictor_permutation_invariance_and_sign_protocol(golden_dir):
    """(d)/(e) of SURVEY.md 7: with B1 off the prediction does not depend on row order inside a
    record; flipping an eigenvector's sign can change the signed column-drop (H1)."""
    d = os.path.join(golden_dir, "tiny_int")
    ud = O.parse_out_eigen(os.path.join(d, "out_eigen_"), bug_b1=False)
    fin = O.parse_fin(open(os.path.join(d, "out_fin_1_of_1")).read())
    graph = O.item_graph(fin)
    test_rat = O.parse_rat(open(os.path.join(d, "out_test_rat_1_of_1")).read())
    user = sorted(ud)[0]
    rec = ud[user]
    ur = {m: test_rat[m][user] for m in rec["items"]}
    m = int(rec["items"][3])
    base = O.predict_pair(rec, m, graph.get(m, set()), ur, ur[m])
    perm = np.random.default_rng(0).permutation(len(rec["items"]))
    rec2 = dict(rec, items=rec["items"][perm], vec=rec["vec"][perm], sigs_min=rec["sigs_min"][perm],
                row_of={int(x): i for i, x in enumerate(rec["items"][perm])})
    alt = O.predict_pair(rec2, m, graph.get(m, set()), ur, ur[m])
    assert base[1] == alt[1] and base[3] == alt[3]
    if base[3] == O.PRED_OK:
        assert abs(base[2] - alt[2]) < 1e-8
    flipped = dict(rec, vec=-rec["vec"])
    f = O.predict_pair(flipped, m, graph.get(m, set()), ur, ur[m], coldrop_signed=False)
    g = O.predict_pair(rec, m, graph.get(m, set()), ur, ur[m], coldrop_signed=False)
    if g[3] == O.PRED_OK:
        assert abs(f[2] - g[2]) < 1e-8                     # |.| drop is sign invariant


def test_fold_split_semantics():
    ratings = [(u, m, 3) for u in range(1, 24) for m in range(1, 4)]
    folds = O.fold_split(ratings, 5, np.random.default_rng(0))
    test_users = [sorted({t[0] for t in te}) for _, te in folds]
    assert sum(len(t) for t in test_users) == 23
    assert all(len(t) == 5 for t in test_users[:4])        # cut when count > 23/5 = 4.6
    for tr, te in folds:
        assert not ({t[0] for t in tr} & {t[0] for t in te})
