"""CPU oracle for the per-user graph-signal interpolation path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference's algorithm for the hot path
(precompute_local / precompute_local_threads / local_calc_precomp / knn / knn2 / knn3).
It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  The product
(``collaborative_filtering_b200``) never imports anything under ``oracle/``.

PARITY UNPINNED.  The reference ships no golden vectors, known-answer tests or fixtures for
this path and it cannot be compiled here: its arithmetic lives in
Eigen 3.x (``SelfAdjointEigenSolver``, ``MatrixXd::inverse``), Boost (``unordered_map``
iteration order, ``threadpool``) and GraphLab PowerGraph v2.x, none of which is vendored in
/root/reference or installed in this image (versions are unpinned in the reference as well:
``CMakeLists.txt:8`` ``requires_eigen``).  This oracle therefore restates the published
algorithms of those calls (Householder tridiagonalisation + implicit QR == LAPACK ``syevd``
semantics via ``numpy.linalg.eigh`` on the lower triangle, ascending, unit-norm vectors) and
anchors parity on the reference's own call sites, cited per function as file:line into
/root/reference.  What it is pinned against instead: closed-form spectra and identities in
``tests/test_oracle.py`` (K_n, isolated items, sig_min closed form, U^T U = I, L U = U diag(lam)).

Defined behaviour where the reference's is unreproducible:
  B5  n == 1          -> k = 2, lambda_2 = 0, U_2 = 0 (reference grows arrays with garbage)
  B6  iteration order -> ascending user', ascending movie id (reference: boost hash order)
  H1  eigenvector sign -> the component of largest magnitude is positive (first on ties)
  B3  singular Gram    -> classified (``status`` in ``predict_pair``), never compared
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

UIMAX = 2147483647  # std::numeric_limits<int>::max(), precompute_local.cpp:19, knn.cpp:19

# --------------------------------------------------------------------------------------
# text formatting: C++ default ostream formatting == printf("%g") with precision 6
# (precompute_local.cpp:263-280, local_calc_precomp.cpp:397-399, knn.cpp:306-311)
# --------------------------------------------------------------------------------------


def fmt_g(x) -> str:
    """operator<<(ostream&, double) with default flags (precision 6, general format)."""
    return "%g" % float(x)


def round6(x):
    """Value a 6-significant-digit text round trip leaves: strtod(printf("%g", x))."""
    x = np.asarray(x, dtype=np.float64)
    out = np.array([float("%g" % v) for v in x.ravel()], dtype=np.float64)
    return out.reshape(x.shape)


# --------------------------------------------------------------------------------------
# P1/K1 ingest: GraphLab-style "user item rating" files  (collaborative_filtering.dox:62-80)
# --------------------------------------------------------------------------------------


def read_rating_files(directory: str):
    """Returns (train, validate): lists of (user, item, rating) from every regular file of
    ``directory`` in sorted name order.  Role by suffix: ``*.validate`` is VALIDATE, everything
    else TRAIN (knn.cpp:88-92).  Empty lines are skipped (precompute_local.cpp:103-104)."""
    train, validate = [], []
    for name in sorted(os.listdir(directory)):
        path = os.path.join(directory, name)
        if not os.path.isfile(path):
            continue
        dst = validate if name.endswith(".validate") else train
        with open(path) as f:
            for line in f:
                tok = line.split()
                if len(tok) < 3:
                    continue
                dst.append((int(tok[0]), int(tok[1]), float(tok[2])))
    return train, validate


def users_from_validate(validate):
    """P1 (precompute_local.cpp:94-112): user' = INT_MAX - user; per user the SET of rated movie
    ids (duplicates collapse, :108).  Order defined as ascending (B6).  Returns
    {user': sorted np.int64 array of movie ids}."""
    users: dict[int, set] = {}
    for u, m, _ in validate:
        users.setdefault(UIMAX - u, set()).add(m)
    return {u: np.array(sorted(s), dtype=np.int64) for u, s in sorted(users.items())}


def weights_from_fin(fin_edges):
    """P2 (precompute_local.cpp:113-158): dense directed table weights(m1,m2)=w, last wins,
    size (N+1)^2 with N = max id seen; row/col 0 unused.  ``fin_edges`` = iterable of
    (m1, m2, w) where w is the double parsed from text.  B4: the 2000 cap is lifted."""
    fin_edges = list(fin_edges)
    n = 0
    for a, b, _ in fin_edges:
        n = max(n, a, b)
    w = np.zeros((n + 1, n + 1), dtype=np.float64)
    for a, b, v in fin_edges:
        w[a, b] = v
    return w


# --------------------------------------------------------------------------------------
# P3-P8 precompute for one user  (precompute_local.cpp:185-261 == precompute_local_threads.cpp:118-194)
# --------------------------------------------------------------------------------------


def canonical_sign(u: np.ndarray) -> np.ndarray:
    """H1 convention: in every column the entry of largest |value| is made positive."""
    u = np.array(u, dtype=np.float64, copy=True)
    if u.size == 0:
        return u
    idx = np.argmax(np.abs(u), axis=0)
    sgn = np.sign(u[idx, np.arange(u.shape[1])])
    sgn[sgn == 0] = 1.0
    return u * sgn


def gather_ww(items: np.ndarray, weights: np.ndarray) -> np.ndarray:
    """P3 (precompute_local.cpp:185-192): ww(i,j)=weights(m_i,m_j), 0 when max(m_i,m_j) >= rows."""
    items = np.asarray(items, dtype=np.int64)
    n = len(items)
    rows = weights.shape[0]
    ok = items < rows
    ww = np.zeros((n, n), dtype=np.float64)
    idx = np.where(ok)[0]
    if len(idx):
        ww[np.ix_(idx, idx)] = weights[np.ix_(items[idx], items[idx])]
    return ww


def normalized_laplacian(ww: np.ndarray):
    """P4+P5 (precompute_local.cpp:196-222).  d_i = sum_j ww(i,j) accumulated j ascending in
    double, 1 when the sum == 0; s_i = sqrt(1/d_i) (dd.inverse() of a diagonal matrix then
    elementwise sqrt, :216-220 -- this is NOT 1/sqrt(d_i) bit for bit);
    ll2_ij = fl(fl(s_i * ll_ij) * s_j) with ll = dd - ww (:211-212, :222)."""
    n = ww.shape[0]
    d = np.zeros(n, dtype=np.float64)
    for j in range(n):  # sequential j order, as the reference's scalar loop :199-203
        d = d + ww[:, j]
    d = np.where(d == 0.0, 1.0, d)
    s = np.sqrt(1.0 / d)
    ll = np.diag(d) - ww
    ll2 = (s[:, None] * ll) * s[None, :]
    return d, s, ll2


def sig_min_rows(ll2: np.ndarray):
    """P7 (precompute_local.cpp:236-249): float accumulator ``sig_min += pow(ll2(i,j),2)`` ==
    (float)((double)acc + x*x), j ascending over the FULL row; sqrt in float; stored value is
    (double)sig_min + 0.01; sig_min_max (float) = max_i sig_min, then += 0.01 in float."""
    n = ll2.shape[0]
    acc = np.zeros(n, dtype=np.float32)
    for j in range(n):
        acc = (acc.astype(np.float64) + ll2[:, j] * ll2[:, j]).astype(np.float32)
    sig = np.sqrt(acc)  # float32 sqrt, correctly rounded
    sigs_min = sig.astype(np.float64) + 0.01
    sig_max = np.float32(0.0)
    for v in sig:  # "if (sig_min_max < sig_min)" :246
        if sig_max < v:
            sig_max = v
    sig_min_max = np.float32(np.float64(sig_max) + 0.01)
    return sigs_min, sig_min_max


def eig_lower(ll2: np.ndarray):
    """P6 (precompute_local.cpp:231-233): SelfAdjointEigenSolver reads the lower triangle only,
    eigenvalues ascending, unit-norm eigenvectors (columns).  Sign fixed by H1 convention."""
    if ll2.shape[0] == 0:
        return np.zeros(0), np.zeros((0, 0))
    lam, u = np.linalg.eigh(ll2, UPLO="L")
    return lam, canonical_sign(u)


def cutoff(lam: np.ndarray, thr) -> int:
    """P8 / R5 (precompute_local.cpp:252-258, local_calc_precomp.cpp:272-279):
    first index with lam > thr (or len), then max(.,2)."""
    thr = float(thr)
    lim = len(lam)
    for i, v in enumerate(lam):
        if v > thr:
            lim = i
            break
    return max(lim, 2)


@dataclass
class UserRecord:
    """One out_eigen_ record (README.md:14-19; precompute_local.cpp:263-280)."""

    user: int                       # user' = INT_MAX - user id
    items: np.ndarray               # n movie ids, ascending (B6)
    sigs_min: np.ndarray            # n doubles
    lam: np.ndarray                 # k eigenvalues
    vec: np.ndarray                 # n x k eigenvectors (row i = movie i)
    ll2: np.ndarray | None = field(default=None, repr=False)


def precompute_user(user: int, items, weights: np.ndarray, keep_ll2: bool = False) -> UserRecord:
    """compute_eigens(user_id, ratings) (precompute_local_threads.cpp:100-213)."""
    items = np.asarray(items, dtype=np.int64)
    n = len(items)
    ww = gather_ww(items, weights)
    _, _, ll2 = normalized_laplacian(ww)
    lam, u = eig_lower(ll2)
    sigs_min, sig_min_max = sig_min_rows(ll2)
    lim = cutoff(lam, sig_min_max)
    if lim > n:  # B5: defined behaviour for n < 2 (reference reads uninitialised memory)
        lam_k = np.zeros(lim)
        lam_k[:n] = lam
        u_k = np.zeros((n, lim))
        u_k[:, :n] = u
    else:
        lam_k, u_k = lam[:lim].copy(), u[:, :lim].copy()
    return UserRecord(user, items, sigs_min, lam_k, u_k, ll2 if keep_ll2 else None)


def precompute_all(users: dict, weights: np.ndarray):
    """Hot loop over users (precompute_local.cpp:165-282), ascending user' (B6)."""
    return [precompute_user(u, items, weights) for u, items in sorted(users.items())]


# --------------------------------------------------------------------------------------
# P9 / R1: out_eigen_ text format
# --------------------------------------------------------------------------------------


def format_record(rec: UserRecord) -> str:
    """precompute_local.cpp:265-278: three lines, every token followed by one space."""
    n, k = len(rec.items), len(rec.lam)
    a = [str(int(rec.user)), str(n), str(k)]
    for m, s in zip(rec.items, rec.sigs_min):
        a.append(str(int(m)))
        a.append(fmt_g(s))
    l1 = " ".join(a) + " \n"
    l2 = "".join(fmt_g(v) + " " for v in rec.lam) + "\n"
    l3 = "".join(fmt_g(v) + " " for v in rec.vec.reshape(-1)) + "\n"
    return l1 + l2 + l3


def write_out_eigen(path: str, records) -> None:
    with open(path, "w") as f:
        for r in records:
            f.write(format_record(r))


def parse_out_eigen(path_or_text, bug_b1: bool = True, is_text: bool = False):
    """load_precomputed_data (local_calc_precomp.cpp:406-482).  Returns {user': dict(items,
    row_of (movie->row), sigs_min, lam, vec)} .  bug_b1=True reproduces :414,437,440 -- the
    local ``sigs_min`` vector is never cleared, so every record carries the concatenation of
    all sig_min values read so far (and ``sigs_min[movie_ind]`` at :271 indexes that)."""
    text = path_or_text if is_text else open(path_or_text).read()
    lines = [ln for ln in text.split("\n") if len(ln)]
    users = {}
    running: list[float] = []
    i = 0
    while i + 3 <= len(lines):  # 3-state machine :425-477 (empty lines already skipped :421)
        t = lines[i].split()
        user, n, k = int(t[0]), int(t[1]), int(t[2])
        assert len(t) >= 3 + 2 * n, "assert(!parseline.fail()) :434"
        items = np.array([int(t[3 + 2 * j]) for j in range(n)], dtype=np.int64)
        sig = [float(t[4 + 2 * j]) for j in range(n)]
        if not bug_b1:
            running = []
        running.extend(sig)
        lam = np.array([float(x) for x in lines[i + 1].split()], dtype=np.float64)
        assert len(lam) >= k, "assert(!parseline.fail()) :447"
        v = np.array([float(x) for x in lines[i + 2].split()], dtype=np.float64)
        assert len(v) >= n * k, "assert(!parseline.fail()) :462"
        users[user] = dict(
            items=items,
            row_of={int(m): j for j, m in enumerate(items)},  # movie_list[movie_id] = i  :436
            sigs_min=np.array(running, dtype=np.float64),
            lam=lam[:k],
            vec=v[: n * k].reshape(n, k),
        )
        i += 3
    return users


# --------------------------------------------------------------------------------------
# R2/R3 loaders + R5-R8 predictor  (local_calc_precomp.cpp:122-160, 217-380)
# --------------------------------------------------------------------------------------


def edge_kept(w) -> bool:
    """R2 / knn3 loader (local_calc_precomp.cpp:129-133, knn3.cpp:88-92): the weight is parsed
    into a ``float`` and compared with the double literal 0.1, so "0.1" passes (B8)."""
    return float(np.float32(w)) > 0.1


def item_graph(fin_edges):
    """{movie: set of out-neighbours} over edges with (float)w > 0.1."""
    g: dict[int, set] = {}
    for a, b, w in fin_edges:
        if edge_kept(w):
            g.setdefault(int(a), set()).add(int(b))
    return g


PRED_OK, PRED_EMPTY, PRED_UNDERDETERMINED, PRED_SINGULAR = 0, 1, 2, 3


def predict_pair(ud: dict, movie: int, neigh: set, user_ratings: dict, rat_real: float,
                 coldrop_signed: bool = True):
    """One (movie, test-user) prediction, local_calc_precomp.cpp:230-360.

    ud           parsed record of the user (parse_out_eigen)
    neigh        out-neighbours of ``movie`` in the thresholded item graph (R4, :206-215)
    user_ratings {movie j: rating of this user for j} from out_test_rat_ (float -> double, R3)
    Returns (err float32, kk, pred double, status, c).  K is taken in record order (B6)."""
    uu = ud["vec"]
    lam = ud["lam"]
    movie_ind = ud["row_of"][movie]                      # :248
    vv = uu[movie_ind, :].copy()                         # :249-250
    rows, usr_rat = [], []
    for j, row in zip(ud["items"], range(len(ud["items"]))):   # :254-265
        if int(j) in neigh:
            usr_rat.append(float(user_ratings[int(j)]))
            rows.append(row)
    usr_rat = np.array(usr_rat, dtype=np.float64)
    uu_hh = uu[rows, :] if len(rows) else np.zeros((0, uu.shape[1]))
    kk = len(rows)
    w_lim = ud["sigs_min"][movie_ind]                    # :271  (B1 lives in ud["sigs_min"])
    lim = min(cutoff(lam, w_lim), uu.shape[1])           # :272-279 (k >= 2 always, so lim <= k)
    vv = vv[:lim]
    uu_hh = uu_hh[:, :lim]
    # column clean :284-304 -- keep column i iff some entry >= 1e-4 (signed test, B2)
    if kk:
        test = uu_hh if coldrop_signed else np.abs(uu_hh)
        keep = (test >= 0.0001).any(axis=0)
    else:
        keep = np.zeros(uu_hh.shape[1], dtype=bool)
    uu_hh = uu_hh[:, keep]
    vv = vv[keep]
    c = int(keep.sum())
    status = PRED_OK
    if kk == 0:
        status = PRED_EMPTY                              # 0/0 -> NaN  (:311)
        pred = float("nan")
    else:
        mm = uu_hh.T @ uu_hh                             # :309
        rat_mean = usr_rat.sum() / kk                    # :311
        rhs = uu_hh.T @ (usr_rat - rat_mean)
        if kk < c:
            status = PRED_UNDERDETERMINED                # rank-deficient Gram, B3/H2
        try:
            x = np.linalg.solve(mm, rhs) if c else np.zeros(0)
            if c and np.linalg.cond(mm) > 1e8 and status == PRED_OK:
                status = PRED_SINGULAR
        except np.linalg.LinAlgError:
            x = np.full(c, np.nan)
            status = PRED_SINGULAR if status == PRED_OK else status
        pred = float(vv @ x) + rat_mean if c else rat_mean   # :314-315
    p = pred
    if p > 5:                                            # :322-325
        p = 5.0
    if p < 1:
        p = 1.0
    err = (rat_real - p) ** 2                            # :327
    return np.float32(err), kk, pred, status, c


def local_calc_precomp(user_data: dict, graph: dict, test_rat: dict, coldrop_signed=True):
    """apply() over every movie vertex that has test ratings, --pct 100
    (local_calc_precomp.cpp:217-380, writer :393-404).  ``test_rat`` = {movie: {user': rating}}
    (R3).  Returns rows (movie, user', err float32, kk, pred, status, c) in ascending
    movie, ascending user' order (B6)."""
    by_user: dict[int, dict] = {}
    for m, d in test_rat.items():
        for u, r in d.items():
            by_user.setdefault(u, {})[m] = r
    out = []
    for m in sorted(test_rat):
        neigh = graph.get(m, set())
        for u in sorted(test_rat[m]):
            if u not in user_data:
                continue
            err, kk, pred, status, c = predict_pair(
                user_data[u], m, neigh, by_user[u], float(test_rat[m][u]), coldrop_signed)
            out.append((m, u, err, kk, pred, status, c))
    return out


# --------------------------------------------------------------------------------------
# local_calc.cpp: the per-MOVIE variant.  One local graph per movie (the movie
# and its out-neighbours), one eigensolve per movie, and per (movie, test user) pair the exact
# cutoff sigma_min(L_h) from one more eigensolve.
# --------------------------------------------------------------------------------------
def item_graph_weights(fin_edges):
    """graph_loader (local_calc.cpp:102-117): edge a -> b kept iff (float)w > 0.1; the edge carries
    ``edge_data(weight)``, i.e. the FLOAT-rounded weight widened to double.  {a: {b: w}}."""
    g: dict[int, dict] = {}
    for a, b, w in fin_edges:
        if edge_kept(w):
            g.setdefault(int(a), {})[int(b)] = float(np.float32(w))
    return g


def local_graph(m: int, gw: dict):
    """Adjacency of the local graph of movie ``m`` (local_calc.cpp:265-335).  Node 0 is m, nodes
    1.. are its out-neighbours (hash order in the reference; ascending id here, B6).
    ww(i, j) = w(i -> j) between neighbours (:326-330); weights to movies outside the local graph
    land on column 0 through ``indices[]``'s default value and are overwritten (:331-333) by the
    fix-up ww(0, i) = ww(i, 0) = w(m -> i).  Self edges are not part of the item graph (knn2 never
    emits one) and are ignored."""
    neigh = sorted(j for j in gw.get(m, {}) if j != m)
    nodes = [m] + neigh
    idx = {v: i for i, v in enumerate(nodes)}
    n = len(nodes)
    ww = np.zeros((n, n), dtype=np.float64)
    for i in neigh:
        for j2, w in gw.get(i, {}).items():
            if j2 in idx and j2 != i and j2 != m:
                ww[idx[i], idx[j2]] = w
        ww[0, idx[i]] = gw[m][i]
        ww[idx[i], 0] = gw[m][i]
    return nodes, ww


def local_calc_movie(m: int, gw: dict, test_rat: dict):
    """vertex_program::apply for one movie, --pct 100 (local_calc.cpp:262-526).  Returns rows
    (movie, user', err float32, kk, pred, status, lim, w_lim, gap) in ascending user' order, or []
    when the local graph has fewer than 3 nodes (:271-272) or the movie has no test ratings.
    ``gap`` = lambda[lim] - lambda[lim-1] (inf when every eigenpair is used): when the cutoff -- usually
    the max(.,2) rule -- falls inside a cluster of equal eigenvalues, U[:, :lim] depends on the
    eigensolver's arbitrary basis of that cluster and so does the reference's own prediction; parity
    tests compare predictions only where gap > 1e-6 (eigenvectors up to subspace rotation)."""
    nodes, ww = local_graph(m, gw)
    n = len(nodes)
    users = sorted(test_rat.get(m, {}))
    if n < 3 or not users:
        return []
    _, _, ll2 = normalized_laplacian(ww)                 # :347-374 (no zero degree: every row holds w(m -> i) > 0.1)
    lam, uu = eig_lower(ll2)                             # :378
    out = []
    for u in users:
        rat = np.array([float(test_rat.get(v, {}).get(u, 0.0)) for v in nodes])   # :305-321
        rat_real = rat[0]
        rat[0] = 0.0                                     # :405
        unrated = rat == 0.0                             # :406-413
        rated = ~unrated
        kk = int(rated.sum())
        ll2_h = ll2[unrated, :]                          # :420-431
        with np.errstate(invalid="ignore"):
            w_lim = float(np.sqrt(np.linalg.eigvalsh(ll2_h @ ll2_h.T).min()))     # :435-436
        lim = cutoff(lam, w_lim)                         # :443-451 (NaN never compares greater: lim = n)
        vv = uu[0, :lim]                                 # :456-478
        uu_hh = uu[rated, :lim]
        usr_rat = rat[rated]
        status = PRED_OK
        if kk == 0:
            status, pred = PRED_EMPTY, float("nan")      # 0/0 :487
        else:
            mm = uu_hh.T @ uu_hh                         # :484-485
            rat_mean = usr_rat.sum() / kk
            rhs = uu_hh.T @ (usr_rat - rat_mean)
            if kk < lim:
                status = PRED_UNDERDETERMINED
            try:
                x = np.linalg.solve(mm, rhs)
                if status == PRED_OK and np.linalg.cond(mm) > 1e8:
                    status = PRED_SINGULAR
            except np.linalg.LinAlgError:
                x = np.full(lim, np.nan)
                status = PRED_SINGULAR if status == PRED_OK else status
            pred = float(vv @ x) + rat_mean              # :490-491
        p = pred
        if p > 5:                                        # :494-497
            p = 5.0
        if p < 1:
            p = 1.0
        err = (rat_real - p) ** 2                        # :499
        gap = float(lam[lim] - lam[lim - 1]) if lim < n else float("inf")
        out.append((m, u, np.float32(err), kk, pred, status, lim, w_lim, gap))
    return out


def local_calc(fin_edges, test_rat: dict):
    """Both engines of local_calc.cpp over every vertex (:614-648), ascending movie id."""
    gw = item_graph_weights(fin_edges)
    out = []
    for m in sorted(test_rat):
        out.extend(local_calc_movie(m, gw, test_rat))
    return out


def format_res(rows) -> str:
    """graph_writer::save_vertex (local_calc_precomp.cpp:393-404): "movie user' mse kk\\n"."""
    return "".join("%d %d %s %d\n" % (m, u, fmt_g(e), kk) for m, u, e, kk, *_ in rows)


def rmse_of(rows, statuses=(PRED_OK,)):
    """RMSE over out_res rows (not computed by any shipped reference code; the user averages
    column 3 offline, run_test_precompute.sh:19).  NaN rows are excluded and counted."""
    errs = [float(r[2]) for r in rows if r[5] in statuses and not np.isnan(r[2])]
    return (float(np.sqrt(np.mean(errs))) if errs else float("nan")), len(errs)


# --------------------------------------------------------------------------------------
# K1: knn  (knn.cpp:83-111, 160-357)
# --------------------------------------------------------------------------------------


def knn1(train, validate):
    """Returns (rat, test_rat, edg):
    rat[m]      = {user': rating} over TRAIN edges          (out_rat_*,      knn.cpp:303-315)
    test_rat[m] = {user': rating} over VALIDATE edges        (out_test_rat_*, knn.cpp:320-332)
    edg[m]      = ascending unique co-rated movie ids != m, co-rating through train AND
                  validate edges (engine2/engine3, knn.cpp:218-281; writer :337-357).
    A key exists for every movie vertex, even with an empty map (a line "m \\n" is written)."""
    rat, test_rat, movies_of = {}, {}, {}
    for (u, m, r) in train:
        rat.setdefault(m, {})[UIMAX - u] = r
        test_rat.setdefault(m, {})
        movies_of.setdefault(UIMAX - u, set()).add(m)
    for (u, m, r) in validate:
        test_rat.setdefault(m, {})[UIMAX - u] = r
        rat.setdefault(m, {})
        movies_of.setdefault(UIMAX - u, set()).add(m)
    raters: dict[int, set] = {}
    for u, ms in movies_of.items():
        for m in ms:
            raters.setdefault(m, set()).add(u)
    edg = {}
    for m in rat:
        s = set()
        for u in raters.get(m, ()):
            s |= movies_of[u]
        s.discard(m)
        edg[m] = sorted(s)
    return rat, test_rat, edg


def format_rat(rat: dict) -> str:
    out = []
    for m in sorted(rat):
        out.append(str(m) + " " + "".join("%d %s " % (u, fmt_g(r)) for u, r in sorted(rat[m].items())) + "\n")
    return "".join(out)


def format_edg(edg: dict) -> str:
    return "".join(str(m) + " " + "".join("%d " % j for j in edg[m]) + "\n" for m in sorted(edg))


def parse_rat(text: str) -> dict:
    """graph_vertex_loader / graph_test_loader (knn2.cpp:79-102, local_calc_precomp.cpp:138-160)."""
    out = {}
    for line in text.split("\n"):
        t = line.split()
        if not t:
            continue
        d = {}
        for i in range(1, len(t) - 1, 2):
            d[int(t[i])] = float(t[i + 1])
        out[int(t[0])] = d
    return out


def parse_fin(text: str):
    out = []
    for line in text.split("\n"):
        t = line.split()
        if len(t) >= 3:
            out.append((int(t[0]), int(t[1]), float(t[2])))
    return out


# --------------------------------------------------------------------------------------
# K2: knn2 cosine weights  (knn2.cpp:127-164)
# --------------------------------------------------------------------------------------


def knn2_weight(ra: dict, rb: dict):
    """weights_calc (knn2.cpp:127-146): float accumulators num, den1, den2 updated as
    (float)((double)acc + a*b); common raters in ascending user' order (B6; exact for integer
    and half-star ratings, whose partial sums are exactly representable in float);
    w = num / (sqrtf(den1) * sqrtf(den2)) in float when cnt > 5, else 0."""
    num = np.float32(0)
    den1 = np.float32(0)
    den2 = np.float32(0)
    cnt = 0
    small, big, swap = (ra, rb, False) if len(ra) <= len(rb) else (rb, ra, True)
    for u in sorted(small):
        if u in big:
            a, b = (small[u], big[u]) if not swap else (big[u], small[u])
            cnt += 1
            num = np.float32(np.float64(num) + a * b)
            den1 = np.float32(np.float64(den1) + a * a)
            den2 = np.float32(np.float64(den2) + b * b)
    if cnt > 5:
        w = np.float32(num / np.float32(np.sqrt(den1) * np.sqrt(den2)))
        return float(w), cnt
    return 0.0, cnt


def knn2(rat: dict, edg: dict):
    """All directed edges a->b of out_edg with their cosine weight; emitted iff w > 0.01
    (knn2.cpp:155-163).  Returns list of (a, b, w_double) in ascending (a, b) order; w is the
    in-memory double (a float value).  The text written is fmt_g(w)."""
    out = []
    for a in sorted(edg):
        ra = rat.get(a, {})
        for b in edg[a]:
            w, _ = knn2_weight(ra, rat.get(b, {}))
            if w > 0.01:
                out.append((a, b, w))
    return out


def format_fin(edges) -> str:
    return "".join("%d %d %s\n" % (a, b, fmt_g(w)) for a, b, w in edges)


# --------------------------------------------------------------------------------------
# K3: knn3 neighbourhood predictor  (knn3.cpp:81-264)
# --------------------------------------------------------------------------------------


def knn3(fin_edges, test_rat: dict):
    """Returns (avg_mse float, per-movie dict).  Edges kept iff (float)w > 0.1 with obs =
    (double)(float)w (knn3.cpp:86-92); test_rat values are float->double (:100-113).
    pred(m,u) = sum_j w_mj r_j(u) / sum_j w_mj over out-neighbours j that have a test rating by
    u (:197-219); error = mean over the movie's own test ratings of (r - round(pred))^2 with
    pred < 0.1 -> 0 (:243-247), NaN -> 0 (:249-251); result = sum / num_vertices (:263) where
    vertices = endpoints of kept edges U movies with >= 1 test rating."""
    out_edges: dict[int, list] = {}
    vertices = set()
    for a, b, w in fin_edges:
        wf = np.float32(w)
        if float(wf) > 0.1:
            out_edges.setdefault(a, []).append((b, float(wf)))
            vertices.add(a)
            vertices.add(b)
    tr = {m: {u: float(np.float32(r)) for u, r in d.items()} for m, d in test_rat.items() if len(d) >= 1}
    vertices |= set(tr)
    total = np.float32(0)
    per_movie = {}
    for m in sorted(vertices):
        sr, sw = {}, {}
        for j, w in sorted(out_edges.get(m, [])):
            for u, r in tr.get(j, {}).items():
                sr[u] = sr.get(u, 0.0) + w * r
                sw[u] = sw.get(u, 0.0) + w
        own = tr.get(m, {})
        if not own:
            per_movie[m] = np.float32(0)
            continue
        err = np.float32(0)
        for u in sorted(own):
            knn = sr[u] / sw[u] if u in sr else 0.0
            if knn < 0.1:
                tmp = np.float32(0)
            else:
                tmp = np.float32(own[u] - np.floor(knn + 0.5))   # boost::math::round, knn>=0.1
            err = np.float32(err + tmp * tmp)
        e = np.float32(0) if np.isnan(err) else np.float32(err / np.float32(len(own)))
        per_movie[m] = e
        total = np.float32(total + e)
    n = len(vertices)
    return (float(total) / n if n else float("nan")), per_movie


# --------------------------------------------------------------------------------------
# fold_cross_validation.py semantics (fold_cross_validation.py:11-56), seeded
# --------------------------------------------------------------------------------------


def fold_split(ratings, num_div: int, rng: np.random.Generator):
    """User-disjoint folds: shuffle users, cut a fold every time the count in the current fold
    exceeds num_usr/num_div (:37-44).  Returns list of (train, test) triples lists."""
    by_user: dict[int, list] = {}
    for u, m, r in ratings:
        by_user.setdefault(u, []).append((u, m, r))
    keys = list(by_user.keys())
    rng.shuffle(keys)
    num_usr = len(keys)
    test = [[]]
    done = 0
    for key in keys:
        test[-1].extend(by_user[key])
        done += 1
        if done > num_usr / num_div:
            done = 0
            test.append([])
    folds = []
    for i in range(len(test)):
        train = [t for j in range(len(test)) if j != i for t in test[j]]
        folds.append((train, test[i]))
    return folds


# ---------------------------------------------------------------------------------------------------
# Chebyshev polynomial graph filter -- cheby.cpp.  Test infrastructure like the rest of this file.
# ---------------------------------------------------------------------------------------------------
def cheby_parse_topology(text: str):
    """graph_loader (cheby.cpp:86-103): lines `a b w`; a line is kept iff w > 0.1 (double) and adds BOTH directions."""
    edges = []
    for line in text.splitlines():
        tok = line.split()
        if len(tok) < 3:
            continue
        a, b, w = int(tok[0]), int(tok[1]), float(tok[2])
        if w > 0.1:
            edges.append((a, b, w))
            edges.append((b, a, w))
    return edges


def cheby_parse_signal(text: str):
    """graph_signal_loader (:105-118): lines `vertex value`."""
    sig = {}
    for line in text.splitlines():
        tok = line.split()
        if len(tok) >= 2:
            sig[int(tok[0])] = float(tok[1])
    return sig


def cheby_parse_coeff(text: str):
    """filter_loader (:120-135): every number of every line, in order."""
    return [float(t) for t in text.split()]


def cheby_filter(edges, signal: dict, coeff):
    """The three engines of cheby.cpp:312-375 on directed edges (source, target, weight) -- degree_program (:155-183),
    init_values_program (:189-227), cheby_program (:232-273, synchronous supersteps) with arange = [0, 2] (:17-19).
    Vertices that only appear in the topology have an uninitialised signal in the reference; here they read 0.
    With two coefficients the reference reads coeff[2] out of bounds (:262); here the filter stops after c1.
    Returns {vertex: filtered value}."""
    a1, a2 = (2.0 - 0.0) / 2, (2.0 + 0.0) / 2
    verts = sorted(set(signal) | {e[0] for e in edges} | {e[1] for e in edges})
    idx = {v: i for i, v in enumerate(verts)}
    nv = len(verts)
    src = np.array([idx[e[0]] for e in edges], dtype=np.int64)
    dst = np.array([idx[e[1]] for e in edges], dtype=np.int64)
    w = np.array([e[2] for e in edges], dtype=np.float64)
    deg = np.zeros(nv)
    np.add.at(deg, src, w)                                             # sum over the out-edges
    wn = w / np.sqrt(deg[dst] * deg[src]) if len(w) else w            # :209-210, :251-252

    def adj(t):                                                        # gather over the out-edges: sum wn * t[target]
        out = np.zeros(nv)
        np.add.at(out, src, wn * t[dst])
        return out

    x = np.array([signal.get(v, 0.0) for v in verts], dtype=np.float64)
    coeff = list(coeff)
    assert len(coeff) >= 2
    t_old = x.copy()
    t_cur = (x - adj(x) - a2 * x) / a1
    val = 0.5 * coeff[0] * t_old + coeff[1] * t_cur
    for k in range(2, len(coeff)):
        t_new = (2 / a1) * (t_cur - adj(t_cur) - a2 * t_cur) - t_old
        val = val + coeff[k] * t_new
        t_old, t_cur = t_cur, t_new
    return {v: float(val[idx[v]]) for v in verts}


def cheby_format(values: dict) -> str:
    """graph_signal_writer (:140-148): `id value\\n` in default stream formatting, ascending id (order is free)."""
    return "".join("%d %s\n" % (v, fmt_g(values[v])) for v in sorted(values))
