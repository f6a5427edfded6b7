"""Chebyshev polynomial graph filter (cheby.cpp): the oracle restatement against the closed form of the
polynomial on the CPU, the GPU entry point (gsi_cheby_filter_host) and the `cheby` CLI against the oracle on the GPU.

Tolerance: the reference's gather order over the out-edges is unspecified, so sums are compared at 1e-12 relative to the
largest value (fp64 everywhere); the CLI's text output must equal the oracle's text (6 significant digits)."""
import os
import subprocess

import numpy as np
import pytest

from oracle import gsi_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "collaborative_filtering_b200", "bin")


def _random_case(seed, nv=60, density=0.2, ncoef=6):
    rng = np.random.default_rng(seed)
    ids = np.sort(rng.choice(np.arange(1, 4 * nv), size=nv, replace=False))
    lines = []
    for i in range(nv):
        for j in range(i + 1, nv):
            if rng.random() < density:
                lines.append((int(ids[i]), int(ids[j]), float(np.round(rng.uniform(0.02, 1.0), 6))))   # some fall under the 0.1 threshold
    lines.append(lines[0])                                                      # a duplicate line: kept twice, like GraphLab's add_edge
    topo = "".join("%d %d %s\n" % (a, b, O.fmt_g(w)) for a, b, w in lines)
    signal = {int(v): float(np.round(rng.normal(3.5, 1.0), 4)) for v in ids}
    signal[int(4 * nv + 7)] = 2.5                                               # an isolated vertex: no edges at all
    sig_text = "".join("%d %s\n" % (v, O.fmt_g(x)) for v, x in signal.items())
    coeff = [float(np.round(c, 5)) for c in rng.normal(0, 1, ncoef)]
    return topo, sig_text, coeff


def _csr(edges, signal):
    verts = sorted(set(signal) | {e[0] for e in edges} | {e[1] for e in edges})
    idx = {v: i for i, v in enumerate(verts)}
    rows = [[] for _ in verts]
    for a, b, w in edges:
        rows[idx[a]].append((idx[b], w))
    off = np.zeros(len(verts) + 1, dtype=np.int64)
    for i, r in enumerate(rows):
        off[i + 1] = off[i] + len(r)
    col = np.array([c for r in rows for c, _ in r], dtype=np.int32)
    w = np.array([x for r in rows for _, x in r], dtype=np.float64)
    x = np.array([signal.get(v, 0.0) for v in verts])
    return verts, off, col, w, x


def test_oracle_matches_the_polynomial_of_the_laplacian():
    topo, sig_text, coeff = _random_case(1)
    edges = O.cheby_parse_topology(topo)
    signal = O.cheby_parse_signal(sig_text)
    assert all(e[2] > 0.1 for e in edges) and len(edges) % 2 == 0
    got = O.cheby_filter(edges, signal, coeff)
    verts = sorted(got)
    idx = {v: i for i, v in enumerate(verts)}
    W = np.zeros((len(verts), len(verts)))
    for a, b, w in edges:
        W[idx[a], idx[b]] += w                                                  # duplicates add up
    d = W.sum(1)
    s = np.where(d > 0, 1.0 / np.sqrt(np.where(d > 0, d, 1.0)), 0.0)
    L = np.eye(len(verts)) - W * s[:, None] * s[None, :]
    lam, V = np.linalg.eigh((L + L.T) / 2)
    c = np.array(coeff, dtype=np.float64)
    c[0] *= 0.5                                                                 # cheby.cpp:222: 0.5 * coeff[0]
    f = np.polynomial.chebyshev.chebval(lam - 1.0, c)                           # arange [0, 2]: a1 = a2 = 1
    x = np.array([signal.get(v, 0.0) for v in verts])
    ref = V @ (f * (V.T @ x))
    out = np.array([got[v] for v in verts])
    assert np.abs(out - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())
    assert abs(got[max(verts)] - f_isolated(coeff, signal[max(verts)])) <= 1e-12


def f_isolated(coeff, x):
    """An isolated vertex has L x = x: T1 = 0, T2 = -x, T3 = 0, T4 = x, ..."""
    t_old, t_cur, val = x, 0.0, 0.5 * coeff[0] * x
    for k in range(2, len(coeff)):
        t_new = -t_old
        val += coeff[k] * t_new
        t_old, t_cur = t_cur, t_new
    return val


def test_oracle_text_formats():
    assert O.cheby_parse_coeff("1 2.5\n-3e-1\n") == [1.0, 2.5, -0.3]
    assert O.cheby_parse_topology("1 2 0.1\n1 3 0.25\n") == [(1, 3, 0.25), (3, 1, 0.25)]      # w > 0.1 is strict
    assert O.cheby_format({7: 1.0, 3: 0.123456789}) == "3 0.123457\n7 1\n"


@pytest.mark.gpu
@pytest.mark.parametrize("seed,nv,density,ncoef", [(2, 40, 0.3, 2), (3, 150, 0.15, 3), (4, 400, 0.5, 12), (5, 33, 0.0, 5)])
def test_gpu_filter_matches_oracle(seed, nv, density, ncoef):
    from collaborative_filtering_b200.api import Context, GsiError
    topo, sig_text, coeff = _random_case(seed, nv, max(density, 0.05) if density else 0.05, ncoef)
    edges = O.cheby_parse_topology(topo) if density else []
    signal = O.cheby_parse_signal(sig_text)
    ref = O.cheby_filter(edges, signal, coeff)
    verts, off, col, w, x = _csr(edges, signal)
    ctx = Context(0)
    try:
        y = ctx.cheby_filter(off, col, w, x, coeff)
        r = np.array([ref[v] for v in verts])
        assert np.abs(y - r).max() <= 1e-12 * max(1.0, np.abs(r).max())
        assert np.array_equal(y, ctx.cheby_filter(off, col, w, x, coeff))       # run-to-run identical
        with pytest.raises(GsiError):
            ctx.cheby_filter(off, col, w, x, coeff[:1])                          # needs two coefficients
        if len(col):
            bad = col.copy()
            bad[0] = len(x)
            with pytest.raises(GsiError):
                ctx.cheby_filter(off, bad, w, x, coeff)                          # column outside the graph
    finally:
        ctx.close()


@pytest.mark.gpu
def test_cheby_cli(tmp_path):
    topo, sig_text, coeff = _random_case(6, 120, 0.2, 7)
    cwd = str(tmp_path)
    open(os.path.join(cwd, "graph_topology"), "w").write(topo)
    open(os.path.join(cwd, "graph_signal"), "w").write(sig_text)
    open(os.path.join(cwd, "coeff"), "w").write(" ".join(O.fmt_g(c) for c in coeff) + "\n")
    p = subprocess.run([os.path.join(BIN, "cheby")], cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
    assert p.returncode == 0, p.stdout.decode()
    ref = O.cheby_filter(O.cheby_parse_topology(topo), O.cheby_parse_signal(sig_text), O.cheby_parse_coeff(open(os.path.join(cwd, "coeff")).read()))
    got = O.cheby_parse_signal(open(os.path.join(cwd, "graph_filtered_signal_1_of_1")).read())
    assert sorted(got) == sorted(ref)
    scale = max(1.0, max(abs(v) for v in ref.values()))
    assert max(abs(got[v] - ref[v]) for v in ref) <= 1e-5 * scale               # the file carries 6 significant digits
    lines = open(os.path.join(cwd, "graph_filtered_signal_1_of_1")).read().splitlines()
    same = sum(1 for a, b in zip(lines, O.cheby_format(ref).splitlines()) if a == b)
    assert same >= 0.99 * len(lines)                                             # identical text except for last-digit rounding ties


def test_golden_fixture_is_reproducible(golden_dir):
    """tests/golden/cheby_small (written by tests/golden/make_golden.py): the oracle reproduces the committed output."""
    d = os.path.join(golden_dir, "cheby_small")
    out = O.cheby_filter(O.cheby_parse_topology(open(os.path.join(d, "graph_topology")).read()),
                         O.cheby_parse_signal(open(os.path.join(d, "graph_signal")).read()),
                         O.cheby_parse_coeff(open(os.path.join(d, "coeff")).read()))
    assert O.cheby_format(out) == open(os.path.join(d, "graph_filtered_signal_1_of_1")).read()
    z = np.load(os.path.join(d, "oracle.npz"))
    assert np.array_equal(z["vertex"], np.array(sorted(out))) and np.allclose(z["value"], [out[v] for v in sorted(out)], rtol=0, atol=1e-13)


@pytest.mark.gpu
def test_cli_on_the_golden_fixture(golden_dir, tmp_path):
    import shutil
    d = os.path.join(golden_dir, "cheby_small")
    for f in ("coeff", "graph_topology", "graph_signal"):
        shutil.copy(os.path.join(d, f), str(tmp_path / f))
    p = subprocess.run([os.path.join(BIN, "cheby")], cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
    assert p.returncode == 0, p.stdout.decode()
    z = np.load(os.path.join(d, "oracle.npz"))
    got = O.cheby_parse_signal(open(str(tmp_path / "graph_filtered_signal_1_of_1")).read())
    assert sorted(got) == z["vertex"].tolist()                                   # incl. the vertex without a graph_signal line
    ref = dict(zip(z["vertex"].tolist(), z["value"].tolist()))
    assert max(abs(got[v] - ref[v]) for v in ref) <= 1e-5 * max(1.0, max(abs(x) for x in ref.values()))
