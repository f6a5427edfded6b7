"""Drop-in boundary: the six CLI hosts run in a scratch cwd on the golden inputs with the reference's
file names, and their outputs are compared with the oracle's golden files
(knn -> knn2 -> precompute_local[_threads] -> local_calc_precomp -> knn3, run_test_precompute.sh:15-19)."""
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import gsi_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "collaborative_filtering_b200", "bin")


def _run(tool, cwd, *args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([os.path.join(BIN, tool), *args], cwd=cwd, env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert p.returncode == 0, p.stdout.decode()
    return p.stdout.decode()


@pytest.mark.parametrize("case", ["tiny_int", "tiny_half"])
def test_pipeline_cli(tmp_path, golden_dir, case):
    g = os.path.join(golden_dir, case)
    cwd = str(tmp_path)
    shutil.copytree(os.path.join(g, "movielens"), os.path.join(cwd, "movielens"))
    z = np.load(os.path.join(g, "oracle.npz"))
    meta = json.load(open(os.path.join(g, "meta.json")))
    # knn: byte-identical rating lists and neighbour lists (ascending order is our stated tie-break)
    _run("knn", cwd)
    for name in ("out_rat_1_of_1", "out_test_rat_1_of_1", "out_edg_1_of_1"):
        assert open(os.path.join(cwd, name)).read() == open(os.path.join(g, name)).read(), name
    # knn2: byte-identical out_fin_
    _run("knn2", cwd)
    assert open(os.path.join(cwd, "out_fin_1_of_1")).read() == open(os.path.join(g, "out_fin_1_of_1")).read()
    # precompute_local_threads: usage + exit 1 without the argument (precompute_local_threads.cpp:217-220)
    p = subprocess.run([os.path.join(BIN, "precompute_local_threads")], cwd=cwd, stdout=subprocess.PIPE)
    assert p.returncode == 1 and b"n_threads" in p.stdout
    for tool, args in (("precompute_local", ["8"]), ("precompute_local_threads", ["3"])):
        _run(tool, cwd, *args)
        got = O.parse_out_eigen(os.path.join(cwd, "out_eigen_"), bug_b1=False)
        ref = O.parse_out_eigen(os.path.join(g, "out_eigen_"), bug_b1=False)
        assert sorted(got) == sorted(ref)
        text = open(os.path.join(cwd, "out_eigen_")).read()
        assert text.count("\n") == 3 * len(ref) and text.count(" \n") == 3 * len(ref)     # trailing space before \n
        for u in ref:
            assert np.array_equal(got[u]["items"], ref[u]["items"])
            assert np.array_equal(got[u]["sigs_min"], ref[u]["sigs_min"])        # identical 6-digit text
            assert got[u]["lam"].shape == ref[u]["lam"].shape
            assert np.abs(got[u]["lam"] - ref[u]["lam"]).max() <= 2e-6          # 6 significant digits
            p1, p2 = got[u]["vec"] @ got[u]["vec"].T, ref[u]["vec"] @ ref[u]["vec"].T
            if len(ref[u]["lam"]) < len(ref[u]["items"]):
                assert np.abs(p1 - p2).max() <= 1e-4
    # predictor, stage-wise: feed the ORACLE's out_eigen_ so both sides consume identical records
    shutil.copy(os.path.join(g, "out_eigen_"), os.path.join(cwd, "out_eigen_"))
    for b1, tag, extra in ((True, "on", []), (False, "off", ["--fix-b1"])):
        out = _run("local_calc_precomp", cwd, "--pct", "100", *extra)
        assert "Update Rate" in out
        rows = {}
        for line in open(os.path.join(cwd, "out_res_1_of_1")):
            m, u, e, kk = line.split()
            rows[(int(m), int(u))] = (float(e), int(kk))
        n = len(z["res_%s_movie" % tag])
        assert len(rows) == n
        se_g = se_o = 0.0
        ok = 0
        for i in range(n):
            key = (int(z["res_%s_movie" % tag][i]), int(z["res_%s_user" % tag][i]))
            e, kk = rows[key]
            assert kk == z["res_%s_kk" % tag][i]
            if z["res_%s_status" % tag][i] == O.PRED_OK:
                eo = float(z["res_%s_err" % tag][i])
                assert abs(e - eo) <= 2e-5 * max(1.0, eo)
                se_g += e
                se_o += eo
                ok += 1
        assert ok > 0 and abs(np.sqrt(se_g / ok) - np.sqrt(se_o / ok)) <= 1e-4     # north_star: RMSE to 1e-4
    # --pct with a fixed seed samples movie vertices reproducibly
    _run("local_calc_precomp", cwd, "--pct", "50", env={"GSI_SEED": "7"})
    a = open(os.path.join(cwd, "out_res_1_of_1")).read()
    _run("local_calc_precomp", cwd, "50", env={"GSI_SEED": "7"})
    assert a == open(os.path.join(cwd, "out_res_1_of_1")).read() and 0 < a.count("\n") < n
    # knn3
    out = _run("knn3", cwd)
    val = float(out.strip().split("Knn Average MSE:")[1])
    assert abs(val - meta["knn3_avg_mse"]) <= 2e-6 * max(1.0, meta["knn3_avg_mse"]) + 5e-6


def _parse_eigen_bin(path):
    """out_eigen_.bin (precompute_common.hpp): "GSIEIG01", then u32 user', i32 n, i32 k, i32 0, i32 movie[n] (+pad),
    f64 sig_min[n], f64 lambda[k], f64 U[n*k]."""
    buf = open(path, "rb").read()
    assert buf[:8] == b"GSIEIG01"
    p, recs = 8, {}
    while p < len(buf):
        uid, n, k, z = np.frombuffer(buf, dtype="<i4", count=4, offset=p)
        assert z == 0
        p += 16
        items = np.frombuffer(buf, dtype="<i4", count=n, offset=p)
        p += 4 * (n + (n & 1))
        sig = np.frombuffer(buf, dtype="<f8", count=n, offset=p); p += 8 * n
        lam = np.frombuffer(buf, dtype="<f8", count=k, offset=p); p += 8 * k
        vec = np.frombuffer(buf, dtype="<f8", count=n * k, offset=p).reshape(n, k); p += 8 * n * k
        recs[int(np.uint32(uid))] = {"items": items, "sigs_min": sig, "lam": lam, "vec": vec}
    return recs


def test_binary_out_eigen(tmp_path, golden_dir):
    """SURVEY.md 8f.1 / README.md:29: GSI_EIGEN_BINARY=1 writes out_eigen_.bin (full doubles) instead of the text and
    local_calc_precomp reads it back; the text path stays the default and agrees with it to its 6 digits."""
    g = os.path.join(golden_dir, "tiny_int")
    cwd = str(tmp_path)
    shutil.copytree(os.path.join(g, "movielens"), os.path.join(cwd, "movielens"))
    _run("knn", cwd)
    _run("knn2", cwd)
    _run("precompute_local", cwd)
    text = O.parse_out_eigen(os.path.join(cwd, "out_eigen_"), bug_b1=False)
    _run("local_calc_precomp", cwd, "--pct", "100", "--fix-b1")
    res_text = open(os.path.join(cwd, "out_res_1_of_1")).read()
    assert not os.path.exists(os.path.join(cwd, "out_eigen_.bin"))
    env = {"GSI_EIGEN_BINARY": "1"}
    _run("precompute_local_threads", cwd, "4", env=env)
    assert os.path.getsize(os.path.join(cwd, "out_eigen_")) == 0          # no stale text next to the binary records
    recs = _parse_eigen_bin(os.path.join(cwd, "out_eigen_.bin"))
    assert sorted(recs) == sorted(text)
    for u, t in text.items():
        b = recs[u]
        assert np.array_equal(b["items"], t["items"]) and b["lam"].shape == t["lam"].shape and b["vec"].shape == t["vec"].shape
        assert np.allclose(b["sigs_min"], t["sigs_min"], rtol=1e-5, atol=0)       # the text keeps 6 significant digits
        assert np.allclose(b["lam"], t["lam"], rtol=1e-5, atol=1e-12)
        n, k = b["vec"].shape
        assert np.abs(b["vec"].T @ b["vec"] - np.eye(k)).max() <= 1e-9 or n < k     # full precision: orthonormal to fp64 level
    _run("local_calc_precomp", cwd, "--pct", "100", "--fix-b1", env=env)
    rows_b = [l.split() for l in open(os.path.join(cwd, "out_res_1_of_1"))]
    rows_t = [l.split() for l in res_text.splitlines()]
    assert [(r[0], r[1], r[3]) for r in rows_b] == [(r[0], r[1], r[3]) for r in rows_t]   # same pairs, same kk
    eb = np.array([float(r[2]) for r in rows_b]); et = np.array([float(r[2]) for r in rows_t])
    fin = np.isfinite(eb) & np.isfinite(et)
    assert fin.sum() > 0 and np.median(np.abs(eb[fin] - et[fin])) <= 1e-3
    # a later text run removes the binary file again
    _run("precompute_local", cwd)
    assert not os.path.exists(os.path.join(cwd, "out_eigen_.bin")) and os.path.getsize(os.path.join(cwd, "out_eigen_")) > 0
