"""Workflow around the six tools: the reference's fold generator, its 5-fold driver
script and the RMSE step the reference leaves to the user, without hard-coded paths and with seeds.

  fold_cross_validation   fold_cross_validation.py:11-56 -- user-disjoint folds of a tab separated
                          `user item rating` file: users shuffled, a fold is cut every time more than
                          num_usr / num_div users have been written (:37-44), fold i is u{i}.test and
                          the other folds together are u{i}.train (:46-56).  The shuffle is seeded.
  rmse_from_out_res       the step run_test_precompute.sh:19 stops before: `cat out_res_*` and average
                          column 3.  Lines are `movie user' mse kk` (local_calc_precomp.cpp:393-404).
  mega_graph, cheby_scale mega_graph.py:27-40 and scale2.sh:5-36: seeded random graphs and the three scaling sweeps of
                          the cheby tool (coefficients, connectivity, nodes), timing lines kept as JSON.
  run_pipeline            run_test_precompute.sh:9-20: per fold, movielens/u{i}.train + u{i}.validate,
                          then knn; knn2; precompute_local 8; local_calc_precomp --pct P;
                          cat out_res_* > out_res.{i}.  The tools are the drop-in binaries under
                          collaborative_filtering_b200/bin (they need a GPU: there is no CPU path).

CLI:  python -m collaborative_filtering_b200.workflow fold  u.data 5 [--out cross_validation] [--seed S]
      python -m collaborative_filtering_b200.workflow rmse  out_res_1_of_1 [more files ...]
      python -m collaborative_filtering_b200.workflow run   cross_validation workdir [--pct 20] [--folds 0,1,2,3,4] [--variant local_calc]
      python -m collaborative_filtering_b200.workflow mega-graph 50000 0.01 [--out DIR]
      python -m collaborative_filtering_b200.workflow cheby-scale workdir [--nodes 50000] [--conn 0.01] [--coeffs 64]
"""
from __future__ import annotations

import argparse
import glob
import json
import math
import os
import random
import shutil
import subprocess
import sys
import time
from collections import OrderedDict

import numpy as np

BIN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bin")
TOOLS = ("knn", "knn2", "precompute_local", "local_calc_precomp")


def fold_cross_validation(filename: str, num_div: int, out_dir: str = "cross_validation", seed: int = 31413) -> int:
    """Writes out_dir/u{i}.test and u{i}.train; returns the number of folds written (the reference's
    `ind + 1`, which is num_div, or num_div + 1 with an empty last fold when num_div divides the
    number of users -- kept, fold_cross_validation.py:40-44)."""
    data: "OrderedDict[int, list]" = OrderedDict()
    with open(filename, "r") as f:
        for line in f:
            if not line.strip():
                continue
            val = line.split("\t")
            user_id, item_id, rating = int(val[0]), int(val[1]), int(val[2])
            data.setdefault(user_id, []).append((item_id, rating))
    num_usr = len(data)
    os.makedirs(out_dir, exist_ok=False)                     # os.mkdir in the reference: an existing directory is an error
    keys = list(data.keys())
    random.Random(seed).shuffle(keys)
    test = {0: []}
    ind, n_usr_done = 0, 0
    for key in keys:
        test[ind].extend("%d\t%d\t%d\n" % (key, p[0], p[1]) for p in data[key])
        n_usr_done += 1
        if n_usr_done > (num_usr / num_div):
            n_usr_done = 0
            ind += 1
            test[ind] = []
    for i in range(ind + 1):
        with open(os.path.join(out_dir, "u%d.test" % i), "w") as f:
            f.writelines(test[i])
        with open(os.path.join(out_dir, "u%d.train" % i), "w") as f:
            for j in range(ind + 1):
                if i != j:
                    f.writelines(test[j])
    return ind + 1


def rmse_from_out_res(paths) -> dict:
    """RMSE over out_res lines `movie user' mse kk`.  NaN lines (kk == 0: the reference's 0/0,
    local_calc_precomp.cpp:311) are counted and left out of the mean, as any offline average has to."""
    if isinstance(paths, str):
        paths = sorted(glob.glob(paths)) or [paths]
    se, n, nan, kk_sum = 0.0, 0, 0, 0
    for path in paths:
        with open(path, "r") as f:
            for line in f:
                tok = line.split()
                if len(tok) < 4:
                    continue
                mse = float(tok[2])
                if mse != mse:
                    nan += 1
                    continue
                se += mse
                n += 1
                kk_sum += int(tok[3])
    return {"rmse": math.sqrt(se / n) if n else float("nan"), "mse": se / n if n else float("nan"), "predictions": n,
            "nan": nan, "mean_kk": kk_sum / n if n else 0.0, "files": len(paths)}


def _link(src: str, dst: str) -> None:
    if os.path.lexists(dst):
        os.remove(dst)
    os.symlink(os.path.abspath(src), dst)


def run_pipeline(cross_dir: str, workdir: str, folds=None, pct: int = 20, seed: int = 31413, bin_dir: str = BIN_DIR,
                 precompute_tool: str = "precompute_local", threads: int = 8, log=sys.stderr, stage_timeout=None,
                 variant: str = "precomp") -> list:
    """One pass of run_test_precompute.sh (variant "precomp": knn; knn2; precompute_local; local_calc_precomp --pct) or of
    run_test.sh (variant "local_calc": knn; knn2; local_calc -- the per-movie tool, run_test.sh:15-17) per fold inside
    `workdir` (the tools are cwd-relative).  Returns one dict per fold: the RMSE summary of its out_res.{i} and the wall
    seconds of every tool."""
    if variant not in ("precomp", "local_calc"):
        raise ValueError("variant must be 'precomp' or 'local_calc'")
    if folds is None:
        folds = sorted(int(os.path.basename(p)[1:-6]) for p in glob.glob(os.path.join(cross_dir, "u*.train")))
    os.makedirs(workdir, exist_ok=True)
    mv = os.path.join(workdir, "movielens")
    results = []
    env = dict(os.environ, GSI_SEED=str(seed))               # --pct sampling is seeded (the reference seeds it with time())
    for i in folds:
        print("Test #%d" % i, file=log)
        shutil.rmtree(mv, ignore_errors=True)
        os.makedirs(mv)
        _link(os.path.join(cross_dir, "u%d.train" % i), os.path.join(mv, "u%d.train" % i))
        _link(os.path.join(cross_dir, "u%d.test" % i), os.path.join(mv, "u%d.validate" % i))
        for stale in glob.glob(os.path.join(workdir, "out_*_of_*")) + glob.glob(os.path.join(workdir, "out_eigen_*")):
            os.remove(stale)
        stage_s = {}
        stages = ((("knn", []), ("knn2", []), (precompute_tool, [str(threads)]), ("local_calc_precomp", ["--pct", str(pct)]))
                  if variant == "precomp" else (("knn", []), ("knn2", []), ("local_calc", ["--pct", str(pct)])))
        for tool, args in stages:
            exe = os.path.join(bin_dir, tool)
            if not os.path.exists(exe):
                raise FileNotFoundError("%s is not built (run `make -C collaborative_filtering_b200/csrc`)" % exe)
            t0 = time.perf_counter()
            subprocess.run([exe] + args, cwd=workdir, env=env, check=True, stdout=subprocess.DEVNULL, timeout=stage_timeout)
            stage_s[tool] = round(time.perf_counter() - t0, 3)
        parts = sorted(glob.glob(os.path.join(workdir, "out_res_*_of_*")))
        merged = os.path.join(workdir, "out_res.%d" % i)
        with open(merged, "w") as out:                       # cat out_res_* > out_res.$i
            for p in parts:
                with open(p, "r") as f:
                    shutil.copyfileobj(f, out)
        summary = dict(rmse_from_out_res([merged]), fold=i, seconds=stage_s)
        print(json.dumps(summary), file=log)
        results.append(summary)
    return results


def mega_graph(size: int, conn: float, out_dir: str = ".", seed: int = 31413) -> int:
    """mega_graph.py:27-40 restated with a seeded generator: `graph_signal.txt` (vertex i+1, value uniform in [0, 10))
    and `graph_topology.txt` with conn * size^2 distinct directed links (a, b), a != b, weight uniform in [0, 1) printed
    with two decimals.  Returns the number of links."""
    rng = np.random.default_rng(seed)
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "graph_signal.txt"), "w") as f:
        f.write("".join("%d %s\n" % (i + 1, repr(float(v))) for i, v in enumerate(rng.uniform(0, 10, size))))
    n_links = int(conn * size * size)
    if n_links > size * (size - 1):
        raise ValueError("more links requested than a graph of %d vertices has" % size)
    codes = np.zeros(0, dtype=np.int64)
    while len(codes) < n_links:                               # rejection sampling of distinct off-diagonal pairs (:31-37)
        need = n_links - len(codes)
        a = rng.integers(0, size, size=need + need // 8 + 16, dtype=np.int64)
        b = rng.integers(0, size, size=len(a), dtype=np.int64)
        fresh = (a * size + b)[a != b]
        codes = np.unique(np.concatenate([codes, fresh]))
        if len(codes) > n_links:
            codes = rng.permutation(codes)[:n_links]
    wei = rng.random(len(codes))
    with open(os.path.join(out_dir, "graph_topology.txt"), "w") as f:
        for lo in range(0, len(codes), 1 << 20):
            c = codes[lo: lo + (1 << 20)]
            f.write("".join("%d %d %.2f\n" % (x // size + 1, x % size + 1, w) for x, w in zip(c.tolist(), wei[lo: lo + (1 << 20)].tolist())))
    return int(len(codes))


def cheby_scale(workdir: str, nodes=50000, conn=0.01, coeffs=64, sweep_coeffs=range(10, 101, 10),
                sweep_conn=tuple(round(0.005 * i, 3) for i in range(1, 11)), sweep_nodes=range(5000, 50001, 5000),
                seed: int = 31413, bin_dir: str = BIN_DIR, log=sys.stderr, stage_timeout=None) -> list:
    """scale2.sh restated: three sweeps of the `cheby` tool on random graphs -- number of coefficients (:5-14), connectivity
    (:17-26) and number of nodes (:30-38) -- each point generating its graph with mega_graph (the first sweep reuses one
    graph), cutting the first k coefficients of a 1000-long line (`_coeff_1000.txt`, not shipped with the reference:
    seeded uniform values here) into coeff.txt, running the tool in `workdir` and keeping the two timing lines the script
    greps ("Finished in", "Final Runtime").  Returns one dict per point; also appended to workdir/scale_res2.jsonl."""
    os.makedirs(workdir, exist_ok=True)
    exe = os.path.join(bin_dir, "cheby")
    if not os.path.exists(exe):
        raise FileNotFoundError("%s is not built (run `make -C collaborative_filtering_b200/csrc`)" % exe)
    all_coeff = np.random.default_rng(seed + 1).uniform(-1.0, 1.0, 1000)
    res_path = os.path.join(workdir, "scale_res2.jsonl")
    open(res_path, "w").close()
    results = []

    def point(n, c, k, regenerate):
        if regenerate:
            mega_graph(n, c, workdir, seed)
        with open(os.path.join(workdir, "coeff.txt"), "w") as f:
            f.write(" ".join(repr(float(x)) for x in all_coeff[:k]) + "\n")
        t0 = time.perf_counter()
        p = subprocess.run([exe], cwd=workdir, check=True, stdout=subprocess.PIPE, timeout=stage_timeout)
        wall = time.perf_counter() - t0
        rec = {"nodes": int(n), "conn": float(c), "coeffs": int(k), "wall_s": round(wall, 3)}
        for line in p.stdout.decode().splitlines():
            if "Finished in" in line:
                rec["load_s"] = float(line.split()[-1])
            elif "Final Runtime" in line:
                rec["runtime_s"] = float(line.split()[-1])
        with open(res_path, "a") as f:
            f.write(json.dumps(rec) + "\n")
        print(json.dumps(rec), file=log)
        results.append(rec)

    for i, k in enumerate(sweep_coeffs):
        point(nodes, conn, k, regenerate=(i == 0))
    for c in sweep_conn:
        point(nodes, c, coeffs, regenerate=True)
    for n in sweep_nodes:
        point(n, conn, coeffs, regenerate=True)
    return results


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="collaborative_filtering_b200.workflow")
    sub = ap.add_subparsers(dest="cmd", required=True)
    a = sub.add_parser("fold", help="user-disjoint k-fold split (fold_cross_validation.py)")
    a.add_argument("filename")
    a.add_argument("num_div", type=int)
    a.add_argument("--out", default="cross_validation")
    a.add_argument("--seed", type=int, default=31413)
    b = sub.add_parser("rmse", help="RMSE over out_res files")
    b.add_argument("paths", nargs="+")
    c = sub.add_parser("run", help="knn; knn2; precompute_local; local_calc_precomp per fold (run_test_precompute.sh), or "
                                   "knn; knn2; local_calc with --variant local_calc (run_test.sh)")
    c.add_argument("cross_dir")
    c.add_argument("workdir")
    c.add_argument("--pct", type=int, default=20)
    c.add_argument("--folds", default=None)
    c.add_argument("--seed", type=int, default=31413)
    c.add_argument("--tool", default="precompute_local", choices=["precompute_local", "precompute_local_threads"])
    c.add_argument("--variant", default="precomp", choices=["precomp", "local_calc"])
    d = sub.add_parser("mega-graph", help="random graph_signal.txt / graph_topology.txt for cheby (mega_graph.py)")
    d.add_argument("size", type=int)
    d.add_argument("conn", type=float)
    d.add_argument("--out", default=".")
    d.add_argument("--seed", type=int, default=31413)
    e = sub.add_parser("cheby-scale", help="the three scaling sweeps of scale2.sh with the cheby tool")
    e.add_argument("workdir")
    e.add_argument("--nodes", type=int, default=50000)
    e.add_argument("--conn", type=float, default=0.01)
    e.add_argument("--coeffs", type=int, default=64)
    e.add_argument("--seed", type=int, default=31413)
    args = ap.parse_args(argv)
    if args.cmd == "mega-graph":
        print(mega_graph(args.size, args.conn, args.out, args.seed))
        return 0
    if args.cmd == "cheby-scale":
        res = cheby_scale(args.workdir, args.nodes, args.conn, args.coeffs,
                          sweep_nodes=range(args.nodes // 10, args.nodes + 1, args.nodes // 10), seed=args.seed)
        print(json.dumps(res))
        return 0
    if args.cmd == "fold":
        print(fold_cross_validation(args.filename, args.num_div, args.out, args.seed))
    elif args.cmd == "rmse":
        print(json.dumps(rmse_from_out_res(args.paths)))
    else:
        folds = [int(x) for x in args.folds.split(",")] if args.folds else None
        res = run_pipeline(args.cross_dir, args.workdir, folds, args.pct, args.seed, precompute_tool=args.tool, variant=args.variant)
        tot = sum(r["mse"] * r["predictions"] for r in res if r["predictions"])
        cnt = sum(r["predictions"] for r in res)
        print(json.dumps({"folds": res, "rmse": math.sqrt(tot / cnt) if cnt else float("nan"), "predictions": cnt}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
