"""ctypes wrapper of oracle/cpu_ref.cpp (C++ restatement of compute_eigens + thread pool).

TEST INFRASTRUCTURE ONLY -- used by tests/ and by bench.py's cpu_baseline / --impl reference legs.
PARITY UNPINNED (see cpu_ref.cpp header)."""
from __future__ import annotations

import ctypes
import os
import subprocess
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "build", "libcpu_ref.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cpu_ref.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.cpuref_precompute.restype = ctypes.c_longlong
        _lib.cpuref_hardware_threads.restype = ctypes.c_int
    return _lib


def hardware_threads() -> int:
    return int(lib().cpuref_hardware_threads())


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def precompute(weights: np.ndarray, offsets: np.ndarray, items: np.ndarray, user_ids=None,
               n_threads: int = 1, honest: bool = True, format_text: bool = False):
    """Runs the restated compute_eigens over all users.  Returns dict with sig_min [nnz], k [U],
    lam (list of k-vectors), vec (list of n x k arrays), seconds, text_bytes."""
    w = np.ascontiguousarray(weights, dtype=np.float64)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    items = np.ascontiguousarray(items, dtype=np.int32)
    nu = len(offsets) - 1
    deg = np.diff(offsets)
    slots = np.maximum(deg, 2)
    lam_off = np.zeros(nu + 1, dtype=np.int64)
    np.cumsum(slots, out=lam_off[1:])
    vec_off = np.zeros(nu + 1, dtype=np.int64)
    np.cumsum(slots * deg, out=vec_off[1:])
    sig = np.zeros(int(offsets[-1]), dtype=np.float64)
    k = np.zeros(nu, dtype=np.int32)
    lam = np.zeros(int(lam_off[-1]), dtype=np.float64)
    vec = np.zeros(int(vec_off[-1]), dtype=np.float64)
    uid = None if user_ids is None else np.ascontiguousarray(user_ids, dtype=np.int32)
    t0 = time.perf_counter()
    nbytes = lib().cpuref_precompute(
        _p(w, ctypes.c_double), ctypes.c_int(w.shape[0]), _p(offsets, ctypes.c_int64),
        _p(items, ctypes.c_int32), None if uid is None else _p(uid, ctypes.c_int32),
        ctypes.c_int(nu), ctypes.c_int(n_threads), ctypes.c_int(int(honest)),
        ctypes.c_int(int(format_text)), _p(sig, ctypes.c_double), _p(k, ctypes.c_int32),
        _p(lam, ctypes.c_double), _p(lam_off, ctypes.c_int64), _p(vec, ctypes.c_double),
        _p(vec_off, ctypes.c_int64))
    dt = time.perf_counter() - t0
    lams = [lam[lam_off[u]: lam_off[u] + k[u]] for u in range(nu)]
    vecs = [vec[vec_off[u]: vec_off[u] + deg[u] * k[u]].reshape(int(deg[u]), int(k[u])) for u in range(nu)]
    return dict(sig_min=sig, k=k, lam=lams, vec=vecs, seconds=dt, text_bytes=int(nbytes))


def local_calc(weights: np.ndarray, offsets: np.ndarray, items: np.ndarray, ratings: np.ndarray, pair_mask=None,
               n_threads: int = 1, honest: bool = True):
    """C++ restatement of local_calc.cpp's vertex program over every requested (user, movie) pair (same arrays as
    gsi_local_calc_host).  Returns dict err, kk, pred, status (4 = no line), lim, w_lim [nnz], seconds, pairs."""
    lib().cpuref_local_calc.restype = ctypes.c_longlong
    w = np.ascontiguousarray(weights, dtype=np.float64)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    items = np.ascontiguousarray(items, dtype=np.int32)
    ratings = np.ascontiguousarray(ratings, dtype=np.float64)
    nnz = int(offsets[-1])
    mask = None if pair_mask is None else np.ascontiguousarray(pair_mask, dtype=np.uint8)
    err = np.zeros(nnz, dtype=np.float32)
    kk = np.zeros(nnz, dtype=np.int32)
    pred = np.zeros(nnz, dtype=np.float64)
    status = np.zeros(nnz, dtype=np.int32)
    lim = np.zeros(nnz, dtype=np.int32)
    w_lim = np.zeros(nnz, dtype=np.float64)
    t0 = time.perf_counter()
    done = lib().cpuref_local_calc(
        _p(w, ctypes.c_double), ctypes.c_int(w.shape[0]), _p(offsets, ctypes.c_int64), _p(items, ctypes.c_int32),
        _p(ratings, ctypes.c_double), ctypes.c_int(len(offsets) - 1), None if mask is None else _p(mask, ctypes.c_uint8),
        ctypes.c_int(n_threads), ctypes.c_int(int(honest)), _p(err, ctypes.c_float), _p(kk, ctypes.c_int32),
        _p(pred, ctypes.c_double), _p(status, ctypes.c_int32), _p(lim, ctypes.c_int32), _p(w_lim, ctypes.c_double))
    return dict(err=err, kk=kk, pred=pred, status=status, lim=lim, w_lim=w_lim, seconds=time.perf_counter() - t0, pairs=int(done))
