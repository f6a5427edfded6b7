"""Python mirror of the C ABI (tests, bench, Python callers).  All compute happens in libgsi.so on
the GPU; numpy / torch are used only for buffers.  The unit of work follows the reference:
``Context.precompute`` == compute_eigens() over a batch of users
(precompute_local_threads.cpp:100-213) and returns what the out_eigen_ records hold."""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np

from . import _lib


class GsiError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("gsi error %d: %s" % (code, msg))
        self.code = code


@dataclass
class Records:
    """Batch of out_eigen_ records (README.md:14-19).  Record of user u:
    sig_min[offsets[u]:offsets[u+1]], lam[lam_off[u]:+k[u]], vec[vec_off[u]:+n*k].reshape(n,k)."""

    offsets: np.ndarray
    items: np.ndarray
    sig_min: np.ndarray
    k: np.ndarray
    lam_off: np.ndarray
    vec_off: np.ndarray
    lam: np.ndarray
    vec: np.ndarray

    def n(self, u: int) -> int:
        return int(self.offsets[u + 1] - self.offsets[u])

    def lam_of(self, u: int) -> np.ndarray:
        return self.lam[self.lam_off[u]: self.lam_off[u] + self.k[u]]

    def vec_of(self, u: int) -> np.ndarray:
        n, k = self.n(u), int(self.k[u])
        return self.vec[self.vec_off[u]: self.vec_off[u] + n * k].reshape(n, k)

    def sig_of(self, u: int) -> np.ndarray:
        return self.sig_min[self.offsets[u]: self.offsets[u + 1]]


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return ctypes.c_void_p(a.ctypes.data)
    return ctypes.c_void_p(int(a.data_ptr()))     # torch tensor


def upper_bounds(offsets: np.ndarray):
    """(lam_cap, vec_cap) that always suffice: sum max(n,2), sum n*max(n,2)."""
    n = np.diff(np.asarray(offsets, dtype=np.int64))
    m = np.maximum(n, 2)
    return int(m.sum()), int((n * m).sum())


def _predict_args(recs: "Records") -> dict:
    """The record arrays in the dtypes / layout the C ABI reads (copies only where needed)."""
    return dict(offsets=np.ascontiguousarray(recs.offsets, dtype=np.int64), items=np.ascontiguousarray(recs.items, dtype=np.int32),
                k=np.ascontiguousarray(recs.k, dtype=np.int32), lam_off=np.ascontiguousarray(recs.lam_off, dtype=np.int64),
                vec_off=np.ascontiguousarray(recs.vec_off, dtype=np.int64), lam=np.ascontiguousarray(recs.lam, dtype=np.float64),
                vec=np.ascontiguousarray(recs.vec, dtype=np.float64))


def _collect_records(offsets, items, call):
    """Runs a streaming precompute (``call(cb)``) and gathers the chunks into one Records (host arrays)."""
    nu = len(offsets) - 1
    lam_cap, vec_cap = upper_bounds(offsets)
    sig = np.zeros(int(offsets[-1]), dtype=np.float64)
    k = np.zeros(nu, dtype=np.int32)
    lam_off = np.zeros(nu, dtype=np.int64)
    vec_off = np.zeros(nu, dtype=np.int64)
    lam = np.zeros(lam_cap, dtype=np.float64)
    vec = np.zeros(vec_cap, dtype=np.float64)
    used = [0, 0]

    def sink(ch):
        nr = ch.n_records
        ui = np.ctypeslib.as_array(ch.user_index, shape=(nr,))
        nn = np.ctypeslib.as_array(ch.n, shape=(nr,))
        kk = np.ctypeslib.as_array(ch.k, shape=(nr,))
        lo = np.ctypeslib.as_array(ch.lam_off, shape=(nr,))
        vo = np.ctypeslib.as_array(ch.vec_off, shape=(nr,))
        for j in range(nr):
            u, n, kj = int(ui[j]), int(nn[j]), int(kk[j])
            k[u] = kj
            lam_off[u], vec_off[u] = used
            lam[used[0]: used[0] + kj] = np.ctypeslib.as_array(ctypes.cast(ctypes.addressof(ch.lam.contents) + 8 * int(lo[j]), c_f64p), shape=(kj,))
            vec[used[1]: used[1] + n * kj] = np.ctypeslib.as_array(ctypes.cast(ctypes.addressof(ch.vec.contents) + 8 * int(vo[j]), c_f64p), shape=(n * kj,))
            o = int(offsets[u])
            sig[o: o + n] = np.ctypeslib.as_array(ctypes.cast(ctypes.addressof(ch.sig_min.contents) + 8 * o, c_f64p), shape=(n,))
            used[0] += kj
            used[1] += n * kj
        return 0

    call(sink)
    return Records(offsets, items, sig, k, lam_off, vec_off, lam[: used[0]], vec[: used[1]])


c_f64p = ctypes.POINTER(ctypes.c_double)


class Group:
    """Every GPU of the box behind one call (gsi_group_*): the analogue of the reference's thread pool over all workers
    (precompute_local_threads.cpp:300-314).  ``devices=None`` takes every visible device."""

    def __init__(self, devices=None):
        self._lib = _lib.load()
        h = ctypes.c_void_p()
        if devices is None:
            rc = self._lib.gsi_group_create(ctypes.byref(h), 0, None)
        else:
            arr = (ctypes.c_int * len(devices))(*devices)
            rc = self._lib.gsi_group_create(ctypes.byref(h), len(devices), ctypes.cast(arr, ctypes.c_void_p))
        if rc != 0:
            raise GsiError(rc, (self._lib.gsi_last_error(None) or b"").decode())
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gsi_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise GsiError(rc, (self._lib.gsi_group_last_error(self._h) or b"").decode())

    @property
    def size(self) -> int:
        return int(self._lib.gsi_group_size(self._h))

    @property
    def broadcast_path(self) -> str:
        return self._lib.gsi_group_broadcast_path(self._h).decode()

    def set_workspace_limit(self, nbytes: int):
        self._check(self._lib.gsi_group_set_workspace_limit(self._h, nbytes))

    def set_weights(self, w: np.ndarray):
        w = np.ascontiguousarray(w, dtype=np.float64)
        assert w.ndim == 2 and w.shape[0] == w.shape[1]
        self._check(self._lib.gsi_group_set_weights_host(self._h, _ptr(w), w.shape[0]))

    def precompute(self, offsets, items) -> Records:
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        items = np.ascontiguousarray(items, dtype=np.int32)

        def call(sink):
            def _cb(_opaque, chunk_p):
                return int(sink(chunk_p.contents) or 0)
            cb = _lib.RECORD_SINK(_cb)
            self._check(self._lib.gsi_group_precompute_stream(self._h, len(offsets) - 1, _ptr(offsets), _ptr(items), cb, None))

        return _collect_records(offsets, items, call)

    def predict(self, recs: Records, ratings, w_lim=None, pair_mask=None) -> dict:
        nnz = int(recs.offsets[-1])
        ratings = np.ascontiguousarray(ratings, dtype=np.float64)
        w_lim = np.ascontiguousarray(recs.sig_min if w_lim is None else w_lim, dtype=np.float64)
        mask = None if pair_mask is None else np.ascontiguousarray(pair_mask, dtype=np.uint8)
        err = np.zeros(nnz, dtype=np.float32)
        kk = np.zeros(nnz, dtype=np.int32)
        pred = np.zeros(nnz, dtype=np.float64)
        status = np.zeros(nnz, dtype=np.int32)
        cols = np.zeros(nnz, dtype=np.int32)
        a = _predict_args(recs)
        self._check(self._lib.gsi_group_predict_host(
            self._h, len(a["offsets"]) - 1, _ptr(a["offsets"]), _ptr(a["items"]), _ptr(w_lim), _ptr(ratings),
            _ptr(a["k"]), _ptr(a["lam_off"]), _ptr(a["vec_off"]), _ptr(a["lam"]), len(a["lam"]), _ptr(a["vec"]), len(a["vec"]),
            _ptr(mask), _ptr(err), _ptr(kk), _ptr(pred), _ptr(status), _ptr(cols)))
        return dict(err=err, kk=kk, pred=pred, status=status, cols=cols)


class Context:
    """One per GPU (gsi_create).  ``stream``: a raw cudaStream_t handle (e.g.
    ``torch.cuda.current_stream().cuda_stream``) or None for a private stream."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._lib = _lib.load()
        h = ctypes.c_void_p()
        rc = self._lib.gsi_create(ctypes.byref(h), device, ctypes.c_void_p(stream) if stream else None)
        if rc != 0:
            raise GsiError(rc, (self._lib.gsi_last_error(None) or b"").decode())
        self._h = h
        self.device = device
        self._keep = []

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gsi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise GsiError(rc, (self._lib.gsi_last_error(self._h) or b"").decode())

    @property
    def version(self) -> str:
        return self._lib.gsi_version().decode()

    def set_workspace_limit(self, nbytes: int):
        self._check(self._lib.gsi_set_workspace_limit(self._h, nbytes))

    @property
    def small_max(self) -> int:
        """Largest n handled by the CTA-resident Jacobi kernel (gsi_small_max)."""
        return int(self._lib.gsi_small_max(self._h))

    def sync(self):
        self._check(self._lib.gsi_sync(self._h))

    # ---- weights (precompute_local.cpp:113-158) ----
    def set_weights(self, w):
        """Dense (N+1)x(N+1) float64 table: numpy array (copied to the device) or CUDA torch tensor
        (borrowed)."""
        if isinstance(w, np.ndarray):
            w = np.ascontiguousarray(w, dtype=np.float64)
            assert w.ndim == 2 and w.shape[0] == w.shape[1]
            self._check(self._lib.gsi_set_weights_host(self._h, _ptr(w), w.shape[0]))
        else:
            assert w.is_cuda and w.is_contiguous() and w.dim() == 2 and w.shape[0] == w.shape[1]
            assert str(w.dtype) == "torch.float64"
            self._keep = [w]
            self._check(self._lib.gsi_set_weights_device(self._h, _ptr(w), w.shape[0]))

    def set_weights_edges(self, m1, m2, w) -> int:
        m1 = np.ascontiguousarray(m1, dtype=np.int32)
        m2 = np.ascontiguousarray(m2, dtype=np.int32)
        w = np.ascontiguousarray(w, dtype=np.float64)
        rows = ctypes.c_int(0)
        self._check(self._lib.gsi_set_weights_edges(self._h, _ptr(m1), _ptr(m2), _ptr(w), len(w), ctypes.byref(rows)))
        return rows.value

    # ---- precompute ----
    def precompute(self, offsets, items) -> Records:
        """Host arrays in, host arrays out (gsi_precompute_host)."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        items = np.ascontiguousarray(items, dtype=np.int32)
        nu = len(offsets) - 1
        lam_cap, vec_cap = upper_bounds(offsets)
        sig = np.zeros(int(offsets[-1]), dtype=np.float64)
        k = np.zeros(nu, dtype=np.int32)
        lam_off = np.zeros(nu, dtype=np.int64)
        vec_off = np.zeros(nu, dtype=np.int64)
        lam = np.zeros(lam_cap, dtype=np.float64)
        vec = np.zeros(vec_cap, dtype=np.float64)
        tot = np.zeros(2, dtype=np.int64)
        self._check(self._lib.gsi_precompute_host(self._h, nu, _ptr(offsets), _ptr(items), _ptr(sig), _ptr(k),
                                                  _ptr(lam_off), _ptr(vec_off), _ptr(lam), lam_cap, _ptr(vec), vec_cap,
                                                  _ptr(tot)))
        return Records(offsets, items, sig, k, lam_off, vec_off, lam[: tot[0]], vec[: tot[1]])

    def precompute_device(self, offsets_host, d_items, d_sig, d_k, d_lam_off, d_vec_off, d_lam, d_vec):
        """Everything resident on the device (torch CUDA tensors); returns (lam_used, vec_used)."""
        offsets_host = np.ascontiguousarray(offsets_host, dtype=np.int64)
        tot = np.zeros(2, dtype=np.int64)
        self._check(self._lib.gsi_precompute_device(
            self._h, len(offsets_host) - 1, _ptr(offsets_host), _ptr(d_items), _ptr(d_sig), _ptr(d_k),
            _ptr(d_lam_off), _ptr(d_vec_off), _ptr(d_lam), d_lam.numel(), _ptr(d_vec), d_vec.numel(), _ptr(tot)))
        return int(tot[0]), int(tot[1])

    def precompute_stream(self, offsets, items, sink):
        """gsi_precompute_stream: ``sink(chunk: RecordChunk) -> int`` is called once per chunk."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        items = np.ascontiguousarray(items, dtype=np.int32)

        def _cb(_opaque, chunk_p):
            return int(sink(chunk_p.contents) or 0)

        cb = _lib.RECORD_SINK(_cb)
        self._check(self._lib.gsi_precompute_stream(self._h, len(offsets) - 1, _ptr(offsets), _ptr(items), cb, None))

    # ---- predict (local_calc_precomp.cpp:217-380) ----
    def predict(self, recs: Records, ratings, w_lim=None, pair_mask=None) -> dict:
        """One prediction per (user, rated movie) pair of ``recs`` (gsi_predict_host).
        ``ratings`` [nnz]: the users' own ratings, aligned with recs.items.  ``w_lim`` [nnz]: cutoff
        per pair (default: the record's own sig_min, i.e. bug B1 off).  Returns arrays [nnz]."""
        nnz = int(recs.offsets[-1])
        ratings = np.ascontiguousarray(ratings, dtype=np.float64)
        w_lim = np.ascontiguousarray(recs.sig_min if w_lim is None else w_lim, dtype=np.float64)
        assert len(ratings) == nnz and len(w_lim) == nnz
        mask = None if pair_mask is None else np.ascontiguousarray(pair_mask, dtype=np.uint8)
        err = np.zeros(nnz, dtype=np.float32)
        kk = np.zeros(nnz, dtype=np.int32)
        pred = np.zeros(nnz, dtype=np.float64)
        status = np.zeros(nnz, dtype=np.int32)
        cols = np.zeros(nnz, dtype=np.int32)
        a = _predict_args(recs)          # converted arrays stay referenced until the call returns
        self._check(self._lib.gsi_predict_host(
            self._h, len(a["offsets"]) - 1, _ptr(a["offsets"]), _ptr(a["items"]), _ptr(w_lim), _ptr(ratings),
            _ptr(a["k"]), _ptr(a["lam_off"]), _ptr(a["vec_off"]), _ptr(a["lam"]), len(a["lam"]), _ptr(a["vec"]), len(a["vec"]),
            _ptr(mask), _ptr(err), _ptr(kk), _ptr(pred), _ptr(status), _ptr(cols)))
        return dict(err=err, kk=kk, pred=pred, status=status, cols=cols)

    def local_calc(self, offsets, items, ratings, pair_mask=None) -> dict:
        """Per-movie variant (local_calc.cpp; gsi_local_calc_host): one prediction per (user, movie) test
        rating given as a user CSR; the item graph is the weight table set with set_weights*.  Returns
        arrays [nnz] aligned with ``items`` (status 4 = no line in the reference's out_res)."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        items = np.ascontiguousarray(items, dtype=np.int32)
        ratings = np.ascontiguousarray(ratings, dtype=np.float64)
        nnz = int(offsets[-1])
        assert len(items) == nnz and len(ratings) == nnz
        mask = None if pair_mask is None else np.ascontiguousarray(pair_mask, dtype=np.uint8)
        err = np.zeros(nnz, dtype=np.float32)
        kk = np.zeros(nnz, dtype=np.int32)
        pred = np.zeros(nnz, dtype=np.float64)
        status = np.zeros(nnz, dtype=np.int32)
        cols = np.zeros(nnz, dtype=np.int32)
        w_lim = np.zeros(nnz, dtype=np.float64)
        self._check(self._lib.gsi_local_calc_host(
            self._h, len(offsets) - 1, _ptr(offsets), _ptr(items), _ptr(ratings), _ptr(mask),
            _ptr(err), _ptr(kk), _ptr(pred), _ptr(status), _ptr(cols), _ptr(w_lim)))
        return dict(err=err, kk=kk, pred=pred, status=status, cols=cols, w_lim=w_lim)

    def local_calc_shard(self, offsets, items, ratings, n_nodes, rank: int, world: int, pair_mask=None) -> dict:
        """This rank's share of a multi-GPU local_calc run: movie vertices are dealt over the ranks by cost
        (shard.shard_movies; ``n_nodes[m]`` = 1 + out-degree of movie m in the thresholded item graph), every rank
        passes the whole test CSR and computes the pairs of its own movies; the other pairs come back with status 4
        and are merged by the caller (all-reduce of counts / squared errors, or a gather of the arrays)."""
        from . import shard
        items = np.ascontiguousarray(items, dtype=np.int32)
        n_nodes = np.asarray(n_nodes)
        keep = np.ones(len(items), dtype=bool) if pair_mask is None else np.asarray(pair_mask).astype(bool)
        inside = items < len(n_nodes)
        n_pairs = np.bincount(items[keep & inside], minlength=len(n_nodes))
        owner = shard.shard_movies(n_nodes, n_pairs, world)
        mask = np.zeros(len(items), dtype=np.uint8)
        mask[inside] = shard.local_calc_pair_mask(items[inside], owner, rank, keep[inside])
        return self.local_calc(offsets, items, ratings, pair_mask=mask)

    # ---- knn chain (knn.cpp / knn2.cpp / knn3.cpp) ----
    def knn_build(self, offsets, items, ratings, rows: int, install_weights: bool = True):
        """knn2 over the TRAIN ratings (CSR by user).  Returns the out_fin_ edges (m1, m2, w float32)
        in ascending (m1, m2) order; optionally installs the dense table as the context's weights."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        items = np.ascontiguousarray(items, dtype=np.int32)
        ratings = np.ascontiguousarray(ratings, dtype=np.float32)
        ne = ctypes.c_int64(0)
        self._check(self._lib.gsi_knn_build_host(self._h, len(offsets) - 1, _ptr(offsets), _ptr(items), _ptr(ratings), rows,
                                                 int(install_weights), ctypes.byref(ne)))
        a = np.zeros(ne.value, dtype=np.int32)
        b = np.zeros(ne.value, dtype=np.int32)
        w = np.zeros(ne.value, dtype=np.float32)
        self._check(self._lib.gsi_knn_edges_host(self._h, _ptr(a), _ptr(b), _ptr(w), ne.value))
        return a, b, w

    def knn_corated(self, offsets, items, rows: int) -> np.ndarray:
        """knn step 1 (out_edg_): dense boolean co-rating matrix over ALL users (train + validate)."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        items = np.ascontiguousarray(items, dtype=np.int32)
        co = np.zeros((rows, rows), dtype=np.uint8)
        self._check(self._lib.gsi_knn_corated_host(self._h, len(offsets) - 1, _ptr(offsets), _ptr(items), rows, _ptr(co)))
        return co

    def knn3(self, offsets, items, ratings):
        """knn3 over the validate users.  Returns (avg_mse, err_sum[rows], cnt[rows], has_edge[rows]);
        avg_mse = sum_m (err_sum/cnt) / num_vertices exactly as knn3.cpp:234-264 (float sums,
        ascending movie id)."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        items = np.ascontiguousarray(items, dtype=np.int32)
        ratings = np.ascontiguousarray(ratings, dtype=np.float32)
        rows = ctypes.c_int(0)
        p = ctypes.c_void_p()
        self._check(self._lib.gsi_get_weights(self._h, ctypes.byref(p), ctypes.byref(rows)))
        err = np.zeros(rows.value, dtype=np.float32)
        cnt = np.zeros(rows.value, dtype=np.int32)
        has = np.zeros(rows.value, dtype=np.uint8)
        self._check(self._lib.gsi_knn3_host(self._h, len(offsets) - 1, _ptr(offsets), _ptr(items), _ptr(ratings), _ptr(err),
                                            _ptr(cnt), _ptr(has)))
        total = np.float32(0)
        for m in np.nonzero(cnt)[0]:
            e = np.float32(err[m] / np.float32(cnt[m]))
            if not np.isnan(e):
                total = np.float32(total + e)
        nv = int(np.count_nonzero((cnt > 0) | (has > 0)))
        return (float(total) / nv if nv else float("nan")), err, cnt, has

    # ---- Chebyshev graph filter (cheby.cpp) ----
    def cheby_filter(self, row_off, col, w, x, coef) -> np.ndarray:
        """y = 0.5 c0 T0 + c1 T1 + ... on the normalised Laplacian of the CSR graph (out-edges, vertices 0..nv-1),
        exactly the three engines of cheby.cpp:312-375."""
        row_off = np.ascontiguousarray(row_off, dtype=np.int64)
        col = np.ascontiguousarray(col, dtype=np.int32)
        w = np.ascontiguousarray(w, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        coef = np.ascontiguousarray(coef, dtype=np.float64)
        y = np.zeros(len(x), dtype=np.float64)
        self._check(self._lib.gsi_cheby_filter_host(self._h, len(x), _ptr(row_off), _ptr(col), _ptr(w), _ptr(x), len(coef), _ptr(coef), _ptr(y)))
        return y

    # ---- measurement ----
    def timing_enable(self, on: bool = True):
        self._check(self._lib.gsi_timing_enable(self._h, int(on)))

    def timing_reset(self):
        self._check(self._lib.gsi_timing_reset(self._h))

    def timing(self) -> dict:
        n = len(_lib.T_NAMES)
        ms = np.zeros(n, dtype=np.float64)
        launches = np.zeros(n, dtype=np.int64)
        samples = np.zeros(n, dtype=np.int64)
        self._check(self._lib.gsi_timing_get(self._h, _ptr(ms), _ptr(launches), _ptr(samples)))
        return {name: dict(ms=float(ms[i]), launches=int(launches[i]), samples=int(samples[i]))
                for i, name in enumerate(_lib.T_NAMES)}

    def measure_fp64_tflops(self, dmma: bool = False) -> float:
        v = ctypes.c_double(0)
        self._check(self._lib.gsi_measure_fp64_tflops(self._h, int(dmma), ctypes.byref(v)))
        return v.value

    # ---- stage-wise test hook of the large-n eigensolver (gsi_debug_eigh) ----
    def debug_eigh(self, a: np.ndarray, thr: float = 1e30, team: int = 0):
        """Runs tridiagonalisation / divide & conquer / back-transform on one dense symmetric matrix
        and returns dict(d, e, tau, v, lam, u, k); see include/gsi.h."""
        a = np.asfortranarray(a, dtype=np.float64)
        n = a.shape[0]
        d = np.zeros(n); e = np.zeros(max(n - 1, 1)); tau = np.zeros(max(n - 1, 1))
        v = np.zeros((n, n), order="F"); lam = np.zeros(n); u = np.zeros((n, n), order="F")
        k = ctypes.c_int32(0)
        self._check(self._lib.gsi_debug_eigh(self._h, n, _ptr(a), ctypes.c_float(thr), team, _ptr(d), _ptr(e), _ptr(tau),
                                             _ptr(v), _ptr(lam), _ptr(u), ctypes.byref(k)))
        kk = int(k.value)
        return dict(d=d, e=e[:n - 1], tau=tau[:n - 1], v=v, lam=lam, u=u[:, :kk], k=kk)

    def debug_tc_gemm(self, a: np.ndarray, b: np.ndarray, slices: int = 8, reps: int = 1):
        """C = A @ B through the tcgen05 INT8-slice engine (gsi_debug_tc_gemm).  Returns (C, ms_slice, ms_gemm)."""
        a = np.asfortranarray(a, dtype=np.float64)
        b = np.asfortranarray(b, dtype=np.float64)
        m, k = a.shape
        k2, n = b.shape
        assert k == k2
        c = np.zeros((m, n), order="F")
        ms_s, ms_g = ctypes.c_double(0.0), ctypes.c_double(0.0)
        self._check(self._lib.gsi_debug_tc_gemm(self._h, m, n, k, _ptr(a), max(m, 1), _ptr(b), max(k, 1), _ptr(c), max(m, 1), slices, reps,
                                                ctypes.byref(ms_s), ctypes.byref(ms_g)))
        return c, ms_s.value, ms_g.value

    def debug_band(self, a: np.ndarray, chase: bool = True):
        """Two-stage test hook (gsi_debug_band): returns dict(band = dense symmetric band matrix after stage 1, d, e)."""
        a = np.asfortranarray(a, dtype=np.float64)
        n = a.shape[0]
        ab = np.zeros((n, 128))
        d = np.zeros(n); e = np.zeros(max(n - 1, 1))
        self._check(self._lib.gsi_debug_band(self._h, n, _ptr(a), _ptr(ab), _ptr(d) if chase else None, _ptr(e) if chase else None))
        band = np.zeros((n, n))
        for dd in range(65):
            v = ab[: n - dd, dd]
            band[np.arange(dd, n), np.arange(0, n - dd)] = v
            band[np.arange(0, n - dd), np.arange(dd, n)] = v
        return dict(band=band, ab=ab, d=d, e=e[: n - 1])
