"""ctypes binding of libgsi.so (the C ABI declared in include/gsi.h).

The library is built in-tree (``collaborative_filtering_b200/libgsi.so``) by
``__graft_entry__.build()`` / ``make -C collaborative_filtering_b200/csrc``.  There is no fallback:
importing this module without the shared library raises, and every call needs a CUDA device.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libgsi.so")

GSI_OK, GSI_ERR_INVALID, GSI_ERR_CUDA, GSI_ERR_NOMEM, GSI_ERR_STATE, GSI_ERR_CAPACITY, GSI_ERR_SINK = range(7)
T_NAMES = ["eig_cta", "lap", "bj_gram", "bj_inner", "bj_update", "finalize", "compact", "predict", "knn",
           "trd", "dc", "dc_gemm", "bt", "cheby", "sbr", "bt2"]

c_i64p = ctypes.POINTER(ctypes.c_int64)
c_i32p = ctypes.POINTER(ctypes.c_int32)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_f32p = ctypes.POINTER(ctypes.c_float)


class RecordChunk(ctypes.Structure):
    _fields_ = [("n_records", ctypes.c_int64), ("user_index", c_i64p), ("n", c_i32p), ("k", c_i32p),
                ("lam_off", c_i64p), ("vec_off", c_i64p), ("lam", c_f64p), ("vec", c_f64p),
                ("sig_min", c_f64p)]


RECORD_SINK = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(RecordChunk))

# every symbol include/gsi.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "gsi_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_void_p]),
    "gsi_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "gsi_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "gsi_version": (ctypes.c_char_p, []),
    "gsi_set_workspace_limit": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64]),
    "gsi_sync": (ctypes.c_int, [ctypes.c_void_p]),
    "gsi_small_max": (ctypes.c_int, [ctypes.c_void_p]),
    "gsi_set_weights_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "gsi_set_weights_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "gsi_set_weights_edges": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_int64, ctypes.POINTER(ctypes.c_int)]),
    "gsi_get_weights": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int)]),
    "gsi_precompute_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64] + [ctypes.c_void_p] * 7 +
                              [ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "gsi_precompute_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64] + [ctypes.c_void_p] * 7 +
                            [ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "gsi_precompute_stream": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                             RECORD_SINK, ctypes.c_void_p]),
    "gsi_predict_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64] + [ctypes.c_void_p] * 8 + [ctypes.c_int64, ctypes.c_void_p,
                                        ctypes.c_int64] + [ctypes.c_void_p] * 6),
    "gsi_predict_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64] + [ctypes.c_void_p] * 18),
    "gsi_local_calc_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64] + [ctypes.c_void_p] * 10),
    "gsi_knn_build_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "gsi_knn_edges_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]),
    "gsi_knn_corated_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                            ctypes.c_void_p]),
    "gsi_knn3_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64] + [ctypes.c_void_p] * 6),
    "gsi_cheby_filter_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64] + [ctypes.c_void_p] * 4 + [ctypes.c_int] + [ctypes.c_void_p] * 2),
    "gsi_group_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_void_p]),
    "gsi_group_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "gsi_group_size": (ctypes.c_int, [ctypes.c_void_p]),
    "gsi_group_ctx": (ctypes.c_void_p, [ctypes.c_void_p, ctypes.c_int]),
    "gsi_group_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "gsi_group_broadcast_path": (ctypes.c_char_p, [ctypes.c_void_p]),
    "gsi_group_set_workspace_limit": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64]),
    "gsi_group_set_weights_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "gsi_group_precompute_stream": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                                   RECORD_SINK, ctypes.c_void_p]),
    "gsi_group_predict_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64] + [ctypes.c_void_p] * 8 + [ctypes.c_int64, ctypes.c_void_p,
                                              ctypes.c_int64] + [ctypes.c_void_p] * 6),
    "gsi_timing_enable": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "gsi_timing_reset": (ctypes.c_int, [ctypes.c_void_p]),
    "gsi_timing_get": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "gsi_measure_fp64_tflops": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_f64p]),
    "gsi_debug_band": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 4),
    "gsi_debug_tc_gemm": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                         ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_void_p]),
    "gsi_debug_eigh": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_float, ctypes.c_int] +
                       [ctypes.c_void_p] * 7),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile libgsi.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc")]
    subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)
    return SO_PATH


def load():
    """Load the library and bind every declared symbol.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            "libgsi.so not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C collaborative_filtering_b200/csrc`).  There is no CPU fallback.")
    lib = ctypes.CDLL(SO_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
