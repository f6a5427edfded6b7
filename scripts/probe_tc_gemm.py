"""Timing of the tcgen05 FP64-equivalent GEMM (gsi_debug_tc_gemm) on D&C-merge-like shapes:
    python scripts/probe_tc_gemm.py [--quick] > gpurun_out/tc_gemm_probe.json
Per shape and slice count: CUDA-event time of the slicing kernels and of the tcgen05 kernel (median-free: `reps` back-to-back
repetitions, time of the last), FP64-equivalent TFLOP/s (2 M N K), the INT8 tensor rate actually executed, max error against
numpy, next to the live DMMA / FMA FP64 peaks of the same device."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from collaborative_filtering_b200.api import Context

quick = "--quick" in sys.argv
ctx = Context(0)
out = {"fp64_fma_tflops": ctx.measure_fp64_tflops(False), "fp64_dmma_tflops": ctx.measure_fp64_tflops(True), "shapes": []}
rng = np.random.default_rng(1)
shapes = [(4096, 4096, 4096)] if quick else [(1024, 1024, 1024), (2048, 2048, 2048), (4096, 4096, 4096), (3712, 4416, 3008), (7424, 4416, 3712), (4096, 4096, 512), (4096, 4096, 128)]
for (m, n, k) in shapes:
    a = rng.standard_normal((m, k)) / np.sqrt(k)
    b = rng.standard_normal((k, n)) / np.sqrt(k)
    t0 = time.time()
    ref = a @ b
    t_np = time.time() - t0
    for s in ((8,) if quick else (8, 7, 6)):
        c, ms_s, ms_g = ctx.debug_tc_gemm(a, b, slices=s, reps=4)
        pairs = s * (s + 1) // 2
        out["shapes"].append({"m": m, "n": n, "k": k, "slices": s, "ms_slice": ms_s, "ms_gemm": ms_g,
                              "fp64_equiv_tflops_gemm": 2.0 * m * n * k / (ms_g * 1e-3) / 1e12,
                              "fp64_equiv_tflops_with_slicing": 2.0 * m * n * k / ((ms_g + ms_s) * 1e-3) / 1e12,
                              "int8_tops_executed": 2.0 * m * n * k * pairs / (ms_g * 1e-3) / 1e12,
                              "max_abs_err": float(np.abs(c - ref).max()), "numpy_s": t_np})
        print(out["shapes"][-1], file=sys.stderr, flush=True)
ctx.close()
print(json.dumps(out))
