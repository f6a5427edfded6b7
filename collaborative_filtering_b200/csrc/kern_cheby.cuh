// Chebyshev polynomial graph filter (cheby.cpp): y = 0.5 c0 T0 + c1 T1 + ... + c_{K-1} T_{K-1} applied to a signal x on the
// normalised Laplacian L = I - D^-1/2 W D^-1/2 of one graph, with T0 = x, T1 = (L x - a2 x) / a1,
// T_{k+1} = (2 / a1)(L T_k - a2 T_k) - T_{k-1} on the interval [0, 2] (a1 = a2 = 1, cheby.cpp:17-19).
//
//   degree_program     (cheby.cpp:155-183)   d_i = sum of the out-edge weights
//   init_values_program (:189-227)           T0, T1, y = 0.5 c0 T0 + c1 T1
//   cheby_program       (:232-273)           one synchronous superstep per remaining coefficient
// The graph is a CSR over out-edges (both directions of every kept line, duplicates kept).  The edge factor
// w / sqrt(d_target * d_source) (:209-211, :251-253) does not change between supersteps and is computed once.
// One warp per row, lanes along the row (contiguous 12 bytes per edge: the stage is HBM bound), fixed shuffle tree:
// results do not depend on timing.  The supersteps are Jacobi style like the reference's "sync" engine (:327-375): every
// row reads the previous superstep's T_k.
#pragma once
#include "gsi_internal.cuh"

__global__ void cheby_degree_kernel(int nv, const int64_t* __restrict__ row_off, const double* __restrict__ w, double* __restrict__ deg) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= nv) return;
    double s = 0.0;
    for (int64_t e = row_off[row] + lane; e < row_off[row + 1]; e += 32) s += w[e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) deg[row] = s;
}

__global__ void cheby_normalise_kernel(int nv, const int64_t* __restrict__ row_off, const int32_t* __restrict__ col,
                                       const double* __restrict__ w, const double* __restrict__ deg, double* __restrict__ wn) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= nv) return;
    const double ds = deg[row];
    for (int64_t e = row_off[row] + lane; e < row_off[row + 1]; e += 32) wn[e] = w[e] / sqrt(deg[col[e]] * ds);     // :209-210
}

// first == 1: init_values_program  (t_old = x, t_cur = (x - A x - a2 x) / a1, y = 0.5 c0 t_old + c1 t_cur)
// first == 0: cheby_program        (t_new = (2 / a1)(t_cur - A t_cur - a2 t_cur) - t_old, y += c t_new)
__global__ void cheby_step_kernel(int nv, const int64_t* __restrict__ row_off, const int32_t* __restrict__ col,
                                  const double* __restrict__ wn, const double* __restrict__ t_cur, const double* __restrict__ t_old,
                                  double* __restrict__ t_new, double* __restrict__ y, double c0, double c1, int first) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= nv) return;
    double s = 0.0;
    for (int64_t e = row_off[row] + lane; e < row_off[row + 1]; e += 32) s = fma(wn[e], __ldg(t_cur + col[e]), s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        const double a1 = 1.0, a2 = 1.0;
        const double tc = t_cur[row];
        if (first) {
            const double t1 = (tc - s - a2 * tc) / a1;
            t_new[row] = t1;
            y[row] = 0.5 * c0 * tc + c1 * t1;
        } else {
            const double tn = (2.0 / a1) * (tc - s - a2 * tc) - t_old[row];
            t_new[row] = tn;
            y[row] = y[row] + c0 * tn;
        }
    }
}
