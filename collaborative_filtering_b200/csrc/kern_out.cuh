// Output offsets (exclusive scan over k and n*k in processing order) and compaction of the padded
// per-user regions into the contiguous record layout of gsi.h, plus small utility kernels.
#pragma once
#include "gsi_internal.cuh"

struct OutJobs {
    int nj;
    const int32_t* n;        // [nj]
    const int32_t* k;        // [nj]
    const int64_t* user;     // [nj] index in the caller's CSR
    const int64_t* vec_pad;  // [nj] offset of the padded n x k block
    const int64_t* lam_pad;  // [nj]
    int64_t* vec_dst;        // [nj] out: offset in the final vec array
    int64_t* lam_dst;        // [nj]
};

// single CTA of 1024 threads.  totals[0..1] (device) hold the running lam / vec totals and are
// advanced by this chunk; totals[2] is a capacity-overflow flag.
__global__ void __launch_bounds__(1024) out_scan_kernel(OutJobs J, int64_t* totals, int64_t lam_cap,
                                                        int64_t vec_cap, int32_t* out_k,
                                                        int64_t* out_lam_off, int64_t* out_vec_off) {
    __shared__ int64_t s_lam[1024], s_vec[1024];
    const int t = threadIdx.x;
    const int per = (J.nj + 1023) / 1024;
    const int b = min(J.nj, t * per), e = min(J.nj, b + per);
    int64_t a_lam = 0, a_vec = 0;
    for (int j = b; j < e; ++j) { a_lam += J.k[j]; a_vec += (int64_t)J.k[j] * J.n[j]; }
    s_lam[t] = a_lam; s_vec[t] = a_vec;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {       // Hillis-Steele inclusive scan
        int64_t x = 0, y = 0;
        if (t >= o) { x = s_lam[t - o]; y = s_vec[t - o]; }
        __syncthreads();
        s_lam[t] += x; s_vec[t] += y;
        __syncthreads();
    }
    int64_t p_lam = totals[0] + s_lam[t] - a_lam, p_vec = totals[1] + s_vec[t] - a_vec;
    for (int j = b; j < e; ++j) {
        J.lam_dst[j] = p_lam; J.vec_dst[j] = p_vec;
        const int64_t u = J.user[j];
        out_k[u] = J.k[j]; out_lam_off[u] = p_lam; out_vec_off[u] = p_vec;
        p_lam += J.k[j]; p_vec += (int64_t)J.k[j] * J.n[j];
    }
    __syncthreads();
    if (t == 1023) {
        const int64_t nl = totals[0] + s_lam[1023], nv = totals[1] + s_vec[1023];
        if (nl > lam_cap || nv > vec_cap) totals[2] = 1;
        totals[0] = nl; totals[1] = nv;
    }
}

// grid (nj, slices), block 256: copy the job's n*k doubles (slice of 16384) and its k eigenvalues
#define GSI_COMPACT_SLICE 16384
__global__ void __launch_bounds__(256) out_compact_kernel(OutJobs J, const double* __restrict__ vec_pad,
                                                          const double* __restrict__ lam_pad,
                                                          double* __restrict__ vec, double* __restrict__ lam,
                                                          int64_t lam_cap, int64_t vec_cap) {
    const int j = blockIdx.x;
    const int k = J.k[j];
    const int64_t cnt = (int64_t)J.n[j] * k;
    const int64_t b = (int64_t)blockIdx.y * GSI_COMPACT_SLICE;
    if (b >= cnt && blockIdx.y > 0) return;
    const int64_t vd = J.vec_dst[j], ld_ = J.lam_dst[j];
    if (vd + cnt > vec_cap || ld_ + k > lam_cap) return;      // overflow flagged by the scan
    const double* src = vec_pad + J.vec_pad[j];
    double* dst = vec + vd;
    const int64_t e = min(cnt, b + GSI_COMPACT_SLICE);
    for (int64_t t = b + threadIdx.x; t < e; t += 256) dst[t] = src[t];
    if (blockIdx.y == 0)
        for (int t = threadIdx.x; t < k; t += 256) lam[ld_ + t] = lam_pad[J.lam_pad[j] + t];
}

// dense table from an edge list (gsi_set_weights_edges)
__global__ void scatter_edges_kernel(const int32_t* __restrict__ a, const int32_t* __restrict__ b,
                                     const double* __restrict__ w, int64_t ne, double* __restrict__ W, int rows) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < ne) W[(size_t)a[e] * rows + b[e]] = w[e];
}

// FP64 throughput probes (roofline denominators for the Jacobi kernels)
__global__ void __launch_bounds__(256) fp64_fma_probe(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void __launch_bounds__(256) fp64_dmma_probe(double* out, int iters) {
    double d[8][2];
#pragma unroll
    for (int t = 0; t < 8; ++t) { d[t][0] = threadIdx.x * 1e-9; d[t][1] = t; }
    const double a = 1.0000001, b = 0.25;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int t = 0; t < 8; ++t)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(d[t][0]), "+d"(d[t][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) s += d[t][0] + d[t][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
