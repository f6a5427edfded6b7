// Item-item similarity stage: knn (co-rated lists), knn2 (cosine weights), knn3 (neighbourhood
// predictor).  The reference walks hash maps per graph edge (knn2.cpp:127-146); here every user
// scatters its rated pairs into dense N x N accumulators with atomics (the ratings matrix is ~1 %
// dense, so the sparse outer-product form does ~75x less work than dense Gram contractions).
//
// Exactness: knn2 accumulates num, den1, den2 in `float` (knn2.cpp:129,136-138).  For integer and
// half-star ratings every partial sum is an integer multiple of 0.25 below 2^24, hence exactly
// representable: the float atomics below are order independent and bit-identical to the
// reference's sequential loop.  sqrtf / mul / div use the IEEE round-to-nearest intrinsics.
#pragma once
#include "gsi_internal.cuh"

// grid (nu, ceil(nmax/128)), block 128: thread per item a of user u, loop over b > a.
//   cnt[a][b], num[a][b] for ia < ib (upper triangle);  S[a][b] += r_a^2, S[b][a] += r_b^2
__global__ void __launch_bounds__(128) knn_accumulate_kernel(const int64_t* __restrict__ off, const int32_t* __restrict__ items,
                                                             const float* __restrict__ ratings, int rows,
                                                             int* __restrict__ cnt, float* __restrict__ num, float* __restrict__ S) {
    const int u = blockIdx.x;
    const int64_t o = off[u];
    const int n = (int)(off[u + 1] - o);
    const int a = blockIdx.y * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const unsigned ia = (unsigned)items[o + a];
    if (ia >= (unsigned)rows) return;
    const double ra = (double)ratings[o + a];
    const float ra2 = __double2float_rn(ra * ra);
    for (int b = a + 1; b < n; ++b) {
        const unsigned ib = (unsigned)items[o + b];
        if (ib >= (unsigned)rows) continue;
        const double rb = (double)ratings[o + b];
        const size_t ab = (size_t)ia * rows + ib, ba = (size_t)ib * rows + ia;
        atomicAdd(cnt + ab, 1);
        atomicAdd(num + ab, __double2float_rn(ra * rb));
        atomicAdd(S + ab, ra2);
        atomicAdd(S + ba, __double2float_rn(rb * rb));
    }
}

// co-rated relation through train AND validate edges (knn.cpp:218-281): co[a][b] = co[b][a] = 1
__global__ void __launch_bounds__(128) knn_corated_kernel(const int64_t* __restrict__ off, const int32_t* __restrict__ items,
                                                          int rows, unsigned char* __restrict__ co) {
    const int u = blockIdx.x;
    const int64_t o = off[u];
    const int n = (int)(off[u + 1] - o);
    const int a = blockIdx.y * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const unsigned ia = (unsigned)items[o + a];
    if (ia >= (unsigned)rows) return;
    for (int b = a + 1; b < n; ++b) {
        const unsigned ib = (unsigned)items[o + b];
        if (ib >= (unsigned)rows) continue;
        co[(size_t)ia * rows + ib] = 1;
        co[(size_t)ib * rows + ia] = 1;
    }
}

// weights_calc (knn2.cpp:127-146) for the directed edge a -> b from the accumulators
__device__ __forceinline__ float knn2_weight(int a, int b, int rows, const int* cnt, const float* num, const float* S) {
    const int lo = min(a, b), hi = max(a, b);
    const int c = cnt[(size_t)lo * rows + hi];
    if (c <= 5) return 0.f;                                    // "if (num_rat > 5)" :142
    const float nm = num[(size_t)lo * rows + hi];
    const float d1 = S[(size_t)a * rows + b], d2 = S[(size_t)b * rows + a];
    return __fdiv_rn(nm, __fmul_rn(__fsqrt_rn(d1), __fsqrt_rn(d2)));
}

// value a 6-significant-digit text round trip leaves: strtod(printf("%g", w)) for w in (0.01, ~1]
__device__ __forceinline__ double round6_text(float wf) {
    const double w = (double)wf;
    const double scale = (w >= 1.0) ? 1e5 : (w >= 0.1 ? 1e6 : 1e7);
    return __ddiv_rn(rint(__dmul_rn(w, scale)), scale);
}

// grid (rows), block 256: pass 0 counts the emitted edges of row a (w > 0.01, knn2.cpp:157),
// pass 1 writes them at edge_off[a] in ascending b and (optionally) fills the dense table with the
// text-rounded weight, which is what precompute_local parses from out_fin_ (:137-145).
__global__ void __launch_bounds__(256) knn_finalize_kernel(int rows, const int* __restrict__ cnt, const float* __restrict__ num,
                                                           const float* __restrict__ S, int pass, int64_t* __restrict__ edge_cnt,
                                                           const int64_t* __restrict__ edge_off, int32_t* __restrict__ ea,
                                                           int32_t* __restrict__ eb, float* __restrict__ ew, double* __restrict__ Wd) {
    const int a = blockIdx.x;
    __shared__ int warp_tot[8];
    __shared__ int base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int b0 = 0; b0 < rows; b0 += 256) {
        const int b = b0 + threadIdx.x;
        float w = 0.f;
        bool emit = false;
        if (b < rows && b != a) { w = knn2_weight(a, b, rows, cnt, num, S); emit = (double)w > 0.01; }
        const unsigned bal = __ballot_sync(0xffffffffu, emit);
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int before = base;
        for (int wv = 0; wv < warp; ++wv) before += warp_tot[wv];
        if (pass == 1 && emit) {
            const int64_t pos = edge_off[a] + before + __popc(bal & ((1u << lane) - 1u));
            ea[pos] = a; eb[pos] = b; ew[pos] = w;
            if (Wd) Wd[(size_t)a * rows + b] = round6_text(w);
        }
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int wv = 0; wv < 8; ++wv) t += warp_tot[wv]; base += t; }
        __syncthreads();
    }
    if (pass == 0 && threadIdx.x == 0) edge_cnt[a] = base;
}

// exclusive scan of per-row edge counts (single CTA)
__global__ void __launch_bounds__(1024) knn_scan_kernel(int rows, const int64_t* __restrict__ cnt, int64_t* __restrict__ off) {
    __shared__ int64_t s[1024];
    const int t = threadIdx.x;
    const int per = (rows + 1023) / 1024;
    const int b = min(rows, t * per), e = min(rows, b + per);
    int64_t a = 0;
    for (int j = b; j < e; ++j) a += cnt[j];
    s[t] = a;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        int64_t x = (t >= o) ? s[t - o] : 0;
        __syncthreads();
        s[t] += x;
        __syncthreads();
    }
    int64_t p = s[t] - a;
    for (int j = b; j < e; ++j) { off[j] = p; p += cnt[j]; }
    if (t == 1023) off[rows] = s[1023];
}

// knn3 (knn3.cpp:185-256).  grid (nu, ceil(nmax/128)), block 128: thread per target movie m of
// test user u; neighbours j = the user's other test movies with (float)W[m][j] > 0.1.
//   pred = sum_j w r_j / sum_j w ; tmp = pred < 0.1 ? 0 : (float)(r_m - round(pred))
//   err_sum[m] += tmp*tmp (float), cnt[m] += 1
__global__ void __launch_bounds__(128) knn3_kernel(const int64_t* __restrict__ off, const int32_t* __restrict__ items,
                                                   const float* __restrict__ ratings, const double* __restrict__ W, int rows,
                                                   float* __restrict__ err_sum, int* __restrict__ cnt) {
    const int u = blockIdx.x;
    const int64_t o = off[u];
    const int n = (int)(off[u + 1] - o);
    const int a = blockIdx.y * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const unsigned m = (unsigned)items[o + a];
    if (m >= (unsigned)rows) return;
    const double* wrow = W + (size_t)m * rows;
    double sr = 0.0, sw = 0.0;
    bool any = false;
    for (int b = 0; b < n; ++b) {
        const unsigned j = (unsigned)items[o + b];
        if (j >= (unsigned)rows) continue;
        const float wf = __double2float_rn(__ldg(wrow + j));
        if ((double)wf > 0.1) {                                 // graph_loader, knn3.cpp:86-92
            const double w = (double)wf;
            sr = __dadd_rn(sr, __dmul_rn(w, (double)ratings[o + b]));
            sw = __dadd_rn(sw, w);
            any = true;
        }
    }
    float tmp = 0.f;
    if (any) {
        const double knn = __ddiv_rn(sr, sw);
        if (!(knn < 0.1)) tmp = __double2float_rn((double)ratings[o + a] - floor(knn + 0.5));
    }
    atomicAdd(err_sum + m, __fmul_rn(tmp, tmp));
    atomicAdd(cnt + m, 1);
}

// vertices of the knn3 graph: endpoints of kept edges ((float)w > 0.1).  grid (rows), block 256
__global__ void __launch_bounds__(256) knn3_vertices_kernel(const double* __restrict__ W, int rows, unsigned char* __restrict__ has_edge) {
    const int a = blockIdx.x;
    bool any = false;
    for (int b = threadIdx.x; b < rows; b += 256) {
        if ((double)__double2float_rn(W[(size_t)a * rows + b]) > 0.1) { any = true; has_edge[b] = 1; }
    }
    if (any) has_edge[a] = 1;
}
