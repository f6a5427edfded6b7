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

// ---- item-stationary form of the same accumulation ------------------------------------------------
// The scatter above issues 4 global atomics per rated pair into three dense N x N arrays and gives the n = 7,359 user's
// first thread 7,358 x 4 of them.  Here a CTA owns (item a, a tile of KNN_CT columns b): it walks the raters of a (item-major
// transpose of the CSR, built by the two kernels below), and for every rater the slice of its ascending item list that
// falls into the tile (start positions per (user, tile) precomputed), accumulating cnt / num / sum r_a^2 / sum r_b^2 of
// row a in SHARED memory.  Every (a, b) is owned by exactly one CTA, so the row is finished in place: weight, emit rule
// and the text-rounded table entry come straight out of shared memory -- the N x N accumulators, their memsets and the
// global atomics are gone.  Ratings on the half-star grid (the host checks; anything else takes the scatter above) make every
// term an integer number of quarters: integer sums, converted once -- bit-identical to the reference's float accumulators as
// long as those are exact (below 2^24 quarters), i.e. exactly where knn_accumulate_kernel + knn_finalize_kernel are.
#define KNN_CT 2048

__global__ void knn_csc_count_kernel(int64_t nnz, const int32_t* __restrict__ items, int rows, unsigned long long* __restrict__ cnt) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nnz) return;
    const unsigned it = (unsigned)items[t];
    if (it < (unsigned)rows) atomicAdd(cnt + it, 1ull);
}

// grid (nu, ceil(nmax/128)), block 128: entry (u, j) -> slot of item items[j]; thread j == 0 .. ntile also writes the tile starts of u
__global__ void __launch_bounds__(128) knn_csc_fill_kernel(const int64_t* __restrict__ off, const int32_t* __restrict__ items,
                                                           const float* __restrict__ ratings, int rows, int ntile,
                                                           const int64_t* __restrict__ coff, unsigned long long* __restrict__ cursor,
                                                           int32_t* __restrict__ cuser, float* __restrict__ crat, int32_t* __restrict__ useg) {
    const int u = blockIdx.x;
    const int64_t o = off[u];
    const int n = (int)(off[u + 1] - o);
    const int j = blockIdx.y * blockDim.x + threadIdx.x;
    if (j <= ntile) {                                            // useg[u][t] = first position with item >= t * KNN_CT
        const int key = j * KNN_CT;
        int lo = 0, hi = n;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (items[o + mid] < key) lo = mid + 1; else hi = mid; }
        useg[(size_t)u * (ntile + 1) + j] = lo;
    }
    if (j >= n) return;
    const unsigned it = (unsigned)items[o + j];
    if (it >= (unsigned)rows) return;
    const unsigned long long slot = atomicAdd(cursor + it, 1ull);
    cuser[coff[it] + (int64_t)slot] = u;
    crat[coff[it] + (int64_t)slot] = ratings[o + j];
}

// grid (rows, ntile), block 256
__global__ void __launch_bounds__(256) knn_row_kernel(const int64_t* __restrict__ off, const int32_t* __restrict__ items,
                                                      const float* __restrict__ ratings, const int64_t* __restrict__ coff,
                                                      const int32_t* __restrict__ cuser, const float* __restrict__ crat,
                                                      const int32_t* __restrict__ useg, int rows, int ntile,
                                                      float* __restrict__ wf, double* __restrict__ Wd) {
    // quarter units: ratings on the half-star grid (checked by the host) -> r_a r_b, r_a^2, r_b^2 are integer multiples of 0.25;
    // integer shared-memory atomics are native (ATOMS.ADD), float ones are a CAS loop (ATOMS.CAST.SPIN)
    __shared__ int s_cnt[KNN_CT], s_num[KNN_CT], s_ab[KNN_CT], s_ba[KNN_CT];
    const int a = blockIdx.x, tile = blockIdx.y, c0 = tile * KNN_CT, c1 = min(rows, c0 + KNN_CT);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int e = tid; e < KNN_CT; e += 256) { s_cnt[e] = 0; s_num[e] = 0; s_ab[e] = 0; s_ba[e] = 0; }
    __syncthreads();
    const int64_t r0 = coff[a];
    const int nr = (int)(coff[a + 1] - r0);
    for (int t = warp; t < nr; t += 8) {
        const int u = cuser[r0 + t];
        const int ca = __float2int_rn(crat[r0 + t] * 2.f);
        const int64_t o = off[u];
        const int j0 = useg[(size_t)u * (ntile + 1) + tile], j1 = useg[(size_t)u * (ntile + 1) + tile + 1];
        for (int j = j0 + lane; j < j1; j += 32) {
            const int b = items[o + j];
            if (b == a || b >= c1) continue;
            const int cb = __float2int_rn(ratings[o + j] * 2.f);
            const int c = b - c0;
            GSI_BOUNDS(c >= 0 && c < KNN_CT);
            atomicAdd(s_cnt + c, 1);
            atomicAdd(s_num + c, ca * cb);
            atomicAdd(s_ab + c, ca * ca);
            atomicAdd(s_ba + c, cb * cb);
        }
    }
    __syncthreads();
    for (int b = c0 + tid; b < c1; b += 256) {
        const int c = b - c0;
        float w = 0.f;
        if (s_cnt[c] > 5)                                        // knn2.cpp:142-145; the float sums of :136-138 (exact below 2^24 quarter units)
            w = __fdiv_rn(__int2float_rn(s_num[c]) * 0.25f,
                          __fmul_rn(__fsqrt_rn(__int2float_rn(s_ab[c]) * 0.25f), __fsqrt_rn(__int2float_rn(s_ba[c]) * 0.25f)));
        const bool emit = (b != a) && ((double)w > 0.01);                                                    // knn2.cpp:157
        wf[(size_t)a * rows + b] = emit ? w : 0.f;
        if (Wd) Wd[(size_t)a * rows + b] = emit ? round6_text(w) : 0.0;
    }
}

// edge compaction from the dense float rows (0 = not emitted).  grid (rows), block 256; pass 0 counts, pass 1 writes at
// edge_off[a] in ascending b
__global__ void __launch_bounds__(256) knn_compact_kernel(int rows, const float* __restrict__ wf, int pass, int64_t* __restrict__ edge_cnt,
                                                          const int64_t* __restrict__ edge_off, int32_t* __restrict__ ea,
                                                          int32_t* __restrict__ eb, float* __restrict__ ew) {
    const int a = blockIdx.x;
    __shared__ int warp_tot[8];
    __shared__ int base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int b0 = 0; b0 < rows; b0 += 256) {
        const int b = b0 + threadIdx.x;
        const float w = (b < rows) ? wf[(size_t)a * rows + b] : 0.f;
        const bool emit = w != 0.f;
        const unsigned bal = __ballot_sync(0xffffffffu, emit);
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int before = base;
        for (int wv = 0; wv < warp; ++wv) before += warp_tot[wv];
        if (pass == 1 && emit) {
            const int64_t pos = edge_off[a] + before + __popc(bal & ((1u << lane) - 1u));
            ea[pos] = a; eb[pos] = b; ew[pos] = w;
        }
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int wv = 0; wv < 8; ++wv) t += warp_tot[wv]; base += t; }
        __syncthreads();
    }
    if (pass == 0 && threadIdx.x == 0) edge_cnt[a] = base;
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
