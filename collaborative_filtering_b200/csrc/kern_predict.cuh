// Band-limited least-squares rating prediction, one CTA per (movie, test-user) pair:
// the body of neigh_program::apply(), local_calc_precomp.cpp:230-360.
//
//   K      = rows j of the user's record whose movie is an out-neighbour of the target movie m in
//            the thresholded item graph ((float)w > 0.1, :129-133); read straight from the dense
//            table row W[m, :] -- the graph and the table come from the same out_fin_ lines
//   lim    = first l with lambda_l > w_lim, >= 2                                  (:271-279)
//   cols   = columns l < lim with some A(row, l) >= 1e-4 on K (signed test)       (:284-304)
//   M      = A^T A, rhs = A^T (r_K - mean), x = M^-1 rhs, pred = v.x + mean       (:308-315)
//   err    = (real - clamp(pred,1,5))^2 as float, kk = #K                          (:318-360)
// M is symmetric positive semi-definite; it is factored by Cholesky (the reference inverts it by
// LU without a singularity check, B3).  A non-positive pivot or kk < c marks the prediction
// ill-posed: status says so and the rule value (mean of the known ratings) is returned.
#pragma once
#include "gsi_internal.cuh"
#include "kern_eig_cta.cuh"

#define GSI_PRED_CHUNK 16     // rows of A staged per Gram step

struct PredParams {
    const double* W; int w_rows;
    const int32_t* items; const int64_t* offsets;
    const double* w_lim;            // [nnz] cutoff used for the pair (user, row)  (B1 lives here)
    const double* ratings;          // [nnz] the user's own rating of each rated movie
    const int32_t* k; const int64_t* lam_off; const int64_t* vec_off;
    const double* lam; const double* vec;
    const int64_t* task_pair;       // [ntasks] flat pair index (into nnz arrays)
    const int32_t* task_user;       // [ntasks]
    float* err; int32_t* kk; double* pred; int32_t* status; int32_t* cols_used;
    double* work; int64_t work_stride;   // per-CTA global scratch for M when it does not fit smem
    int cmax, nmax, m_in_smem, task_base;
};

// deterministic block-wide sum (fixed order: lanes by shuffle tree, warps ascending)
__device__ __forceinline__ double block_sum_256(double v, double* wsum) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    const int nw = blockDim.x >> 5;
    for (int w = 0; w < nw; ++w) s += wsum[w];
    return s;
}

__global__ void __launch_bounds__(256) predict_kernel(PredParams P) {
    extern __shared__ double sm[];
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const int task = blockIdx.x + P.task_base;
    const int64_t pair = P.task_pair[task];
    const int u = P.task_user[task];
    const int64_t off = P.offsets[u];
    const int n = (int)(P.offsets[u + 1] - off);
    const int k = P.k[u];
    const double* U = P.vec + P.vec_off[u];
    const double* lam = P.lam + P.lam_off[u];
    const int mrow = (int)(pair - off);
    const unsigned m = (unsigned)P.items[pair];
    // shared layout
    const int cmax = P.cmax;
    double* M = P.m_in_smem ? sm : P.work + (size_t)blockIdx.x * P.work_stride;
    double* As = sm + (P.m_in_smem ? (size_t)cmax * cmax : 0);      // [CHUNK][cmax]
    double* rhs = As + (size_t)GSI_PRED_CHUNK * cmax;               // [cmax]
    double* yv = rhs + cmax;                                        // [CHUNK]
    int* cols = (int*)(yv + GSI_PRED_CHUNK);                        // [cmax]
    int* rowsK = cols + cmax;                                       // [nmax]
    int* wcnt = rowsK + P.nmax;                                     // [nmax/32 + 1]
    __shared__ int sh_lim, sh_kk, sh_c, sh_bad;
    __shared__ double wsum[8];
    if (tid == 0) { sh_lim = k; sh_bad = 0; }
    __syncthreads();
    // ---- lim ----
    const double w_lim = P.w_lim[pair];
    for (int l = tid; l < k; l += T)
        if (lam[l] > w_lim) { atomicMin(&sh_lim, l); break; }
    // ---- K: ordered compaction of the member rows ----
    const bool mok = m < (unsigned)P.w_rows;
    const double* wrow = P.W + (size_t)m * P.w_rows;
    const int nchunks = (n + 31) >> 5;
    for (int ch = warp; ch < nchunks; ch += nwarps) {
        const int j = ch * 32 + lane;
        bool member = false;
        if (j < n && mok) {
            const unsigned mj = (unsigned)P.items[off + j];
            if (mj < (unsigned)P.w_rows) member = (double)__double2float_rn(__ldg(wrow + mj)) > 0.1;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, member);
        if (lane == 0) wcnt[ch] = __popc(bal);
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int ch = 0; ch < nchunks; ++ch) { const int c = wcnt[ch]; wcnt[ch] = run; run += c; }
        sh_kk = run;
        sh_lim = max(sh_lim, 2) < k ? max(sh_lim, 2) : k;
    }
    __syncthreads();
    for (int ch = warp; ch < nchunks; ch += nwarps) {
        const int j = ch * 32 + lane;
        bool member = false;
        if (j < n && mok) {
            const unsigned mj = (unsigned)P.items[off + j];
            if (mj < (unsigned)P.w_rows) member = (double)__double2float_rn(__ldg(wrow + mj)) > 0.1;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, member);
        if (member) rowsK[wcnt[ch] + __popc(bal & ((1u << lane) - 1u))] = j;
    }
    __syncthreads();
    const int kk = sh_kk, lim = sh_lim;
    // ---- mean of the known ratings ----
    double rsum = 0.0;
    for (int r = tid; r < kk; r += T) rsum += P.ratings[off + rowsK[r]];
    rsum = block_sum_256(rsum, wsum);
    // ---- column clean: keep l < lim iff some A(row, l) >= 1e-4 ----
    for (int l = tid; l < cmax; l += T) cols[l] = -1;
    __syncthreads();
    for (int l = tid; l < lim; l += T) {
        bool keep = false;
        for (int r = 0; r < kk && !keep; ++r) keep = U[(size_t)rowsK[r] * k + l] >= 0.0001;
        cols[l] = keep ? 1 : 0;
    }
    __syncthreads();
    if (tid == 0) {                                 // ordered compaction of kept columns (lim <= cmax)
        int c = 0;
        for (int l = 0; l < lim; ++l) if (cols[l] == 1) cols[c++] = l;
        sh_c = c;
    }
    __syncthreads();
    const int c = sh_c;
    const double mean = (kk > 0) ? rsum / (double)kk : 0.0;
    int status = GSI_PRED_OK;
    double pred;
    if (kk == 0) { status = GSI_PRED_EMPTY; pred = __longlong_as_double(0x7ff8000000000000LL); }
    else if (c == 0) { pred = mean; }
    else if (kk < c) { status = GSI_PRED_UNDERDETERMINED; pred = mean; }     // rank(M) <= kk < c: singular by construction,
    else {                                                                   // the stated rule value without forming M
        // ---- Gram M = A^T A (lower), rhs = A^T y ----
        for (int e = tid; e < c * c; e += T) M[e] = 0.0;
        for (int a = tid; a < c; a += T) rhs[a] = 0.0;
        __syncthreads();
        for (int r0 = 0; r0 < kk; r0 += GSI_PRED_CHUNK) {
            const int nr = min(GSI_PRED_CHUNK, kk - r0);
            for (int e = tid; e < nr * c; e += T) {
                const int r = e / c, a = e - r * c;
                As[r * cmax + a] = U[(size_t)rowsK[r0 + r] * k + cols[a]];
            }
            if (tid < nr) yv[tid] = P.ratings[off + rowsK[r0 + tid]] - mean;
            __syncthreads();
            for (int e = tid; e < c * c; e += T) {
                const int a = e / c, b = e - a * c;
                if (b <= a) {
                    double acc = M[e];
                    for (int r = 0; r < nr; ++r) acc = fma(As[r * cmax + a], As[r * cmax + b], acc);
                    M[e] = acc;
                }
            }
            for (int a = tid; a < c; a += T) {
                double acc = rhs[a];
                for (int r = 0; r < nr; ++r) acc = fma(As[r * cmax + a], yv[r], acc);
                rhs[a] = acc;
            }
            __syncthreads();
        }
        // ---- Cholesky M = L L^T in place (lower, row-major M[a*c + b], b <= a) ----
        for (int j = 0; j < c; ++j) {
            const double piv = M[j * c + j];
            if (!(piv > 1e-14)) { if (tid == 0) sh_bad = 1; break; }      // uniform: every thread reads the same pivot
            const double d = sqrt(piv);
            __syncthreads();
            for (int a = j + tid; a < c; a += T) M[a * c + j] = (a == j) ? d : M[a * c + j] / d;
            __syncthreads();
            // trailing update: M[a][b] -= L[a][j] * L[b][j] for j < b <= a
            const int rem = c - j - 1;
            for (int e = tid; e < rem * rem; e += T) {
                const int a = j + 1 + e / rem, b = j + 1 + e % rem;
                if (b <= a) M[a * c + b] -= M[a * c + j] * M[b * c + j];
            }
            __syncthreads();
        }
        __syncthreads();
        if (sh_bad) { if (status == GSI_PRED_OK) status = GSI_PRED_SINGULAR; pred = mean; }
        else {
            // forward L z = rhs, backward L^T x = z
            for (int j = 0; j < c; ++j) {
                if (tid == 0) rhs[j] = rhs[j] / M[j * c + j];
                __syncthreads();
                const double xj = rhs[j];
                for (int a = j + 1 + tid; a < c; a += T) rhs[a] -= M[a * c + j] * xj;
                __syncthreads();
            }
            for (int j = c - 1; j >= 0; --j) {
                if (tid == 0) rhs[j] = rhs[j] / M[j * c + j];
                __syncthreads();
                const double xj = rhs[j];
                for (int a = tid; a < j; a += T) rhs[a] -= M[j * c + a] * xj;
                __syncthreads();
            }
            double s = 0.0;
            for (int a = tid; a < c; a += T) s = fma(U[(size_t)mrow * k + cols[a]], rhs[a], s);
            pred = block_sum_256(s, wsum) + mean;
            if (status != GSI_PRED_OK) pred = mean;      // stated rule for ill-posed pairs
        }
    }
    if (tid == 0) {
        double p = pred;
        if (p > 5.0) p = 5.0;
        if (p < 1.0) p = 1.0;
        const double real = P.ratings[pair];
        const double d = real - p;
        P.err[pair] = __double2float_rn(d * d);
        P.kk[pair] = kk;
        P.pred[pair] = pred;
        P.status[pair] = status;
        P.cols_used[pair] = c;
    }
}

static inline size_t predict_smem_bytes(int cmax, int nmax, bool m_in_smem) {
    size_t d = (m_in_smem ? (size_t)cmax * cmax : 0) + (size_t)GSI_PRED_CHUNK * cmax + cmax + GSI_PRED_CHUNK;
    size_t i = (size_t)cmax + nmax + (nmax / 32 + 2);
    return d * sizeof(double) + i * sizeof(int) + 16;
}
