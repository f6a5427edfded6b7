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
// LU without a singularity check, B3).  A pivot <= 1e-14 or kk < c marks the prediction
// ill-posed: status says so and the rule value (mean of the known ratings) is returned.
//
// Complement-row form.  The pairs of one user share U and differ in the row subset K, and in real item graphs K is almost
// all of the user's rows (ML-1M-shaped fold: kk / n = 0.93 .. 0.98).  The columns of the record are orthonormal over ALL n
// rows, so with R = the rows NOT in K (the target movie itself and the rated movies that are not its neighbours) and
// B = U[R, cols]:
//     M = A^T A = I - B^T B,     M^-1 = I + B^T (I - B B^T)^-1 B                  (Woodbury)
//     rhs = A^T (r_K - mean) = g - mean h - B^T (r_R - mean),   g = U^T r,  h = U^T 1  (per USER, pred_user_sums_kernel)
//     v . M^-1 rhs = v . rhs + (B v) . (I - B B^T)^-1 (B rhs)
// i.e. an |R| x |R| Cholesky and O(|R|^2 c) Gram work instead of c x c and O(kk c^2): 50 x fewer flops on that fold.  The
// same bordered-Gram / Cholesky machinery runs on N = I - B B^T with the border rows -(B rhs)^T and (B v)^T and the
// corner v . rhs.  Used when |R| < c AND the record is orthonormal to working precision (user_exact: records that went
// through the 6-digit text of out_eigen_ are only orthonormal to ~1e-7 and keep the direct form, so that the result
// stays what the reference computes from those digits).
#pragma once
#include "gsi_internal.cuh"
#include "kern_eig_cta.cuh"
#include "ptx.cuh"

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
    int chunk_rows;                 // rows of A staged per Gram step of predict2_kernel (multiple of 4)
    int kmax;                       // largest lim of the launch (capacity of the column list)
    const double* gsum; const double* hsum;   // [sum_off[u] + l]: U^T r and U^T 1 of the user (complement-row form)
    const int64_t* sum_off;         // [users] prefix sums of k
    const uint8_t* user_exact;      // [users] 1: the record's columns are orthonormal to working precision
    int32_t* plan_lim;              // [nnz] lim of the pair (pred_plan_kernel)
    const uint8_t* pair_mask_dev;   // optional [nnz]
};

// Per user: g = U^T r, h = U^T 1 (k-vectors, fixed summation order) and whether the columns are unit vectors to working
// precision (the proxy for "this record did not go through 6-digit text").  grid = users, block 256.
__global__ void __launch_bounds__(256) pred_user_sums_kernel(PredParams P, double* __restrict__ gsum, double* __restrict__ hsum,
                                                             uint8_t* __restrict__ user_exact, int n_users) {
    const int u = blockIdx.x;
    if (u >= n_users) return;
    const int64_t off = P.offsets[u];
    const int n = (int)(P.offsets[u + 1] - off), k = P.k[u];
    const double* U = P.vec + P.vec_off[u];
    const double* r = P.ratings + off;
    __shared__ int bad;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    for (int l = threadIdx.x; l < k; l += 256) {
        double g = 0.0, h = 0.0, q = 0.0;
        for (int i = 0; i < n; ++i) {
            const double x = U[(size_t)i * k + l];
            g = fma(x, r[i], g); h += x; q = fma(x, x, q);
        }
        gsum[P.sum_off[u] + l] = g;
        hsum[P.sum_off[u] + l] = h;
        if (!(fabs(q - 1.0) <= 1e-12)) bad = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) user_exact[u] = bad ? 0 : 1;
}

// Per pair: kk = #K and lim (what the host needs to size the pair's solve).  grid = users, block 256 (warp per pair, round robin).
__global__ void __launch_bounds__(256) pred_plan_kernel(PredParams P, int n_users) {
    const int u = blockIdx.x;
    if (u >= n_users) return;
    const int64_t off = P.offsets[u];
    const int n = (int)(P.offsets[u + 1] - off), k = P.k[u];
    const double* lam = P.lam + P.lam_off[u];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int mrow = warp; mrow < n; mrow += 8) {
        const int64_t pair = off + mrow;
        if (P.pair_mask_dev && !P.pair_mask_dev[pair]) continue;
        const unsigned m = (unsigned)P.items[pair];
        int cnt = 0;
        if (m < (unsigned)P.w_rows) {
            const double* wrow = P.W + (size_t)m * P.w_rows;
            for (int j = lane; j < n; j += 32) {
                const unsigned mj = (unsigned)P.items[off + j];
                if (mj < (unsigned)P.w_rows && (double)__double2float_rn(__ldg(wrow + mj)) > 0.1) ++cnt;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        const double w_lim = P.w_lim[pair];
        int lo = 0, hi = k;                               // first l with lam[l] > w_lim (lam ascending)
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (lam[mid] > w_lim) hi = mid; else lo = mid + 1; }
        if (lane == 0) { P.kk[pair] = cnt; P.plan_lim[pair] = min(max(lo, 2), k); }
    }
}

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

// ---------------------------------------------------------------------------------------------------
// The kernel.  How the normal equations are solved:
//
//   * A = U[K, cols] is staged 32 rows at a time and  M = A^T A  is accumulated with FP64 MMA (m8n8k4) into a
//     tile-packed lower triangle of 8 x 8 tiles in shared memory (padded columns: zero, padded diagonal: 1);
//   * the matrix is bordered by one more tile row that holds  rhs^T = (A^T (r_K - mean))^T  (it falls out of the
//     same MMA pass: r_K - mean rides along as column `cpad` of the staging buffer) and  v^T = U[m, cols];
//   * a right-looking blocked Cholesky (8 wide: diagonal tile by one warp in registers, panel by one thread
//     per row, trailing update by MMA) runs over the bordered matrix.  The border rows turn into
//     (L^-1 rhs)^T and (L^-1 v)^T and the corner tile receives  -(L^-1 v).(L^-1 rhs) = -v^T M^-1 rhs,
//     i.e. the prediction minus the mean -- no triangular solves (local_calc_precomp.cpp:308-315).
#define P2_R 32
__host__ __device__ __forceinline__ int p2_ld(int cpad) { return ((cpad + 8 + 15) / 16) * 16 + 4; }    // == 4 mod 16: conflict-free fragments
__host__ __device__ __forceinline__ int p2_tile(int ta, int tb) { return ((ta * (ta + 1) / 2) + tb) << 6; }
static inline size_t predict2_tiles_dbl(int cmax) { const int nt = (cmax + 7) / 8; return (size_t)(nt + 1) * (nt + 2) / 2 * 64; }
// M in shared memory when it fits; otherwise it lives in a per-CTA scratch in global memory (L2 resident) and only the
// staging buffer is on chip
// dim = order of the matrix that is factored (c, or |R| in the complement-row form); kmax = capacity of the column list
static inline size_t predict2_smem_bytes(int cmax, int nmax, bool m_in_smem, int chunk_rows = P2_R, int kmax = -1) {
    if (kmax < 0) kmax = cmax;
    const int nt = (cmax + 7) / 8;
    const size_t d = (m_in_smem ? predict2_tiles_dbl(cmax) : 0) + (size_t)chunk_rows * p2_ld(nt * 8) + (size_t)(nt * 8) + 8 * 32;
    const size_t i = (size_t)kmax + nmax + (nmax / 32 + 2);
    return d * sizeof(double) + i * sizeof(int) + 16;
}

__global__ void __launch_bounds__(256) predict2_kernel(PredParams P) {
    extern __shared__ double sm[];
    const int tid = threadIdx.x, T = 256, lane = tid & 31, warp = tid >> 5, nwarps = 8;
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
    const int cmax = P.cmax, ntmax = (cmax + 7) >> 3;
    const int R = P.chunk_rows;
    double* M = P.m_in_smem ? sm : P.work + (size_t)blockIdx.x * P.work_stride;     // (ntmax+1)(ntmax+2)/2 tiles of 64
    double* As = sm + (P.m_in_smem ? (size_t)(ntmax + 1) * (ntmax + 2) / 2 * 64 : 0);   // [R][ld]
    double* rt = As + (size_t)R * p2_ld(ntmax * 8);                  // [ntmax*8] r_j - mean of the complement rows
    double* rpart = rt + ntmax * 8;                                  // [8][32] partial dot products of the complement staging
    int* cols = (int*)(rpart + 8 * 32);                              // [kmax]
    int* rowsK = cols + P.kmax;                                      // [nmax] K rows from the front, complement rows from the back
    int* wcnt = rowsK + P.nmax;                                      // [nmax/32 + 1]
    __shared__ int sh_lim, sh_kk, sh_c, sh_bad;
    __shared__ double wsum[8];
    if (tid == 0) { sh_lim = k; sh_bad = 0; }
    __syncthreads();
    // ---- lim, K, mean, column clean ----
    const double w_lim = P.w_lim[pair];
    for (int l = tid; l < k; l += T)
        if (lam[l] > w_lim) { atomicMin(&sh_lim, l); break; }
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
        // the membership bit is recomputed below (no per-row scratch)
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int ch = 0; ch < nchunks; ++ch) { const int cc = wcnt[ch]; wcnt[ch] = run; run += cc; }
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
        const int before = wcnt[ch] + __popc(bal & ((1u << lane) - 1u));          // members before row j
        if (member) rowsK[before] = j;
        else if (j < n) rowsK[n - 1 - (j - before)] = j;                           // complement rows, ascending from the back
    }
    __syncthreads();
    const int kk = sh_kk, lim = sh_lim, nr = n - kk;
    const int* rowsR = rowsK + (n - 1);                                            // rowsR[-q] = q-th complement row
    double rsum = 0.0;
    for (int r = tid; r < kk; r += T) rsum += P.ratings[off + rowsK[r]];
    rsum = block_sum_256(rsum, wsum);
    for (int l = tid; l < P.kmax; l += T) cols[l] = -1;
    __syncthreads();
    for (int l = tid; l < lim; l += T) {
        bool keep = false;
        for (int r = 0; r < kk && !keep; ++r) keep = U[(size_t)rowsK[r] * k + l] >= 0.0001;
        cols[l] = keep ? 1 : 0;
    }
    __syncthreads();
    if (tid == 0) {
        int cc = 0;
        for (int l = 0; l < lim; ++l) if (cols[l] == 1) cols[cc++] = l;
        sh_c = cc;
    }
    __syncthreads();
    const int c = sh_c;
    const double mean = (kk > 0) ? rsum / (double)kk : 0.0;
    int status = GSI_PRED_OK;
    double pred;
    if (kk == 0) { status = GSI_PRED_EMPTY; pred = __longlong_as_double(0x7ff8000000000000LL); }
    else if (c == 0) { pred = mean; }
    else if (kk < c) { status = GSI_PRED_UNDERDETERMINED; pred = mean; }
    else {
        // complement-row form when it is the smaller problem and the record is orthonormal to working precision
        const bool wood = nr < c && P.user_exact != nullptr && P.user_exact[u] != 0;
        const int dim = wood ? nr : c;
        const int nt = (dim + 7) >> 3, cpad = nt * 8, ld = p2_ld(cpad), wcols = cpad + 8;
        const int fr = lane >> 2, fk = lane & 3;                      // MMA fragment row / k index of this lane
        for (int e = tid; e < ((nt + 1) * (nt + 2) / 2) * 64; e += T) M[e] = 0.0;
        if (wood) for (int q = tid; q < cpad; q += T) rt[q] = (q < nr) ? P.ratings[off + rowsR[-q]] - mean : 0.0;
        const double* gs = wood ? P.gsum + P.sum_off[u] : nullptr;
        const double* hs = wood ? P.hsum + P.sum_off[u] : nullptr;
        const int nsteps = wood ? c : kk;                             // staged rows: columns of B^T, or rows of A
        // ---- bordered Gram: tiles (ta, tb), tb <= ta <= nt; row `nt` is the border ----
        for (int r0 = 0; r0 < nsteps; r0 += R) {
            __syncthreads();                                          // the previous chunk is consumed (first pass: M is zero)
            if (wood) {
                // staged row rr <-> column l = r0 + rr of the record: [ B[q][l], q < nr | rhs_l | v_l ].  A warp covers 8 columns x 4
                // complement rows per step (coalesced 64-byte runs of a U row, two-way conflicts on the shared stores at most).
                for (int rb = 0; rb < R; rb += 32) {
                    const int rr = rb + (warp & 3) * 8 + (lane & 7), l = r0 + rr;
                    const int jres = (lane >> 3) + 4 * (warp >> 2);   // this thread's residue of q mod 8
                    const bool lok = rr < R && l < c;
                    const int cl = lok ? cols[l] : 0;
                    double part = 0.0;
                    if (rr < R) {
                        for (int q = jres; q < cpad; q += 8) {
                            const double val = (lok && q < nr) ? U[(size_t)rowsR[-q] * k + cl] : 0.0;
                            As[rr * ld + q] = val;
                            part = fma(val, rt[q], part);
                        }
                    }
                    rpart[jres * 32 + (rr & 31)] = part;
                    __syncthreads();
                    if (tid < 32 && rb + tid < R) {
                        const int r2 = rb + tid, l2 = r0 + r2;
                        double rhs = 0.0, vl = 0.0;
                        if (l2 < c) {
                            const int c2 = cols[l2];
                            double sub = 0.0;
#pragma unroll
                            for (int q8 = 0; q8 < 8; ++q8) sub += rpart[q8 * 32 + tid];
                            rhs = (gs[c2] - mean * hs[c2]) - sub;
                            vl = U[(size_t)mrow * k + c2];
                        }
                        double* brow = As + r2 * ld + cpad;
                        brow[0] = rhs; brow[1] = vl;
#pragma unroll
                        for (int q8 = 2; q8 < 8; ++q8) brow[q8] = 0.0;
                    }
                    __syncthreads();
                }
            } else
            // a warp stages 4 rows; 8 independent gathers in flight per lane (the loads are L2-latency bound)
            {
                const double* urow[4];
                double yv4[4];
                bool live[4];
                for (int rg = warp * 4; rg < R; rg += 32) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int r = rg + q;
                    live[q] = r0 + r < kk;
                    const int row = live[q] ? rowsK[r0 + r] : 0;
                    urow[q] = U + (size_t)row * k;
                    yv4[q] = live[q] ? P.ratings[off + row] - mean : 0.0;
                }
                for (int a0 = 0; a0 < wcols; a0 += 64) {
                    double v[4][2];
                    int ca[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) { const int a = a0 + lane + 32 * i; ca[i] = (a < c) ? cols[a] : -1; }
#pragma unroll
                    for (int q = 0; q < 4; ++q)
#pragma unroll
                        for (int i = 0; i < 2; ++i) v[q][i] = (live[q] && ca[i] >= 0) ? urow[q][ca[i]] : 0.0;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const int a = a0 + lane + 32 * i;
                            if (a < wcols) As[(rg + q) * ld + a] = (a == cpad) ? yv4[q] : v[q][i];
                        }
                }
                }
            }
            __syncthreads();
            int seg = 0;
            for (int ta = 0; ta <= nt; ++ta)
                for (int s0 = 0; s0 <= ta; s0 += 4, ++seg) {
                    if ((seg & 7) != warp) continue;
                    const int ns = min(4, ta - s0 + 1);
                    double2 acc[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        acc[i] = (i < ns) ? *(const double2*)(M + p2_tile(ta, s0 + i) + fr * 8 + 2 * fk) : make_double2(0.0, 0.0);
                    const double* ap = As + fk * ld + 8 * ta + fr;
                    const double* bp = As + fk * ld + 8 * s0 + fr;
#pragma unroll 2
                    for (int rr = 0; rr < R; rr += 4) {
                        const double a = ap[rr * ld];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (i < ns) dmma(acc[i].x, acc[i].y, a, bp[rr * ld + 8 * i]);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (i < ns) *(double2*)(M + p2_tile(ta, s0 + i) + fr * 8 + 2 * fk) = acc[i];
                }
        }
        __syncthreads();
        if (wood) {
            // G = B B^T (+ borders) -> N = I - G; border row 0 = -(B rhs)^T; row 1 = (B v)^T and the corner v . rhs stay
            const int ntiles = (nt + 1) * (nt + 2) / 2;
            for (int e = tid; e < ntiles * 64; e += T) {
                const int tile = e >> 6, i = (e >> 3) & 7, j = e & 7;
                int ta = (int)((sqrt(8.0 * (double)tile + 1.0) - 1.0) * 0.5);
                while ((ta + 1) * (ta + 2) / 2 <= tile) ++ta;
                while (ta * (ta + 1) / 2 > tile) --ta;
                const int tb = tile - ta * (ta + 1) / 2;
                if (ta < nt) M[e] = ((ta == tb && i == j) ? 1.0 : 0.0) - M[e];
                else if (i == 0) M[e] = -M[e];
            }
        } else
        {
        // border row 1 = v^T; padded diagonal = 1
        for (int a = tid; a < cpad; a += T) {
            const int tb = a >> 3, j = a & 7;
            M[p2_tile(nt, tb) + 8 + j] = (a < c) ? U[(size_t)mrow * k + cols[a]] : 0.0;
            if (a >= c) M[p2_tile(tb, tb) + j * 8 + j] = 1.0;
        }
        }
        __syncthreads();
        // ---- blocked Cholesky over the bordered matrix ----
        for (int jb = 0; jb < nt; ++jb) {
            double* D = M + p2_tile(jb, jb);
            if (warp == 0) {                                          // diagonal tile: lane i (mod 8) holds row i
                const int i = lane & 7;
                double r[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) r[q] = D[i * 8 + q];
                bool ok = true;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const double piv = __shfl_sync(0xffffffffu, r[j], j);
                    if (!(piv > 1e-14)) { ok = false; break; }        // uniform over the warp
                    const double d = sqrt(piv);
                    r[j] = (i == j) ? d : r[j] / d;
#pragma unroll
                    for (int q = j + 1; q < 8; ++q) {
                        const double lq = __shfl_sync(0xffffffffu, r[j], q);          // L[q][j]
                        r[q] = fma(-r[j], lq, r[q]);
                    }
                }
                if (!ok) { if (lane == 0) sh_bad = 1; }
                else if (lane < 8) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) D[i * 8 + q] = (q <= i) ? r[q] : 0.0;
                }
            }
            __syncthreads();
            if (sh_bad) break;
            // panel: rows of the tiles (ta, jb), ta = jb+1 .. nt (border included): x L_jj^T = row
            const int nrows = (nt - jb) * 8;
            for (int rr = tid; rr < nrows; rr += T) {
                double* X = M + p2_tile(jb + 1 + (rr >> 3), jb) + (rr & 7) * 8;
                double x[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) x[q] = X[q];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    double sacc = x[j];
#pragma unroll
                    for (int q = 0; q < j; ++q) sacc = fma(-x[q], D[j * 8 + q], sacc);
                    x[j] = sacc / D[j * 8 + j];
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) X[q] = x[q];
            }
            __syncthreads();
            // trailing update: tile (ta, tb) -= L(ta, jb) L(tb, jb)^T for jb < tb <= ta <= nt
            int seg = 0;
            for (int ta = jb + 1; ta <= nt; ++ta)
                for (int s0 = jb + 1; s0 <= ta; s0 += 4, ++seg) {
                    if ((seg & 7) != warp) continue;
                    const int ns = min(4, ta - s0 + 1);
                    const double* LA = M + p2_tile(ta, jb) + fr * 8 + fk;
                    const double a0 = -LA[0], a1 = -LA[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (i < ns) {
                            const double* LB = M + p2_tile(s0 + i, jb) + fr * 8 + fk;
                            double2* C = (double2*)(M + p2_tile(ta, s0 + i) + fr * 8 + 2 * fk);
                            double2 cv = *C;
                            dmma(cv.x, cv.y, a0, LB[0]);
                            dmma(cv.x, cv.y, a1, LB[4]);
                            *C = cv;
                        }
                }
            __syncthreads();
        }
        __syncthreads();
        if (sh_bad) { status = GSI_PRED_SINGULAR; pred = mean; }
        else pred = wood ? mean + M[p2_tile(nt, nt) + 8] : mean - M[p2_tile(nt, nt) + 8];
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
