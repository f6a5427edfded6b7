// Large path (n > 160): the user's matrix lives in HBM/L2, all SMs cooperate on a chunk of users.
//
//   lap_* kernels      gather W -> degree -> normalised Laplacian -> sig_min -> sym(lower)+I
//                      (precompute_local.cpp:185-249), column-major G[i + j*ld]
//   bj_* kernels       one-sided BLOCK Jacobi on G: column blocks of B = M/2; each round pairs the
//                      blocks with the circle ordering and every pair (I,J) does
//                          H = P^T P            P = [G_I G_J]   (n x M)      bj_gram   (DMMA)
//                          Q = one sweep of two-sided Jacobi rotations on H   bj_inner
//                          P <- P Q                                           bj_update (DMMA)
//                      FP64 tensor-core mma.sync m8n8k4 carries both GEMM-shaped steps.
//   fin_* kernels      column norms -> eigenvalues, ascending rank, cutoff (:252-261), emit
//
// All users of a chunk share nb (block count, even) so that every user finishes a sweep on the
// same round; blocks beyond a user's n are phantom and their pairs exit at once.
#pragma once
#include "gsi_internal.cuh"
#include "kern_eig_cta.cuh"

struct LChunk {
    int nu;        // users in the chunk
    int nb;        // column blocks per user (even), uniform over the chunk
    int ncols;     // nb * B
    int splits;    // row splits of a panel (uniform upper bound)
    const int32_t* n;          // [nu]
    const int32_t* ld;         // [nu] leading dimension, multiple of 8, rows >= n are zero
    const int64_t* g_off;      // [nu] offset of the user's G (ld x ncols doubles)
    const int64_t* item_off;   // [nu] offset into items / sig_min
    const int64_t* row_off;    // [nu] offset into per-row scratch (deg/scale) and per-column scratch
    const int64_t* vec_off;    // [nu] offset into vec_pad
    const int64_t* lam_off;    // [nu] offset into lam_pad
    double* G;
    double* deg;               // per-row scratch
    double* scale;             // per-row scratch
    double* colnorm;           // [nu * ncols] signed norms
    double* colval;            // [nu * ncols] eigenvalues
    int32_t* perm;             // [nu * ncols]
    unsigned int* sigmax;      // [nu] float bits
    int32_t* done;             // [nu]
    unsigned long long* smax;  // [nu]
    int32_t* sweeps;           // [nu]
    int32_t* k;                // [nu]
    int32_t* remaining;        // [1]
    double* Hpart;             // [nu][nb/2][splits][M*M]
    double* Q;                 // [nu][nb/2][M*M]
    int tiled;                 // 1: G is tile-major (64x64 tiles, ld = 64-padded n) -- the Householder path
};

// offset of element (i, j) of a user's matrix
__device__ __forceinline__ size_t lap_idx(const LChunk& C, int ld, int i, int j) {
    if (C.tiled) return (((size_t)(j >> 6) * (ld >> 6) + (i >> 6)) << 12) + ((j & 63) << 6) + (i & 63);
    return (size_t)i + (size_t)j * ld;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------
// Laplacian stage
// ------------------------------------------------------------------------------------------

// grid (tiles_x * tiles_y, 1, nu), block (32, 8): 32x32 tile through smem so that both the table
// read (along a W row) and the G write (along a G column) are contiguous.
__global__ void lap_gather_kernel(LChunk C, const double* __restrict__ W, int w_rows,
                                  const int32_t* __restrict__ items, int tiles_per_dim) {
    __shared__ double tile[32][33];
    __shared__ int idi[32], idj[32];
    const int u = blockIdx.z;
    const int n = C.n[u];
    const int ti = blockIdx.x / tiles_per_dim, tj = blockIdx.x % tiles_per_dim;
    if (ti * 32 >= n || tj * 32 >= n) return;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int64_t ioff = C.item_off[u];
    if (ty == 0) idi[tx] = (ti * 32 + tx < n) ? items[ioff + ti * 32 + tx] : -1;
    if (ty == 1) idj[tx] = (tj * 32 + tx < n) ? items[ioff + tj * 32 + tx] : -1;
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        const int il = ty + 8 * rr;
        const unsigned mi = (unsigned)idi[il], mj = (unsigned)idj[tx];
        double v = 0.0;
        if (mi < (unsigned)w_rows && mj < (unsigned)w_rows) v = __ldg(W + (size_t)mi * w_rows + mj);
        tile[il][tx] = v;
    }
    __syncthreads();
    double* G = C.G + C.g_off[u];
    const int ld = C.ld[u];
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        const int jl = ty + 8 * rr;
        const int i = ti * 32 + tx, j = tj * 32 + jl;
        if (i < n && j < n) G[lap_idx(C, ld, i, j)] = tile[tx][jl];
    }
}

// grid (ceil(nmax/128), 1, nu), block 128: thread per row, sequential j order (reference :199-203)
__global__ void lap_degree_kernel(LChunk C) {
    const int u = blockIdx.z;
    const int n = C.n[u];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* G = C.G + C.g_off[u];
    const int ld = C.ld[u];
    double d = 0.0;
    for (int j = 0; j < n; ++j) d = __dadd_rn(d, G[lap_idx(C, ld, i, j)]);
    if (d == 0.0) d = 1.0;
    C.deg[C.row_off[u] + i] = d;
    C.scale[C.row_off[u] + i] = __dsqrt_rn(__ddiv_rn(1.0, d));
}

// grid (ceil(nmax/128), ceil(nmax/8), nu), block 128: 128 rows x 8 columns per CTA
__global__ void lap_transform_kernel(LChunk C) {
    const int u = blockIdx.z;
    const int n = C.n[u];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double* G = C.G + C.g_off[u];
    const int ld = C.ld[u];
    const double* deg = C.deg + C.row_off[u];
    const double* sc = C.scale + C.row_off[u];
    const double si = sc[i], di = deg[i];
    const int j0 = blockIdx.y * 8;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
        const int j = j0 + jj;
        if (j < n) {
            const double w = G[lap_idx(C, ld, i, j)];
            const double ll = (i == j) ? __dsub_rn(di, w) : __dsub_rn(0.0, w);
            G[lap_idx(C, ld, i, j)] = __dmul_rn(__dmul_rn(si, ll), sc[j]);
        }
    }
}

// thread per row: float accumulator over the full row, j ascending (reference :236-249)
__global__ void lap_sigmin_kernel(LChunk C, double* __restrict__ sig_min) {
    const int u = blockIdx.z;
    const int n = C.n[u];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* G = C.G + C.g_off[u];
    const int ld = C.ld[u];
    float acc = 0.f;
    for (int j = 0; j < n; ++j) {
        const double x = G[lap_idx(C, ld, i, j)];
        acc = __double2float_rn(__dadd_rn((double)acc, __dmul_rn(x, x)));
    }
    const float sig = __fsqrt_rn(acc);
    sig_min[C.item_off[u] + i] = __dadd_rn((double)sig, 0.01);
    atomicMax(&C.sigmax[u], __float_as_uint(sig));
}

// upper <- lower through a 32x32 smem tile (only tiles with ti <= tj do work), diagonal += shift.
// grid (tiles*tiles, 1, nu), block (32, 8)
__global__ void lap_symmetrize_kernel(LChunk C, int tiles_per_dim, double shift) {
    __shared__ double tile[32][33];
    const int u = blockIdx.z;
    const int n = C.n[u];
    const int ti = blockIdx.x / tiles_per_dim, tj = blockIdx.x % tiles_per_dim;
    if (ti > tj || tj * 32 >= n) return;
    double* G = C.G + C.g_off[u];
    const int ld = C.ld[u];
    const int tx = threadIdx.x, ty = threadIdx.y;
    // read the lower tile (rows of block tj, columns of block ti): contiguous along rows
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        const int cl = ty + 8 * rr;
        const int i = tj * 32 + tx, j = ti * 32 + cl;
        tile[cl][tx] = (i < n && j < n) ? G[lap_idx(C, ld, i, j)] : 0.0;
    }
    __syncthreads();
    // write the upper tile (rows of block ti, columns of block tj)
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        const int cl = ty + 8 * rr;
        const int i = ti * 32 + tx, j = tj * 32 + cl;
        if (i < n && j < n) {
            if (i < j) G[lap_idx(C, ld, i, j)] = tile[tx][cl];
            else if (i == j) G[lap_idx(C, ld, i, j)] += shift;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Fused Laplacian stage of the Householder path (tile-major G): two passes over a user's matrix instead of five.
// A CTA owns a block of 64 rows and walks the 64 x 64 tiles of that block row in column order, so the row sums keep
// the reference's sequential j order (degree :199-203, sig_min :236-249) -- bit-exact with the kernels above.
//   pass 1  gather W (:185-192) tile by tile through shared memory (table read along a W row, G write along a
//           tile column, both contiguous), row sums -> deg, scale
//   pass 2  ll2 = (s_i * ll_ij) * s_j (:211-222) on every tile of the block row, float-accumulated row norms ->
//           sig_min, sigmax; written back: the tiles on and below the diagonal only (the diagonal tile with its
//           upper triangle mirrored from the lower one) -- the Householder path never reads a tile above the
//           diagonal, those keep the gathered weights
// grid (ceil(nmax/64), 1, nu), block 256
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lap_fused_gather_kernel(LChunk C, const double* __restrict__ W, int w_rows,
                                                               const int32_t* __restrict__ items) {
    __shared__ double t[64][65];
    __shared__ int idi[64], idj[64];
    const int u = blockIdx.z;
    const int n = C.n[u];
    const int I = blockIdx.x, i0 = I * 64;
    if (i0 >= n) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t ioff = C.item_off[u];
    double* G = C.G + C.g_off[u];
    const int NT = C.ld[u] >> 6, NTn = (n + 63) >> 6;
    if (tid < 64) idi[tid] = (i0 + tid < n) ? items[ioff + i0 + tid] : -1;
    double d = 0.0;                                           // row sum of row i0 + tid (threads 0..63)
    for (int J = 0; J < NTn; ++J) {
        __syncthreads();                                      // the previous tile is consumed (first pass: idi is written)
        if (tid < 64) idj[tid] = (J * 64 + tid < n) ? items[ioff + J * 64 + tid] : -1;
        __syncthreads();
        double v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {                        // warp w: rows 8w..8w+7, two 32-column halves each
            const int r = 8 * warp + (q >> 1), c = 32 * (q & 1) + lane;
            const unsigned mi = (unsigned)idi[r], mj = (unsigned)idj[c];
            v[q] = (mi < (unsigned)w_rows && mj < (unsigned)w_rows) ? __ldg(W + (size_t)mi * w_rows + mj) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) t[8 * warp + (q >> 1)][32 * (q & 1) + lane] = v[q];
        __syncthreads();
        double* tile = G + (((size_t)J * NT + I) << 12);
#pragma unroll
        for (int q = 0; q < 16; ++q) {                        // tile-major write: lanes along the rows of a tile column
            const int e = tid + 256 * q, r = e & 63, c = e >> 6;
            if (i0 + r < n && J * 64 + c < n) tile[(c << 6) + r] = t[r][c];
        }
        if (tid < 64 && i0 + tid < n) {
            const int cn = min(64, n - J * 64);
            for (int c = 0; c < cn; ++c) d = __dadd_rn(d, t[tid][c]);
        }
    }
    if (tid < 64 && i0 + tid < n) {
        if (d == 0.0) d = 1.0;
        C.deg[C.row_off[u] + i0 + tid] = d;
        C.scale[C.row_off[u] + i0 + tid] = __dsqrt_rn(__ddiv_rn(1.0, d));
    }
}

__global__ void __launch_bounds__(256) lap_fused_transform_kernel(LChunk C, double* __restrict__ sig_min) {
    __shared__ double t[64][65];
    __shared__ double srow[64], drow[64], scol[64];
    const int u = blockIdx.z;
    const int n = C.n[u];
    const int I = blockIdx.x, i0 = I * 64;
    if (i0 >= n) return;
    const int tid = threadIdx.x;
    double* G = C.G + C.g_off[u];
    const int NT = C.ld[u] >> 6, NTn = (n + 63) >> 6;
    const double* deg = C.deg + C.row_off[u];
    const double* sc = C.scale + C.row_off[u];
    if (tid < 64) { srow[tid] = (i0 + tid < n) ? sc[i0 + tid] : 0.0; drow[tid] = (i0 + tid < n) ? deg[i0 + tid] : 0.0; }
    float acc = 0.f;
    for (int J = 0; J < NTn; ++J) {
        __syncthreads();
        if (tid < 64) scol[tid] = (J * 64 + tid < n) ? sc[J * 64 + tid] : 0.0;
        double* tile = G + (((size_t)J * NT + I) << 12);
        double w[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) { const int e = tid + 256 * q; w[q] = tile[e]; }     // (c << 6) + r == e
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int e = tid + 256 * q, r = e & 63, c = e >> 6;
            const int i = i0 + r, j = J * 64 + c;
            const double ll = (i == j) ? __dsub_rn(drow[r], w[q]) : __dsub_rn(0.0, w[q]);
            t[r][c] = (i < n && j < n) ? __dmul_rn(__dmul_rn(srow[r], ll), scol[c]) : 0.0;
        }
        __syncthreads();
        if (J <= I) {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int e = tid + 256 * q, r = e & 63, c = e >> 6;
                if (i0 + r < n && J * 64 + c < n) tile[e] = (J < I || r >= c) ? t[r][c] : t[c][r];
            }
        }
        if (tid < 64 && i0 + tid < n) {
            const int cn = min(64, n - J * 64);
            for (int c = 0; c < cn; ++c) {
                const double x = t[tid][c];
                acc = __double2float_rn(__dadd_rn((double)acc, __dmul_rn(x, x)));
            }
        }
    }
    if (tid < 64 && i0 + tid < n) {
        const float sig = __fsqrt_rn(acc);
        sig_min[C.item_off[u] + i0 + tid] = __dadd_rn((double)sig, 0.01);
        atomicMax(&C.sigmax[u], __float_as_uint(sig));
    }
}

// ------------------------------------------------------------------------------------------
// Block Jacobi.  Templated on the panel width M (two column blocks of B = M/2).
//   round >= 0 : circle ordering over the nb blocks, inner sweep over the B*B CROSS pairs only
//   round <  0 : "diagonal" round, pairs (2s, 2s+1), full cyclic sweep over all M(M-1)/2 pairs --
//                once per sweep, so that every column pair of the matrix is rotated once per sweep
// ------------------------------------------------------------------------------------------

template <int M>
__device__ __forceinline__ bool bj_task(const LChunk& C, int u, int round, int slot, int& I, int& J) {
    if (C.done[u]) return false;
    if (round < 0) { I = 2 * slot; J = 2 * slot + 1; }
    else circle_pair(C.nb, round, slot, I, J);
    // I < J.  Circle rounds: nothing to do when block J is phantom (all-zero columns).  Diagonal
    // round: block I still has to be orthogonalised within itself.
    return (round < 0 ? I : J) * (M / 2) < C.n[u];
}

template <int M>
__device__ __forceinline__ int panel_col(int cc, int I, int J) {
    return (cc < M / 2) ? I * (M / 2) + cc : J * (M / 2) + (cc - M / 2);
}

// grid (nb/2, splits, nu), block 128.  Partial Gram H = P^T P of the M-column panel over this
// CTA's rows; DMMA fragments come straight from global/L2 (each element is loaded exactly once per
// warp), next k-step prefetched while the current one is multiplied.
template <int M>
__global__ void __launch_bounds__(128) bj_gram_kernel(LChunk C, int round) {
    constexpr int T = M / 8, NT = T * (T + 1) / 2;
    const int u = blockIdx.z, slot = blockIdx.x, split = blockIdx.y;
    int I, J;
    if (!bj_task<M>(C, u, round, slot, I, J)) return;
    const int ld = C.ld[u];
    const int r_begin = split * GSI_BJ_ROWS;
    if (r_begin >= ld) return;
    const int r_end = min(ld, r_begin + GSI_BJ_ROWS);
    const double* G = C.G + C.g_off[u];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lr = lane & 3, lc = lane >> 2;
    const double* colp[T];
#pragma unroll
    for (int t = 0; t < T; ++t) colp[t] = G + (size_t)panel_col<M>(8 * t + lc, I, J) * ld + lr;
    double acc[NT][2];
#pragma unroll
    for (int t = 0; t < NT; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
    int r = r_begin + 4 * warp;
    double f[T], g[T];
    if (r < r_end) {
#pragma unroll
        for (int t = 0; t < T; ++t) f[t] = colp[t][r];
    }
    for (; r < r_end; r += 16) {
        const int rn = r + 16;
        if (rn < r_end) {
#pragma unroll
            for (int t = 0; t < T; ++t) g[t] = colp[t][rn];
        }
        int idx = 0;
#pragma unroll
        for (int ti = 0; ti < T; ++ti)
#pragma unroll
            for (int tj = ti; tj < T; ++tj) { dmma884(acc[idx][0], acc[idx][1], f[ti], f[tj]); ++idx; }
#pragma unroll
        for (int t = 0; t < T; ++t) f[t] = g[t];
    }
    // deterministic reduction over the 4 warps: warp w adds its tiles in turn
    __shared__ double Hs[M * M];
    for (int w = 0; w < 4; ++w) {
        if (warp == w) {
            int idx = 0;
#pragma unroll
            for (int ti = 0; ti < T; ++ti)
#pragma unroll
                for (int tj = ti; tj < T; ++tj) {
                    const int a = 8 * ti + lc, b = 8 * tj + 2 * lr;
                    if (w == 0) { Hs[a * M + b] = acc[idx][0]; Hs[a * M + b + 1] = acc[idx][1]; }
                    else { Hs[a * M + b] += acc[idx][0]; Hs[a * M + b + 1] += acc[idx][1]; }
                    ++idx;
                }
        }
        __syncthreads();
    }
    double* H = C.Hpart + (((size_t)u * (C.nb >> 1) + slot) * C.splits + split) * (M * M);
    for (int e = threadIdx.x; e < M * M; e += 128) {
        const int a = e / M, b = e % M;
        if ((a >> 3) <= (b >> 3)) H[e] = Hs[e];
    }
}

// grid (nb/2, 1, nu), block 8*M.  H = sum of partials (upper tiles mirrored); one sweep of
// two-sided Jacobi on the M x M panel Gram, rotations accumulated into Q.
template <int M>
__global__ void __launch_bounds__(8 * M) bj_inner_kernel(LChunk C, int round) {
    constexpr int B = M / 2, LDH = M + 1, NTHR = 8 * M;
    const int u = blockIdx.z, slot = blockIdx.x;
    int I, J;
    if (!bj_task<M>(C, u, round, slot, I, J)) return;
    extern __shared__ double sm_inner[];
    double* H = sm_inner;                 // [M][LDH]
    double* Q = H + M * LDH;              // [M][LDH]
    __shared__ double cs[B][2];
    __shared__ int pq[B][2];
    __shared__ unsigned long long sh_max;
    const int tid = threadIdx.x;
    const int ld = C.ld[u];
    const int nsplit = (ld + GSI_BJ_ROWS - 1) / GSI_BJ_ROWS;
    const double* Hp = C.Hpart + ((size_t)u * (C.nb >> 1) + slot) * C.splits * (M * M);
    for (int e = tid; e < M * M; e += NTHR) {
        const int a = e / M, b = e % M;
        if ((a >> 3) <= (b >> 3)) {
            double v = 0.0;
            for (int s = 0; s < nsplit; ++s) v += Hp[(size_t)s * (M * M) + e];
            H[a * LDH + b] = v;
            if ((a >> 3) < (b >> 3)) H[b * LDH + a] = v;
        }
        Q[a * LDH + b] = (a == b) ? 1.0 : 0.0;
    }
    if (tid == 0) sh_max = 0ull;
    __syncthreads();
    const bool full = round < 0;
    const int nrounds = full ? M - 1 : B;
    for (int r = 0; r < nrounds; ++r) {
        if (tid < B) {
            int p, q;
            if (full) circle_pair(M, r, tid, p, q);
            else { p = tid; q = B + ((tid + r) & (B - 1)); }
            const double a = H[p * LDH + p], b = H[q * LDH + q], g = H[p * LDH + q];
            double c = 1.0, s = 0.0, t;
            const double ab = a * b, g2 = g * g;
            if (ab > 0.0) {
                if (g2 > GSI_STOP2 * ab) atomicMax(&sh_max, dbits(g2 / ab));
                if (g2 > GSI_ROT2 * ab) jacobi_cs(a, b, g, c, s, t);
            }
            cs[tid][0] = c; cs[tid][1] = s; pq[tid][0] = p; pq[tid][1] = q;
        }
        __syncthreads();
        // column rotations of H and Q: B pairs x M rows x 2 matrices
        for (int e = tid; e < 2 * B * M; e += NTHR) {
            const int i = e % M, t = (e / M) % B, which = e / (M * B);
            const double c = cs[t][0], s = cs[t][1];
            if (s != 0.0) {
                const int p = pq[t][0], q = pq[t][1];
                double* Mx = which ? Q : H;
                const double x = Mx[i * LDH + p], y = Mx[i * LDH + q];
                Mx[i * LDH + p] = c * x - s * y;
                Mx[i * LDH + q] = s * x + c * y;
            }
        }
        __syncthreads();
        // row rotations of H
        for (int e = tid; e < B * M; e += NTHR) {
            const int j = e % M, t = e / M;
            const double c = cs[t][0], s = cs[t][1];
            if (s != 0.0) {
                const int p = pq[t][0], q = pq[t][1];
                const double x = H[p * LDH + j], y = H[q * LDH + j];
                H[p * LDH + j] = c * x - s * y;
                H[q * LDH + j] = s * x + c * y;
            }
        }
        __syncthreads();
    }
    double* Qg = C.Q + ((size_t)u * (C.nb >> 1) + slot) * (M * M);
    for (int e = tid; e < M * M; e += NTHR) Qg[e] = Q[(e / M) * LDH + (e % M)];
    if (tid == 0 && sh_max) atomicMax(&C.smax[u], sh_max);
}

// grid (nb/2, splits, nu), block 128.  P <- P Q for this CTA's rows, in place.  Q sits in shared
// memory with a leading dimension that makes the B-fragment loads conflict free; every warp
// handles 8-row groups: M/4 A-fragment loads, M/8 * M/4 DMMAs, M/8 * 2 stores per lane.
template <int M>
__global__ void __launch_bounds__(128) bj_update_kernel(LChunk C, int round) {
    constexpr int T = M / 8, KS = M / 4, LDQ = M + 4;
    const int u = blockIdx.z, slot = blockIdx.x, split = blockIdx.y;
    int I, J;
    if (!bj_task<M>(C, u, round, slot, I, J)) return;
    const int ld = C.ld[u];
    const int r_begin = split * GSI_BJ_ROWS;
    if (r_begin >= ld) return;
    const int r_end = min(ld, r_begin + GSI_BJ_ROWS);
    double* G = C.G + C.g_off[u];
    const double* Qg = C.Q + ((size_t)u * (C.nb >> 1) + slot) * (M * M);
    __shared__ double Qs[M * LDQ];
    for (int e = threadIdx.x; e < M * M; e += 128) Qs[(e / M) * LDQ + (e % M)] = Qg[e];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lr = lane & 3, lc = lane >> 2;
    // A fragment kk: P[r0 + lc][4kk + lr];  D fragment nt: out[r0 + lc][8nt + 2lr + {0,1}]
    const int c_lo = panel_col<M>(lr, I, J), c_hi = panel_col<M>(M / 2 + lr, I, J);       // columns 4kk+lr: kk < KS/2 in block I, else J
    const int d_lo = panel_col<M>(2 * lr, I, J), d_hi = panel_col<M>(M / 2 + 2 * lr, I, J);
    for (int r0 = r_begin + 8 * warp; r0 < r_end; r0 += 32) {
        double a[KS];
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) {
            const int gc = (kk < KS / 2 ? c_lo : c_hi - M / 2) + 4 * kk;
            a[kk] = G[(size_t)gc * ld + r0 + lc];
        }
        double d[T][2];
#pragma unroll
        for (int nt = 0; nt < T; ++nt) {
            d[nt][0] = 0.0; d[nt][1] = 0.0;
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) dmma884(d[nt][0], d[nt][1], a[kk], Qs[(4 * kk + lr) * LDQ + 8 * nt + lc]);
        }
        __syncwarp();
#pragma unroll
        for (int nt = 0; nt < T; ++nt) {
            const int gc = (nt < T / 2 ? d_lo : d_hi - M / 2) + 8 * nt;
            G[(size_t)gc * ld + r0 + lc] = d[nt][0];
            G[(size_t)(gc + 1) * ld + r0 + lc] = d[nt][1];
        }
    }
}

// one thread per user: close the sweep
__global__ void bj_check_kernel(LChunk C) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= C.nu) return;
    if (!C.done[u]) {
        const int sw = ++C.sweeps[u];
        if (C.smax[u] <= dbits(GSI_STOP2) || sw >= GSI_MAX_SWEEPS) C.done[u] = 1;
        else atomicAdd(C.remaining, 1);
        C.smax[u] = 0ull;
    }
}

// ------------------------------------------------------------------------------------------
// Finalisation
// ------------------------------------------------------------------------------------------

// grid (ceil(ncols/8), 1, nu), block 256: warp per column -> signed norm and eigenvalue
__global__ void fin_colnorm_kernel(LChunk C) {
    const int u = blockIdx.z;
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (c >= C.ncols) return;
    const int lane = threadIdx.x & 31;
    const int n = C.n[u], ld = C.ld[u];
    const double* col = C.G + C.g_off[u] + (size_t)c * ld;
    double a = 0.0, best = -1.0;
    int arg = 0;
    for (int i = lane; i < n; i += 32) {
        const double v = col[i];
        a = fma(v, v, a);
        if (fabs(v) > best) { best = fabs(v); arg = i; }
    }
    a = warp_sum(a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    if (lane == 0) {
        const double nr = sqrt(a);
        // real columns have norm = lambda + 1 >= 1; phantom (zero) columns are marked by norm 0
        const bool real = nr > 0.5;
        C.colnorm[(size_t)u * C.ncols + c] = real ? ((col[arg] < 0.0) ? -nr : nr) : 0.0;
        C.colval[(size_t)u * C.ncols + c] = real ? nr - 1.0 : 1e300;
    }
}

// grid (nu), block 1024: ascending rank by counting, cutoff, eigenvalues out
__global__ void __launch_bounds__(1024) fin_rank_kernel(LChunk C, double* __restrict__ lam_pad) {
    const int u = blockIdx.x;
    const int n = C.n[u];
    const double* val = C.colval + (size_t)u * C.ncols;
    int32_t* perm = C.perm + (size_t)u * C.ncols;
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    const float thr = __double2float_rn(__dadd_rn((double)__uint_as_float(C.sigmax[u]), 0.01));
    int local = 0;
    for (int c = threadIdx.x; c < C.ncols; c += blockDim.x) {
        const double v = val[c];
        int rank = 0;
        for (int o = 0; o < C.ncols; ++o) {
            const double w = val[o];
            rank += (w < v || (w == v && o < c)) ? 1 : 0;
        }
        perm[rank] = c;
        if (v < 1e299 && !(v > (double)thr)) ++local;
    }
    if (local) atomicAdd(&cnt, local);
    __syncthreads();
    const int k = max(cnt, 2);
    if (threadIdx.x == 0) C.k[u] = k;
    double* lam = lam_pad + C.lam_off[u];
    for (int r = threadIdx.x; r < k; r += blockDim.x) lam[r] = (r < n) ? val[perm[r]] : 0.0;
}

// grid (ceil(kmax/32) * ceil(nmax/32), 1, nu), block (32, 8): vec[i*k + r] = G[i, perm[r]] / norm
__global__ void fin_emit_kernel(LChunk C, double* __restrict__ vec_pad, int tiles_r) {
    __shared__ double tile[32][33];
    const int u = blockIdx.z;
    const int n = C.n[u], k = C.k[u];
    const int ti = blockIdx.x / tiles_r, tr = blockIdx.x % tiles_r;
    if (ti * 32 >= n || tr * 32 >= k) return;
    const int ld = C.ld[u];
    const double* G = C.G + C.g_off[u];
    const int32_t* perm = C.perm + (size_t)u * C.ncols;
    const double* cn = C.colnorm + (size_t)u * C.ncols;
    const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        const int rl = ty + 8 * rr;
        const int r = tr * 32 + rl, i = ti * 32 + tx;
        double v = 0.0;
        if (r < k && r < n && i < n) { const int c = perm[r]; v = G[i + (size_t)c * ld] / cn[c]; }
        tile[rl][tx] = v;
    }
    __syncthreads();
    double* vec = vec_pad + C.vec_off[u];
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        const int il = ty + 8 * rr;
        const int i = ti * 32 + il, r = tr * 32 + tx;
        if (i < n && r < k) vec[(size_t)i * k + r] = tile[tx][il];
    }
}
