// Large path, two-stage tridiagonalisation for the big users (the first half of what Eigen's SelfAdjointEigenSolver does at
// precompute_local.cpp:231).  The one-stage kernel (kern_trd.cuh) must stream the trailing matrix
// once per COLUMN (1.375 n^3 bytes per user: HBM bound); here the matrix is first reduced to a band of half-width 64
// with level-3 work only (one symm-like and one syr2k-like pass over the trailing matrix per 64-column PANEL: n^3 / 12
// bytes, 16 x less), the band is reduced to tridiagonal form by bulge chasing on L2-resident data, and the eigenvectors
// of T are transformed back through both stages:
//
//   stage 1  (sy2sb)  per panel p (columns 64 p .., rows r0 = 64 (p+1) ..):
//       sbr_panel_qr_kernel   Householder QR of the m x 64 panel below the band by a team of CTAs that keep the panel in
//                             shared memory (256 rows each; two team barriers per column), compact-WY T built on the way
//       sbr_symm_kernel       Y = A22 V            (FP64 tensor cores; block row I reads the tiles of row I and column I)
//       sbr_w1/w2/w3_kernel   W = Y T;  S = V^T W;  X = W - 1/2 V (T^T S)
//       sbr_syr2k_kernel      A22 -= X V^T + V X^T (FP64 tensor cores, lower tiles)
//     sbr_band_kernel         band -> compact storage AB; A -> "reflector form" (diagonal tiles 0, unit diagonal in the tile
//                             below) so that bt_formt / bt_apply (kern_bt.cuh) apply Q1 unchanged
//   stage 2  (sb2st)  sbr_chase_kernel: bulge chasing, one CTA per sweep, sweeps pipelined three tasks apart through
//                             progress counters; the length-64 reflectors are written as (95 x 32) parallelogram blocks
//     sbr_t2_kernel           compact-WY T of every block
//     sbr_bt2_kernel          Z <- Q2 Z on the kept eigenvectors (FP64 tensor cores), before bt_apply applies Q1
//
// Every reduction has a fixed order: results do not depend on timing, team size or what else runs.
#pragma once
#include "gsi_internal.cuh"
#include "kern_trd.cuh"
#include "ptx.cuh"

#define SBR_B 64              // band half-width = panel width = reflector length of the bulge chase
#define SBR_RB 256            // panel rows per CTA of the QR team
#define SBR_LDP 257           // leading dimension of the panel in shared memory
#define SBR_QP 72             // doubles per (user, part) of the QR scratch: [0] norm^2, [1] alpha, [8..72) column dots
#define SBR_LDB 128           // leading dimension of the compact band AB[j * 128 + (i - j)], 0 <= i - j <= 127
#define SBR_LD 68             // k-stride of staged 64-wide operand blocks (conflict-free DMMA fragments)
#define SBR_G 32              // sweeps per reflector block of stage 2
#define SBR_VROWS 96          // rows of a reflector block (64 + 32 - 1, padded)
#define SBR_BLK_DBL (SBR_VROWS * SBR_G + SBR_G * SBR_G)     // V (96 x 32, column-major) then T (32 x 32, column-major)

struct SbrUser {              // one user of a stage-1 wave / of the stage-2 launch
    int job;                  // index into HJob
    int pad_;
};

struct SbrParams {
    const HJob* jobs;
    const int* users;         // [nusers] job indices of this launch
    double* A;                // tile-major matrices (in place: band + reflectors)
    double* tau;              // [r_off + j]
    double* Vp; double* Wp; double* Xp;     // [r_off * 64 + c * np + r] explicit panels (rows >= r0 used)
    double* Yp;               // [seg][r_off * 64 + c * np + r] partial products of the symm kernel
    double* Sp;               // [r_off * 64 + blk * 4096] partial V^T W per 64-row block
    double* T1;               // [job * 4096] compact-WY factor of the current panel (column-major 64 x 64)
    double* TS;               // [job * 4096] T^T S
    double* qr_part;          // [job][part][SBR_QP]
    unsigned* qr_bar;         // [job]
    int64_t ystride;          // doubles between two segments of Yp
    int p;                    // panel index
    int seg;                  // segments of the symm kernel
    int qr_parts_max;         // leading dimension of qr_part
};

// ---------------------------------------------------------------------------------------------------------------------
// stage 1a: panel QR.  grid (parts_max, nusers), block 256.  CTAs with part >= ceil(m / 256) leave at once; the others
// must be co-resident (the host keeps their number <= the SM count; ~133 KB of shared memory: one CTA per SM).
// ---------------------------------------------------------------------------------------------------------------------
static inline size_t sbr_qr_smem_bytes() { return ((size_t)64 * SBR_LDP + SBR_RB + 64 * 65 + 4 * 64 + 64 + 64 + 16) * sizeof(double); }

__global__ void __launch_bounds__(256, 1) sbr_panel_qr_kernel(SbrParams P) {
    extern __shared__ __align__(16) double qsm[];
    const int job = P.users[blockIdx.y];
    const HJob jb = P.jobs[job];
    const int n = jb.n, np = jb.np, NT = np >> 6, p = P.p;
    const int r0 = (p + 1) * 64, m = n - r0;
    if (m < 2) return;
    const int C = (m + SBR_RB - 1) / SBR_RB, part = blockIdx.x;
    if (part >= C) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* Bs = qsm;                         // [64][SBR_LDP]
    double* vs = Bs + 64 * SBR_LDP;           // [256] current reflector, this CTA's rows
    double* Tm = vs + SBR_RB;                 // [64][65] T (row-major: Tm[i * 65 + j])
    double* part4 = Tm + 64 * 65;             // [4][64] per row group partial dots
    double* wv = part4 + 4 * 64;              // [64] team totals of the column dots
    double* taus = wv + 64;                   // [64]
    double* red = taus + 64;                  // [16]
    const int ncol = min(64, m - 1);
    const int row_lo = r0 + part * SBR_RB;                       // first global row of this CTA
    const int rows = min(SBR_RB, n - row_lo);                    // live rows (the rest of the 256 are padding inside np)
    double* A = P.A + jb.m_off;
    double* scratch = P.qr_part + ((size_t)job * P.qr_parts_max) * SBR_QP;
    double* mine = scratch + (size_t)part * SBR_QP;
    unsigned* bar = P.qr_bar + job;
    unsigned bar_target = 0;
    // ---- load the panel rows: tiles (p + 1 + 4 part + q, p)
    for (int e = tid; e < 64 * SBR_RB; e += 256) {
        const int c = e >> 8, rl = e & 255;
        const int gr = row_lo + rl;
        Bs[c * SBR_LDP + rl] = (gr < n) ? __ldcg(A + hh_tidx(gr, p * 64 + c, NT)) : 0.0;
    }
    for (int e = tid; e < 64 * 65; e += 256) Tm[e] = 0.0;
    if (tid < 64) taus[tid] = 0.0;
    __syncthreads();
    const int g4 = tid >> 6, jj = tid & 63;                      // dot / update role: row group of 64, column jj
    for (int c = 0; c < ncol; ++c) {
        const int dg = r0 + c;                                   // global row of the diagonal of column c (always in part 0)
        // ---- (a) partial ||x[dg+1:]||^2 (thread per row), alpha
        {
            const int gr = row_lo + tid;
            const double x = Bs[c * SBR_LDP + tid];
            double s = (gr > dg && tid < rows) ? x * x : 0.0;
            s = cta_sum_d(s, red);
            if (tid == 0) { mine[0] = s; if (part == 0) mine[1] = Bs[c * SBR_LDP + c]; }
        }
        team_barrier(bar, bar_target, C);
        double beta, tj, scale;
        {
            double s = 0.0;
            for (int q = 0; q < C; ++q) s += __ldcg(scratch + (size_t)q * SBR_QP);      // identical order in every CTA
            const double alpha = __ldcg(scratch + 1);
            if (s == 0.0) { beta = alpha; tj = 0.0; scale = 0.0; }
            else {
                beta = -copysign(sqrt(fma(alpha, alpha, s)), alpha);
                tj = (beta - alpha) / beta;
                scale = 1.0 / (alpha - beta);
            }
        }
        // ---- (b) v (this CTA's rows), stored in place below the diagonal; the diagonal keeps beta (R)
        {
            const int gr = row_lo + tid;
            double v = 0.0;
            if (tid < rows) {
                if (gr > dg) { v = scale * Bs[c * SBR_LDP + tid]; Bs[c * SBR_LDP + tid] = v; }
                else if (gr == dg) { v = 1.0; Bs[c * SBR_LDP + tid] = beta; }
            }
            vs[tid] = v;
            if (tid == 0) taus[c] = tj;
        }
        __syncthreads();
        // ---- (c) column dots with v over my rows: jj > c: w_jj = v . B[:, jj];  jj < c: g_jj = v . V[:, jj] (for T)
        {
            double acc = 0.0;
            if (jj != c) {
                const double* col = Bs + jj * SBR_LDP + 64 * g4;
                const double* vv = vs + 64 * g4;
                if (jj > c) {
#pragma unroll 8
                    for (int r = 0; r < 64; ++r) acc = fma(vv[r], col[r], acc);
                } else {
                    const int dj = r0 + jj - row_lo - 64 * g4;   // local row (inside this group) of column jj's diagonal
#pragma unroll 8
                    for (int r = 0; r < 64; ++r) {
                        const double x = (r > dj) ? col[r] : (r == dj ? 1.0 : 0.0);
                        acc = fma(vv[r], x, acc);
                    }
                }
            }
            part4[g4 * 64 + jj] = acc;
        }
        __syncthreads();
        if (tid < 64) mine[8 + tid] = (part4[tid] + part4[64 + tid]) + (part4[128 + tid] + part4[192 + tid]);
        team_barrier(bar, bar_target, C);
        if (tid < 64) {
            double s = 0.0;
            for (int q = 0; q < C; ++q) s += __ldcg(scratch + (size_t)q * SBR_QP + 8 + tid);
            wv[tid] = s;
        }
        __syncthreads();
        // ---- (d) B[:, jj] -= tau v w_jj for jj > c;  T(0:c, c) = -tau T(0:c, 0:c) g(0:c)
        if (jj > c && tj != 0.0) {
            double* col = Bs + jj * SBR_LDP + 64 * g4;
            const double* vv = vs + 64 * g4;
            const double f = tj * wv[jj];
#pragma unroll 8
            for (int r = 0; r < 64; ++r) col[r] = fma(-f, vv[r], col[r]);
        }
        if (tid < c) {
            double s = 0.0;
            for (int l = tid; l < c; ++l) s = fma(Tm[tid * 65 + l], wv[l], s);
            Tm[tid * 65 + c] = -tj * s;
        } else if (tid == c) Tm[c * 65 + c] = tj;
        __syncthreads();
    }
    // ---- write back: panel in place (R on / above the diagonal of the top block, V below), explicit V, tau, T
    double* Vp = P.Vp + jb.r_off * 64;
    for (int e = tid; e < 64 * SBR_RB; e += 256) {
        const int c = e >> 8, rl = e & 255;
        const int gr = row_lo + rl;
        if (gr < np) {
            const double x = Bs[c * SBR_LDP + rl];
            if (gr < n) A[hh_tidx(gr, p * 64 + c, NT)] = x;
            const int dg = r0 + c;
            Vp[(size_t)c * np + gr] = (c < ncol && gr < n) ? (gr > dg ? x : (gr == dg ? 1.0 : 0.0)) : 0.0;
        }
    }
    if (part == 0) {
        if (tid < 64) P.tau[jb.r_off + p * 64 + tid] = taus[tid];
        double* T1 = P.T1 + (size_t)job * 4096;
        for (int e = tid; e < 4096; e += 256) { const int j = e >> 6, i = e & 63; T1[e] = Tm[i * 65 + j]; }    // column-major
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// stage 1b: Y = A22 V on the FP64 tensor cores.  grid (block rows * seg, nusers), block 256.  CTA (bi, s) owns the 64
// rows of block row I = p + 1 + bi and the block columns J of segment s:  J < I: tile (I, J);  J == I: the diagonal tile
// (stored full);  J > I: tile (J, I) transposed.  Every tile of the lower triangle is read twice per panel (once per role)
// -- mostly from L2 -- and there is no atomics / no ordering problem: Y_I[seg] is produced by one CTA in a fixed order.
// Operand blocks are staged with cp.async (3 stages of A tile + V block, k-stride 68).
// ---------------------------------------------------------------------------------------------------------------------
#define SBR_SYMM_STAGES 3
#define SBR_SYMM_STAGE_DBL (2 * 64 * SBR_LD)
static inline size_t sbr_symm_smem_bytes() { return (size_t)SBR_SYMM_STAGES * SBR_SYMM_STAGE_DBL * sizeof(double); }

__global__ void __launch_bounds__(256, 1) sbr_symm_kernel(SbrParams P) {
    extern __shared__ __align__(16) double ssm[];
    const int job = P.users[blockIdx.y];
    const HJob jb = P.jobs[job];
    const int np = jb.np, NT = np >> 6, p = P.p;
    const int n = jb.n, r0 = (p + 1) * 64;
    if (n - r0 < 2) return;
    const int nb = NT - (p + 1);                                 // trailing block rows
    const int bi = blockIdx.x / P.seg, sg = blockIdx.x % P.seg;
    if (bi >= nb) return;
    const int I = p + 1 + bi;
    const int per = (nb + P.seg - 1) / P.seg;
    const int jb0 = p + 1 + sg * per, jb1 = min(NT, jb0 + per);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, fk = lane & 3, fr = lane >> 2;
    const double* A = P.A + jb.m_off;
    const double* Vp = P.Vp + jb.r_off * 64;
    double* Y = P.Yp + (size_t)sg * P.ystride + jb.r_off * 64;
    double acc[8][2];
#pragma unroll
    for (int rb = 0; rb < 8; ++rb) { acc[rb][0] = 0.0; acc[rb][1] = 0.0; }
    auto load = [&](int J, int stg) {
        double* As = ssm + (size_t)stg * SBR_SYMM_STAGE_DBL;
        double* Vs = As + 64 * SBR_LD;
        const double* tile = (J <= I) ? A + (((size_t)J * NT + I) << 12) : A + (((size_t)I * NT + J) << 12);
#pragma unroll
        for (int t = 0; t < 8; ++t) {                            // 64 columns x 32 16-byte chunks
            const int e = tid + t * 256, c = e >> 5, r2 = (e & 31) * 2;
            cp_async16(As + c * SBR_LD + r2, tile + (c << 6) + r2);
            cp_async16(Vs + c * SBR_LD + r2, Vp + (size_t)c * np + J * 64 + r2);
        }
    };
    const int nj = jb1 - jb0;
    if (nj > 0) load(jb0, 0);
    cp_async_commit();
    if (nj > 1) load(jb0 + 1, 1);
    cp_async_commit();
    for (int t = 0; t < nj; ++t) {
        const int J = jb0 + t;
        cp_async_wait<1>();
        __syncthreads();
        if (t + 2 < nj) load(J + 2, (t + 2) % SBR_SYMM_STAGES);
        cp_async_commit();
        const double* As = ssm + (size_t)(t % SBR_SYMM_STAGES) * SBR_SYMM_STAGE_DBL;
        const double* Vs = As + 64 * SBR_LD;
        // D[c][r] += V[k][c] * A[r][k]:  A operand = Vs[c][k], B operand = A[r][k] (direct: As[k][r];  transposed role: As[r][k])
        const double* va = Vs + (8 * warp + fr) * SBR_LD + fk;
        if (J <= I) {
            const double* ab = As + fk * SBR_LD + fr;
#pragma unroll 4
            for (int q = 0; q < 16; ++q) {
                const double a = va[4 * q];
#pragma unroll
                for (int rb = 0; rb < 8; ++rb) dmma(acc[rb][0], acc[rb][1], a, ab[(4 * q) * SBR_LD + 8 * rb]);
            }
        } else {
            const double* ab = As + fr * SBR_LD + fk;
#pragma unroll 4
            for (int q = 0; q < 16; ++q) {
                const double a = va[4 * q];
#pragma unroll
                for (int rb = 0; rb < 8; ++rb) dmma(acc[rb][0], acc[rb][1], a, ab[(8 * rb) * SBR_LD + 4 * q]);
            }
        }
    }
    cp_async_wait<0>();
    // lane holds D[c = 8 warp + fr][r = 8 rb + 2 fk + {0, 1}]
    double* yc = Y + (size_t)(8 * warp + fr) * np + I * 64 + 2 * fk;
#pragma unroll
    for (int rb = 0; rb < 8; ++rb) *(double2*)(yc + 8 * rb) = make_double2(acc[rb][0], acc[rb][1]);
}

// ---------------------------------------------------------------------------------------------------------------------
// stage 1c: W = Y T (w1, also the per-block partial of S = V^T W), TS = T^T S (w2), X = W - 1/2 V TS (w3).  Scalar FP64:
// 64 x 64 x 64 products per CTA, a 4 x 4 register block per thread.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void sbr_mm64(const double* __restrict__ Am, int lda, bool ta, const double* __restrict__ Bm, int ldb, double (&acc)[4][4], int tr, int tc) {
    // acc[a][b] += sum_k A(4 tr + a, k) B(k, 4 tc + b);  A(i, k) = ta ? Am[i * lda + k] : Am[k * lda + i];  B(k, j) = Bm[j * ldb + k]
#pragma unroll 4
    for (int k = 0; k < 64; ++k) {
        double x[4], y[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) x[a] = ta ? Am[(4 * tr + a) * lda + k] : Am[k * lda + 4 * tr + a];
#pragma unroll
        for (int b = 0; b < 4; ++b) y[b] = Bm[(4 * tc + b) * ldb + k];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = fma(x[a], y[b], acc[a][b]);
    }
}

// grid (block rows, nusers), block 256, smem 3 x 64 x 65 doubles
static inline size_t sbr_w_smem_bytes() { return (size_t)3 * 64 * 65 * sizeof(double); }
__global__ void __launch_bounds__(256) sbr_w1_kernel(SbrParams P) {
    extern __shared__ __align__(16) double wsm[];
    const int job = P.users[blockIdx.y];
    const HJob jb = P.jobs[job];
    const int np = jb.np, NT = np >> 6, p = P.p, r0 = (p + 1) * 64;
    if (jb.n - r0 < 2 || (int)blockIdx.x >= NT - (p + 1)) return;
    const int I = p + 1 + blockIdx.x, tid = threadIdx.x, tr = tid >> 4, tc = tid & 15;
    double* Ys = wsm;                 // [c][r] stride 65
    double* Ts = Ys + 64 * 65;        // T column-major: Ts[j * 65 + i]
    double* Ws = Ts + 64 * 65;
    const double* T1 = P.T1 + (size_t)job * 4096;
    for (int e = tid; e < 4096; e += 256) {
        const int c = e >> 6, r = e & 63;
        double y = 0.0;
        for (int s = 0; s < P.seg; ++s) y += __ldcg(P.Yp + (size_t)s * P.ystride + jb.r_off * 64 + (size_t)c * np + I * 64 + r);
        Ys[c * 65 + r] = y;
        Ts[c * 65 + r] = T1[e];
    }
    __syncthreads();
    double acc[4][4] = {};
    sbr_mm64(Ys, 65, false, Ts, 65, acc, tr, tc);             // W(r, j) = sum_k Y(r, k) T(k, j)
    double* Wp = P.Wp + jb.r_off * 64;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            Ws[(4 * tc + b) * 65 + 4 * tr + a] = acc[a][b];
            Wp[(size_t)(4 * tc + b) * np + I * 64 + 4 * tr + a] = acc[a][b];
        }
    // V block into Ys
    __syncthreads();
    const double* Vp = P.Vp + jb.r_off * 64;
    for (int e = tid; e < 4096; e += 256) { const int c = e >> 6, r = e & 63; Ys[c * 65 + r] = __ldcg(Vp + (size_t)c * np + I * 64 + r); }
    __syncthreads();
    double s2[4][4] = {};
    sbr_mm64(Ys, 65, true, Ws, 65, s2, tr, tc);               // S(i, j) = sum_r V(r, i) W(r, j):  A(i, k) = Ys[i * 65 + k]
    double* Sp = P.Sp + jb.r_off * 64 + (size_t)blockIdx.x * 4096;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) Sp[(4 * tc + b) * 64 + 4 * tr + a] = s2[a][b];
}

// grid (nusers), block 256: S = sum of the block partials (ascending), TS = T^T S
__global__ void __launch_bounds__(256) sbr_w2_kernel(SbrParams P) {
    extern __shared__ __align__(16) double wsm[];
    const int job = P.users[blockIdx.x];
    const HJob jb = P.jobs[job];
    const int NT = jb.np >> 6, p = P.p, r0 = (p + 1) * 64;
    if (jb.n - r0 < 2) return;
    const int nb = NT - (p + 1), tid = threadIdx.x, tr = tid >> 4, tc = tid & 15;
    double* Ss = wsm;
    double* Ts = Ss + 64 * 65;
    const double* T1 = P.T1 + (size_t)job * 4096;
    const double* Sp = P.Sp + jb.r_off * 64;
    for (int e = tid; e < 4096; e += 256) {
        double s = 0.0;
        for (int b = 0; b < nb; ++b) s += __ldcg(Sp + (size_t)b * 4096 + e);
        Ss[(e >> 6) * 65 + (e & 63)] = s;
        Ts[(e >> 6) * 65 + (e & 63)] = T1[e];
    }
    __syncthreads();
    double acc[4][4] = {};
    sbr_mm64(Ts, 65, true, Ss, 65, acc, tr, tc);              // TS(i, j) = sum_k T(k, i) S(k, j):  A(i, k) = Ts[i * 65 + k] = T(k, i)
    double* TS = P.TS + (size_t)job * 4096;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) TS[(4 * tc + b) * 64 + 4 * tr + a] = acc[a][b];
}

// grid (block rows, nusers), block 256: X = W - 1/2 V TS
__global__ void __launch_bounds__(256) sbr_w3_kernel(SbrParams P) {
    extern __shared__ __align__(16) double wsm[];
    const int job = P.users[blockIdx.y];
    const HJob jb = P.jobs[job];
    const int np = jb.np, NT = np >> 6, p = P.p, r0 = (p + 1) * 64;
    if (jb.n - r0 < 2 || (int)blockIdx.x >= NT - (p + 1)) return;
    const int I = p + 1 + blockIdx.x, tid = threadIdx.x, tr = tid >> 4, tc = tid & 15;
    double* Vs = wsm;
    double* Ss = Vs + 64 * 65;
    const double* Vp = P.Vp + jb.r_off * 64;
    const double* TS = P.TS + (size_t)job * 4096;
    for (int e = tid; e < 4096; e += 256) {
        const int c = e >> 6, r = e & 63;
        Vs[c * 65 + r] = __ldcg(Vp + (size_t)c * np + I * 64 + r);
        Ss[c * 65 + r] = __ldcg(TS + e);
    }
    __syncthreads();
    double acc[4][4] = {};
    sbr_mm64(Vs, 65, false, Ss, 65, acc, tr, tc);             // (V TS)(r, j)
    const double* Wp = P.Wp + jb.r_off * 64;
    double* Xp = P.Xp + jb.r_off * 64;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const size_t o = (size_t)(4 * tc + b) * np + I * 64 + 4 * tr + a;
            Xp[o] = __ldcg(Wp + o) - 0.5 * acc[a][b];
        }
}

// ---------------------------------------------------------------------------------------------------------------------
// stage 1d: A22 -= X V^T + V X^T on the tiles I >= J >= p + 1 (diagonal tiles in full), FP64 tensor cores.
// grid (ctas, nusers), block 256: CTA c walks the tiles c, c + ctas, ... of the user in column-major tile order; operand
// k-chunks of 16 are double buffered with cp.async (the scheme of the trailing update in kern_trd.cuh).
// ---------------------------------------------------------------------------------------------------------------------
static inline size_t sbr_syr2k_smem_bytes() { return (size_t)TRD_SYR_DBL * sizeof(double); }
__global__ void __launch_bounds__(256, 2) sbr_syr2k_kernel(SbrParams P) {
    extern __shared__ __align__(16) double ksm[];
    const int job = P.users[blockIdx.y];
    const HJob jb = P.jobs[job];
    const int np = jb.np, NT = np >> 6, p = P.p, r0 = (p + 1) * 64;
    if (jb.n - r0 < 2) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* A = P.A + jb.m_off;
    const double* Vp = P.Vp + jb.r_off * 64;
    const double* Xp = P.Xp + jb.r_off * 64;
    const int ld = np;
    auto stage_unit = [&](int I, int J, int q, int buf) {      // rows of V_I, X_I, V_J, X_J x 16 panel columns -> S[buf][which][k][68]
        double* dst = ksm + buf * (4 * TRD_SYR_LD * TRD_SYR_KH);
        const int kh = q * TRD_SYR_KH;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int e = tid + t * 256;                        // 4 blocks x 16 k x 32 16-byte chunks
            const int which = e >> 9, k = (e >> 5) & 15, r2 = (e & 31) * 2;
            const double* src = ((which & 1) ? Xp : Vp) + (size_t)(kh + k) * ld + ((which >> 1) ? J : I) * 64 + r2;
            cp_async16(dst + which * (TRD_SYR_LD * TRD_SYR_KH) + k * TRD_SYR_LD + r2, src);
        }
    };
    TileWalk w, wn;
    w.init(p + 1, NT, blockIdx.x, gridDim.x);
    wn = w;
    int buf = 0;
    if (w.valid()) stage_unit(w.I, w.J, 0, 0);
    cp_async_commit();
    const int fk = lane & 3, fr = lane >> 2;
    for (; w.valid(); w.next()) {
        const int I = w.I, J = w.J;
        wn.next();
        const int lc = 8 * warp + fr;
        double* ctile = A + (((size_t)J * NT + I) << 12) + lc * 64;
        double2 cv[8];
#pragma unroll
        for (int rb = 0; rb < 8; ++rb) cv[rb] = __ldcg((const double2*)(ctile + 8 * rb + 2 * fk));
        double acc[8][2];
#pragma unroll
        for (int rb = 0; rb < 8; ++rb) { acc[rb][0] = 0.0; acc[rb][1] = 0.0; }
        for (int q = 0; q < 4; ++q) {
            if (q + 1 < 4) stage_unit(I, J, q + 1, buf ^ 1);
            else if (wn.valid()) stage_unit(wn.I, wn.J, 0, buf ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
            __syncthreads();
            const double* S = ksm + buf * (4 * TRD_SYR_LD * TRD_SYR_KH);
            const double* VIs = S, *XIs = S + TRD_SYR_LD * TRD_SYR_KH, *VJs = S + 2 * TRD_SYR_LD * TRD_SYR_KH,
                         *XJs = S + 3 * TRD_SYR_LD * TRD_SYR_KH;
#pragma unroll
            for (int k0 = 0; k0 < TRD_SYR_KH; k0 += 4) {
                const double aX = XJs[(k0 + fk) * TRD_SYR_LD + 8 * warp + fr];
                const double aV = VJs[(k0 + fk) * TRD_SYR_LD + 8 * warp + fr];
#pragma unroll
                for (int rb = 0; rb < 8; ++rb) {
                    const double bV = VIs[(k0 + fk) * TRD_SYR_LD + 8 * rb + fr];
                    const double bX = XIs[(k0 + fk) * TRD_SYR_LD + 8 * rb + fr];
                    dmma(acc[rb][0], acc[rb][1], aX, bV);
                    dmma(acc[rb][0], acc[rb][1], aV, bX);
                }
            }
            __syncthreads();
            buf ^= 1;
        }
#pragma unroll
        for (int rb = 0; rb < 8; ++rb) {
            double2 v = cv[rb];
            v.x -= acc[rb][0]; v.y -= acc[rb][1];
            *(double2*)(ctile + 8 * rb + 2 * fk) = v;
        }
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------------------------------
// after the last panel: band -> AB (compact, zero padded), A -> reflector form.  grid (NT_max, nusers), block 256.
// CTA p handles the 64 columns of block column p: AB[j * 128 + d] = A(j + d, j), d = 0 .. 64; then tile (p, p) = 0 and tile
// (p + 1, p) = unit lower trapezoid of the panel's reflectors (zeros above the diagonal, zero columns where tau == 0 because
// there is no reflector).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sbr_band_kernel(SbrParams P, double* __restrict__ ABall) {
    const int job = P.users[blockIdx.y];
    const HJob jb = P.jobs[job];
    const int n = jb.n, np = jb.np, NT = np >> 6, p = blockIdx.x;
    if (p >= NT) return;
    double* A = P.A + jb.m_off;
    double* AB = ABall + jb.r_off * SBR_LDB;
    const int tid = threadIdx.x;
    for (int e = tid; e < 64 * SBR_LDB; e += 256) {
        const int c = e >> 7, d = e & 127, j = p * 64 + c, i = j + d;
        double x = 0.0;
        if (d <= 64 && i < n && j < n) x = __ldcg(A + hh_tidx(i, j, NT));
        AB[(size_t)j * SBR_LDB + d] = x;
    }
    __syncthreads();
    double* dt = A + (((size_t)p * NT + p) << 12);
    for (int e = tid; e < 4096; e += 256) dt[e] = 0.0;
    if (p + 1 < NT) {
        const int m = n - (p + 1) * 64, ncol = min(64, m - 1);
        double* st = A + (((size_t)p * NT + p + 1) << 12);
        for (int e = tid; e < 4096; e += 256) {
            const int c = e >> 6, r = e & 63;
            if (c >= ncol || r < c) st[e] = 0.0;
            else if (r == c) st[e] = 1.0;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// stage 2: band -> tridiagonal by bulge chasing.
//
// Sweep s annihilates column s below the sub-diagonal with a reflector on the rows s+1 .. s+64 (task 0) and chases the
// bulge this creates down the band in steps of 64 rows (tasks 1, 2, ...).  Task t of sweep s works on
//     r0 = s + 1 + 64 t,  r1 = min(r0 + 64, n),  col = t ? r0 - 64 : s,   x = A[r0:r1, col]  ->  H = I - tau v v^T,  H x = beta e1
//     L = A[r0:r1, col+1 : r0]  <- H L          (t > 0: the block the previous task filled in)
//     D = A[r0:r1, r0:r1]       <- H D H
//     B = A[r1:r1+64, r0:r1]    <- B H          (fills in below the band: the next task's L and x)
// One CTA runs a whole sweep, keeping B in shared memory as the next task's L.  Sweep s may run task t once sweep s-1 has
// completed task t+2 (they overlap in one element), which is tracked by per-sweep progress counters in global memory
// (release / acquire); the sweeps are handed out in order from a queue, so whoever waits, waits for a running CTA.
// The reflector of (s, t) is column s % 32 of the (96 x 32) block (s / 32, t), rows (s % 32) .. (s % 32) + 63.
// ---------------------------------------------------------------------------------------------------------------------
struct ChaseParams {
    const HJob* jobs;
    const int2* list;         // (job, sweep), sweep-major so that sweep s of a user is handed out before sweep s + 1
    int nitems;
    int* queue;
    double* AB;               // [r_off * 128 ..]
    int* prog;                // [r_off + s] completed tasks of sweep s (1 << 30: finished)
    double* V2;               // reflector blocks
    const int64_t* v2_off;    // [job] first block of the user (in doubles)
    const int* goff;          // [goff_off[job] + G] blocks before group G of the user
    const int* goff_off;      // [job]
    double* d; double* e;     // tridiagonal out (sbr_de_kernel)
};
#define SBR_DONE (1 << 30)
#define CH_LD 65
static inline size_t sbr_chase_smem_bytes() { return ((size_t)3 * 64 * CH_LD + 8 * 64 + 64) * sizeof(double); }

__global__ void __launch_bounds__(256, 2) sbr_chase_kernel(ChaseParams P) {
    extern __shared__ __align__(16) double csm[];
    __shared__ int item_s, stop_s;
    double* Lb = csm;                         // [64][65] column-major blocks
    double* Db = Lb + 64 * CH_LD;
    double* Bb = Db + 64 * CH_LD;
    double* part = Bb + 64 * CH_LD;           // [4][64] partial sums
    double* vv = part + 4 * 64;               // [64] v
    double* pw = vv + 64;                     // [64] p / w / u
    double* xs = pw + 64;                     // [64] x
    double* sc = xs + 64;                     // [8] scalars
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ci = tid & 63, cq = tid >> 6;   // (column / row index, quarter)
    for (;;) {
        __syncthreads();
        if (tid == 0) item_s = atomicAdd(P.queue, 1);
        __syncthreads();
        const int it = item_s;
        if (it >= P.nitems) break;
        const int2 item = P.list[it];
        const HJob jb = P.jobs[item.x];
        const int n = jb.n, s = item.y;
        double* AB = P.AB + jb.r_off * SBR_LDB;
        int* prog = P.prog + jb.r_off;
        const int G = s >> 5, g = s & 31;
        double* V2 = P.V2 + P.v2_off[item.x] + (size_t)P.goff[P.goff_off[item.x] + G] * SBR_BLK_DBL;
        for (int t = 0;; ++t) {
            const int r0 = s + 1 + 64 * t;
            if (r0 > n - 2) break;
            const int r1 = min(r0 + 64, n), len = r1 - r0, hi = min(n, r1 + 64), lenB = hi - r1;
            const int col = t ? r0 - 64 : s;
            // ---- dependency: sweep s-1 has completed task t+2 (or has finished)
            if (s > 0) {
                if (tid == 0) { while (ld_acquire_u32((const unsigned*)(prog + s - 1)) < (unsigned)(t + 3)) {} __threadfence(); }
                __syncthreads();
            }
            // ---- load x, D (mirrored), B
            if (t == 0) { if (tid < 64) xs[tid] = (tid < len) ? __ldcg(AB + (size_t)col * SBR_LDB + 1 + tid) : 0.0; }
            else if (tid < 64) xs[tid] = (tid < len) ? Lb[tid] : 0.0;
            for (int e = tid; e < 64 * 64; e += 256) {
                const int c = e >> 6, r = e & 63;
                if (r >= c) {
                    const double x = (r < len) ? __ldcg(AB + (size_t)(r0 + c) * SBR_LDB + (r - c)) : 0.0;
                    Db[c * CH_LD + r] = x; Db[r * CH_LD + c] = x;
                }
                Bb[c * CH_LD + r] = (r < lenB && c < len) ? __ldcg(AB + (size_t)(r0 + c) * SBR_LDB + (len + r - c)) : 0.0;
            }
            __syncthreads();
            // ---- reflector (warp 0)
            if (warp == 0) {
                const double x0 = xs[lane], x1 = xs[lane + 32];
                double sg = ((lane > 0) ? x0 * x0 : 0.0) + x1 * x1;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sg += __shfl_xor_sync(0xffffffffu, sg, o);
                const double alpha = xs[0];
                double beta = alpha, tj = 0.0, scale = 0.0;
                if (sg != 0.0) {
                    beta = -copysign(sqrt(fma(alpha, alpha, sg)), alpha);
                    tj = (beta - alpha) / beta;
                    scale = 1.0 / (alpha - beta);
                }
                vv[lane] = (lane == 0) ? 1.0 : scale * x0;
                vv[lane + 32] = scale * x1;
                if (lane == 0) { sc[0] = beta; sc[1] = tj; stop_s = (sg == 0.0); }
            }
            __syncthreads();
            const double beta = sc[0], tj = sc[1];
            const bool stop = stop_s != 0;
            if (stop) {
                // nothing to annihilate: the sweep ends; the block filled in by the previous task still has to go home
                if (t > 0)
                    for (int e = tid; e < 64 * 64; e += 256) {
                        const int c = e >> 6, r = e & 63;
                        if (r < len) AB[(size_t)(col + c) * SBR_LDB + (64 - c + r)] = Lb[c * CH_LD + r];
                    }
                break;
            }
            // ---- the reflector goes to its block; column col = beta e1
            {
                double* blk = V2 + (size_t)t * SBR_BLK_DBL;
                if (tid < len) blk[g * SBR_VROWS + g + tid] = vv[tid];
                if (tid == 64) blk[SBR_VROWS * SBR_G + g * SBR_G + g] = tj;
                if (t == 0) { if (tid < len) AB[(size_t)col * SBR_LDB + 1 + tid] = (tid == 0) ? beta : 0.0; }
                else if (tid < 64) Lb[tid] = (tid == 0) ? beta : 0.0;
            }
            // ---- L <- H L (columns 1 .. 63 of the previous task's B), then home
            if (t > 0) {
                double acc = 0.0;
                const double* lc = Lb + ci * CH_LD + 16 * cq;
#pragma unroll
                for (int r = 0; r < 16; ++r) acc = fma(vv[16 * cq + r], lc[r], acc);
                part[cq * 64 + ci] = acc;
                __syncthreads();
                if (ci > 0) {
                    const double w = tj * ((part[ci] + part[64 + ci]) + (part[128 + ci] + part[192 + ci]));
                    double* lw = Lb + ci * CH_LD + 16 * cq;
#pragma unroll
                    for (int r = 0; r < 16; ++r) lw[r] = fma(-w, vv[16 * cq + r], lw[r]);
                }
                __syncthreads();
                for (int e = tid; e < 64 * 64; e += 256) {
                    const int c = e >> 6, r = e & 63;
                    if (r < len) AB[(size_t)(col + c) * SBR_LDB + (64 - c + r)] = Lb[c * CH_LD + r];
                }
            }
            // ---- D <- H D H:  p = tau D v,  w = p - (tau / 2)(p . v) v,  D -= v w^T + w v^T
            {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < 16; ++j) acc = fma(Db[(16 * cq + j) * CH_LD + ci], vv[16 * cq + j], acc);
                part[cq * 64 + ci] = acc;
                __syncthreads();
                if (tid < 64) pw[tid] = tj * ((part[tid] + part[64 + tid]) + (part[128 + tid] + part[192 + tid]));
                __syncthreads();
                if (warp == 0) {
                    double pv = pw[lane] * vv[lane] + pw[lane + 32] * vv[lane + 32];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) pv += __shfl_xor_sync(0xffffffffu, pv, o);
                    if (lane == 0) sc[2] = -0.5 * tj * pv;
                }
                __syncthreads();
                const double a2 = sc[2];
                if (tid < 64) pw[tid] = fma(a2, vv[tid], pw[tid]);
                __syncthreads();
                const double vi = vv[ci], wi = pw[ci];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int jj = 16 * cq + j;
                    double* dp = Db + jj * CH_LD + ci;
                    *dp = *dp - (vi * pw[jj] + wi * vv[jj]);
                }
                __syncthreads();
                for (int e = tid; e < 64 * 64; e += 256) {
                    const int c = e >> 6, r = e & 63;
                    if (r >= c && r < len) AB[(size_t)(r0 + c) * SBR_LDB + (r - c)] = Db[c * CH_LD + r];
                }
            }
            // ---- B <- B H:  u = B v,  B -= tau u v^T
            {
                double acc = 0.0;
#pragma unroll
                for (int c = 0; c < 16; ++c) acc = fma(Bb[(16 * cq + c) * CH_LD + ci], vv[16 * cq + c], acc);
                part[cq * 64 + ci] = acc;
                __syncthreads();
                const double u = tj * ((part[ci] + part[64 + ci]) + (part[128 + ci] + part[192 + ci]));
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    double* bp = Bb + (16 * cq + c) * CH_LD + ci;
                    *bp = fma(-u, vv[16 * cq + c], *bp);
                }
                __syncthreads();
            }
            const bool more = (r0 + 64 <= n - 2);
            if (!more && lenB > 0)
                for (int e = tid; e < 64 * 64; e += 256) {
                    const int c = e >> 6, r = e & 63;
                    if (r < lenB && c < len) AB[(size_t)(r0 + c) * SBR_LDB + (len + r - c)] = Bb[c * CH_LD + r];
                }
            // ---- progress
            __syncthreads();
            if (tid == 0) { __threadfence(); atomicExch(prog + s, t + 1); }
            double* tmp = Lb; Lb = Bb; Bb = tmp;
        }
        __syncthreads();
        if (tid == 0) { __threadfence(); atomicExch(prog + s, SBR_DONE); }
    }
}

// d, e of the finished band.  grid (ceil(npmax / 256), nusers)
__global__ void sbr_de_kernel(SbrParams P, const double* __restrict__ ABall, double* __restrict__ dall, double* __restrict__ eall) {
    const HJob jb = P.jobs[P.users[blockIdx.y]];
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= jb.n) return;
    const double* AB = ABall + jb.r_off * SBR_LDB;
    dall[jb.r_off + j] = AB[(size_t)j * SBR_LDB];
    if (j < jb.n - 1) eall[jb.r_off + j] = AB[(size_t)j * SBR_LDB + 1];
}

// ---------------------------------------------------------------------------------------------------------------------
// compact-WY factor of every reflector block:  H_0 H_1 ... H_31 = I - V T V^T  (forward, columnwise: dlarft),
// T(i, i) = tau_i (left there by the chase), T(0:i, i) = -tau_i T(0:i, 0:i) (V(:, 0:i)^T v_i).  grid (blocks), block 256.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sbr_t2_kernel(double* __restrict__ V2, int64_t nblocks) {
    __shared__ double Vs[SBR_G][SBR_VROWS + 1];
    __shared__ double Gm[SBR_G][SBR_G + 1];
    __shared__ double Tm[SBR_G][SBR_G + 1];
    const int64_t b = blockIdx.x;
    if (b >= nblocks) return;
    double* blk = V2 + b * SBR_BLK_DBL;
    double* T = blk + SBR_VROWS * SBR_G;
    const int tid = threadIdx.x;
    for (int e = tid; e < SBR_VROWS * SBR_G; e += 256) Vs[e / SBR_VROWS][e % SBR_VROWS] = blk[e];
    __syncthreads();
    for (int e = tid; e < SBR_G * SBR_G; e += 256) {
        const int i = e >> 5, j = e & 31;
        double s = 0.0;
        if (i < j) {
            // v_i lives in rows i .. i+63, v_j in rows j .. j+63
            for (int r = j; r < i + 64; ++r) s = fma(Vs[i][r], Vs[j][r], s);
        }
        Gm[i][j] = s;
        Tm[i][j] = 0.0;
    }
    __syncthreads();
    for (int j = 0; j < SBR_G; ++j) {
        const double tj = T[j * SBR_G + j];
        if (tid < j) {
            double s = 0.0;
            for (int l = tid; l < j; ++l) s = fma(Tm[tid][l], Gm[l][j], s);
            Tm[tid][j] = -tj * s;
        } else if (tid == j) Tm[j][j] = tj;
        __syncthreads();
    }
    for (int e = tid; e < SBR_G * SBR_G; e += 256) { const int j = e >> 5, i = e & 31; T[e] = Tm[i][j]; }    // column-major
}

// ---------------------------------------------------------------------------------------------------------------------
// Z <- Q2 Z on the kept eigenvectors.  Q2 = prod over groups G ascending of (B_G[K] ... B_G[1] B_G[0]), B_G[t] = I - V T V^T on
// the rows 32 G + 1 + 64 t .. + 95 (blocks of one group with different t commute where they have to: kern_sbr.cuh header).
// A work item is a 32-column block of Z of one user; its CTA walks the groups from the last to the first and, inside a
// group, t upwards:   X = V^T Z_rows (32 x 32),  X <- T X,  Z_rows -= V X    -- all three on the FP64 tensor cores,
// operands in shared memory (k-strides == 4 mod 16: conflict-free fragments), V and T of the next block prefetched with
// cp.async while this one is applied.
// ---------------------------------------------------------------------------------------------------------------------
struct Bt2Params {
    const HJob* jobs;
    const int2* items;        // (job, column block)
    int nitems;
    int* queue;
    const double* V2;
    const int64_t* v2_off;
    const int* goff; const int* goff_off;
    double* Qa; double* Qb;
    const int32_t* kuser;
};
#define BT2_LDZ 100
#define BT2_LDX 36
#define BT2_VT_DBL (SBR_G * BT2_LDZ + SBR_G * BT2_LDX)       // staged V (32 x 100) + T (32 x 36)
static inline size_t sbr_bt2_smem_bytes() { return ((size_t)SBR_G * BT2_LDZ + 2 * BT2_VT_DBL + 2 * SBR_G * BT2_LDX) * sizeof(double); }

__global__ void __launch_bounds__(256, 2) sbr_bt2_kernel(Bt2Params P) {
    extern __shared__ __align__(16) double bsm[];
    __shared__ int item_s;
    double* Zs = bsm;                                  // [32 z columns][100]
    double* VT0 = Zs + SBR_G * BT2_LDZ;                // two stages of V [32][100] + T [32][36]
    double* Xs = VT0 + 2 * BT2_VT_DBL;                 // [32 z columns][36]
    double* X2 = Xs + SBR_G * BT2_LDX;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, fk = lane & 3, fr = lane >> 2;
    for (;;) {
        __syncthreads();
        if (tid == 0) item_s = atomicAdd(P.queue, 1);
        __syncthreads();
        const int it = item_s;
        if (it >= P.nitems) break;
        const int2 item = P.items[it];
        const HJob jb = P.jobs[item.x];
        const int lim = P.kuser[item.x], c0 = item.y * 32;
        if (c0 >= lim) continue;
        const int n = jb.n, np = jb.np;
        double* Z = ((jb.levels & 1) ? P.Qb : P.Qa) + jb.m_off + (size_t)c0 * np;
        const double* V2 = P.V2 + P.v2_off[item.x];
        const int* goff = P.goff + P.goff_off[item.x];
        const int ngroups = (n - 2 + 31) / 32;                           // sweeps 0 .. n-3
        auto ntasks = [&](int G) { const int r = n - 3 - 32 * G; return r < 0 ? 0 : r / 64 + 1; };
        auto load_vt = [&](int G, int t, int stg) {
            double* Vs = VT0 + (size_t)stg * BT2_VT_DBL;
            double* Ts = Vs + SBR_G * BT2_LDZ;
            const double* blk = V2 + (size_t)(goff[G] + t) * SBR_BLK_DBL;
            for (int e = tid; e < SBR_G * (SBR_VROWS / 2); e += 256) {       // 32 columns x 48 16-byte chunks
                const int c = e / 48, r2 = (e % 48) * 2;
                cp_async16(Vs + c * BT2_LDZ + r2, blk + c * SBR_VROWS + r2);
            }
            for (int e = tid; e < SBR_G * (SBR_G / 2); e += 256) {           // T: 32 columns x 16 chunks
                const int c = e >> 4, r2 = (e & 15) * 2;
                cp_async16(Ts + c * BT2_LDX + r2, blk + SBR_VROWS * SBR_G + c * SBR_G + r2);
            }
        };
        // flat walk over (G descending, t ascending)
        int G = ngroups - 1, t = 0;
        while (G >= 0 && ntasks(G) == 0) --G;
        if (G < 0) continue;
        int stg = 0;
        load_vt(G, 0, 0);
        cp_async_commit();
        while (G >= 0) {
            // next block (for the prefetch)
            int Gn = G, tn = t + 1;
            if (tn >= ntasks(G)) { Gn = G - 1; tn = 0; }
            const int first = 32 * G + 1 + 64 * t;                        // first row of the block
            // ---- Z rows -> shared memory (zero beyond np)
            for (int e = tid; e < 32 * SBR_VROWS; e += 256) {              // (first is odd: scalar 8-byte accesses)
                const int c = e / SBR_VROWS, r = e % SBR_VROWS, gr = first + r;
                Zs[c * BT2_LDZ + r] = (gr < np) ? Z[(size_t)c * np + gr] : 0.0;
            }
            if (Gn >= 0) load_vt(Gn, tn, stg ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
            __syncthreads();
            const double* Vs = VT0 + (size_t)stg * BT2_VT_DBL;
            const double* Ts = Vs + SBR_G * BT2_LDZ;
            // ---- X[zc][vc] = sum_r Z[r][zc] V[r][vc]:  16 tiles (zc tile, vc tile), 2 per warp, K = 96
            {
                const int zt = warp & 3, vt0 = (warp >> 2) * 2;
                double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
                const double* za = Zs + (8 * zt + fr) * BT2_LDZ + fk;
                const double* vb = Vs + (8 * vt0 + fr) * BT2_LDZ + fk;
#pragma unroll 4
                for (int q = 0; q < SBR_VROWS / 4; ++q) {
                    const double a = za[4 * q];
                    dmma(acc[0][0], acc[0][1], a, vb[4 * q]);
                    dmma(acc[1][0], acc[1][1], a, vb[8 * BT2_LDZ + 4 * q]);
                }
                // lane holds D[zc = 8 zt + fr][vc = 8 vt + 2 fk + {0, 1}]
                *(double2*)(Xs + (8 * zt + fr) * BT2_LDX + 8 * vt0 + 2 * fk) = make_double2(acc[0][0], acc[0][1]);
                *(double2*)(Xs + (8 * zt + fr) * BT2_LDX + 8 * (vt0 + 1) + 2 * fk) = make_double2(acc[1][0], acc[1][1]);
            }
            __syncthreads();
            // ---- X2[zc][i] = sum_k X[zc][k] T[i][k]   (T upper triangular: k >= i)
            {
                const int zt = warp & 3, it0 = (warp >> 2) * 2;
                double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
                const double* xa = Xs + (8 * zt + fr) * BT2_LDX + fk;
#pragma unroll
                for (int q = 0; q < SBR_G / 4; ++q) {
                    const double a = xa[4 * q];
                    // B[k][i] = T(i, k) = Ts[k * LDX + i]
                    dmma(acc[0][0], acc[0][1], a, Ts[(4 * q + fk) * BT2_LDX + 8 * it0 + fr]);
                    dmma(acc[1][0], acc[1][1], a, Ts[(4 * q + fk) * BT2_LDX + 8 * (it0 + 1) + fr]);
                }
                *(double2*)(X2 + (8 * zt + fr) * BT2_LDX + 8 * it0 + 2 * fk) = make_double2(acc[0][0], acc[0][1]);
                *(double2*)(X2 + (8 * zt + fr) * BT2_LDX + 8 * (it0 + 1) + 2 * fk) = make_double2(acc[1][0], acc[1][1]);
            }
            __syncthreads();
            // ---- Z[r][zc] -= sum_k V[r][k] X2[zc][k]:  D[zc][r]: 4 x 12 tiles, 6 per warp (one zc tile, 6 row tiles), K = 32
            {
                const int zt = warp & 3, rt0 = (warp >> 2) * 6;
                double acc[6][2];
#pragma unroll
                for (int a = 0; a < 6; ++a) { acc[a][0] = 0.0; acc[a][1] = 0.0; }
                const double* xa = X2 + (8 * zt + fr) * BT2_LDX + fk;
#pragma unroll
                for (int q = 0; q < SBR_G / 4; ++q) {
                    const double a = xa[4 * q];
                    const double* vb = Vs + (4 * q + fk) * BT2_LDZ + 8 * rt0 + fr;          // B[k][r] = V(r, k)
#pragma unroll
                    for (int rt = 0; rt < 6; ++rt) dmma(acc[rt][0], acc[rt][1], a, vb[8 * rt]);
                }
                // lane holds D[zc = 8 zt + fr][r = 8 (rt0 + rt) + 2 fk + {0, 1}]
                const int zc = 8 * zt + fr;
#pragma unroll
                for (int rt = 0; rt < 6; ++rt) {
                    const int r = 8 * (rt0 + rt) + 2 * fk, gr = first + r;
                    const double z0 = Zs[zc * BT2_LDZ + r] - acc[rt][0], z1 = Zs[zc * BT2_LDZ + r + 1] - acc[rt][1];
                    if (gr < np) Z[(size_t)zc * np + gr] = z0;
                    if (gr + 1 < np) Z[(size_t)zc * np + gr + 1] = z1;
                }
            }
            __syncthreads();                                              // Zs, Xs and stage `stg` are free again
            stg ^= 1;
            G = Gn; t = tn;
            while (G >= 0 && ntasks(G) == 0) --G;                         // (cannot happen below the first non-empty group)
        }
        cp_async_wait<0>();
    }
}
