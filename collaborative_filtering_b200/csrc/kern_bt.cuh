// Large path, stage 3: back-transformation  U = H_0 H_1 ... H_{n-2} Z  of the kept eigenvectors of T
// (Eigen does this inside SelfAdjointEigenSolver, precompute_local.cpp:231) and the record emit
// (cutoff :252-261, row-major n x k as the text writer :263-280 walks it).
//
//   bt_formt_kernel   per 64-reflector panel: G = V^T V, the compact-WY T (dlarft, forward/columnwise),
//                     and VT = V T stored in the (now free) S buffer
//   bt_apply_kernel   persistent; a work item is a 32-column block of Z of one user.  The columns of Z
//                     are independent, so a CTA walks all panels (last to first) of its block with no
//                     inter-CTA synchronisation:   X = V^T Z_blk ;  Z_blk -= VT X    (DMMA m8n8k4,
//                     operands staged with cp.async, double buffered)
//   emit_sign_kernel / emit_vec_kernel   sign convention (largest |component| positive) + transpose
#pragma once
#include "gsi_internal.cuh"
#include "kern_trd.cuh"
#include "ptx.cuh"

#define BT_NB 64
#define BT_CB 32          // columns of Z per work item
#define BT_LD 68          // shared-memory leading dimension (conflict-free DMMA fragments, 16-byte aligned)

struct BtParams {
    const HJob* jobs;
    int njobs;
    double* A;            // reflectors, tile-major: column j holds v_j (1 at row j+1, zeros above within the panel)
    const double* tau;    // [r_off + j]
    double* S;            // VT (same layout as A)
    double* Qa; double* Qb;   // Z lives in Qb when the user's level count is odd, else Qa
    const int32_t* kuser;     // [jobs] kept columns
    // work list
    const int2* items;    // (job, column block), sorted by cost descending
    int nitems;
    int* queue;
    const uint8_t* skip;  // optional [jobs]: 1 = this user's back-transform ran on the tensor-core engine (kern_bt_tc.cuh)
};

// grid (max panels, njobs), block 256, dynamic smem 3 * 64*65 doubles
static inline size_t bt_formt_smem_bytes() { return (size_t)3 * 64 * 65 * sizeof(double); }
__global__ void __launch_bounds__(256) bt_formt_kernel(BtParams P) {
    extern __shared__ __align__(16) double formt_smem[];
    if (P.skip && P.skip[blockIdx.y]) return;
    const HJob jb = P.jobs[blockIdx.y];
    const int n = jb.n, np = jb.np, ld = jb.np;
    const int j0 = blockIdx.x * BT_NB;
    if (j0 >= n - 1) return;
    const int b = min(BT_NB, n - 1 - j0);
    const int NT = np >> 6;
    const double* Auser = P.A + jb.m_off;                     // tile-major: rows r0.., panel columns = tile (r0/64, j0/64)
    double* VT = P.S + jb.m_off + (size_t)j0 * ld;            // column-major
    const double* tau = P.tau + jb.r_off + j0;
    double* Vs = formt_smem;            // chunk: 64 rows x 64 cols, [row][col] padded to 65
    double* G = Vs + 64 * 65;
    double* Tm = G + 64 * 65;
    const int tid = threadIdx.x;
    const int tr = tid >> 4, tc = tid & 15;                    // thread owns rows 4tr.., cols 4tc.. of a 64x64 result
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = 0.0;
    for (int r0 = j0; r0 < np; r0 += 64) {
        __syncthreads();
        for (int e = tid; e < 64 * 64; e += 256) {
            const int c = e >> 6, r = e & 63;
            Vs[r * 65 + c] = (c < b) ? Auser[(((size_t)(j0 >> 6) * NT + (r0 >> 6)) << 12) + (c << 6) + r] : 0.0;
        }
        __syncthreads();
        for (int r = 0; r < 64; ++r) {
            double x[4], y[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) { x[a] = Vs[r * 65 + 4 * tr + a]; y[a] = Vs[r * 65 + 4 * tc + a]; }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[a][c] = fma(x[a], y[c], acc[a][c]);
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) { G[(4 * tr + a) * 65 + 4 * tc + c] = acc[a][c]; Tm[(4 * tr + a) * 65 + 4 * tc + c] = 0.0; }
    __syncthreads();
    // T(j,j) = tau_j ;  T(0:j, j) = -tau_j * T(0:j, 0:j) * G(0:j, j)
    for (int j = 0; j < b; ++j) {
        const double tj = tau[j];
        if (tid < j) {
            double s = 0.0;
            for (int l = tid; l < j; ++l) s = fma(Tm[tid * 65 + l], G[l * 65 + j], s);
            Tm[tid * 65 + j] = -tj * s;
        } else if (tid == j) Tm[j * 65 + j] = tj;
        __syncthreads();
    }
    // VT = V T, 64-row chunks (T upper triangular, zero elsewhere)
    for (int r0 = j0; r0 < np; r0 += 64) {
        __syncthreads();
        for (int e = tid; e < 64 * 64; e += 256) {
            const int c = e >> 6, r = e & 63;
            Vs[r * 65 + c] = (c < b) ? Auser[(((size_t)(j0 >> 6) * NT + (r0 >> 6)) << 12) + (c << 6) + r] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][c] = 0.0;
        for (int l = 0; l < b; ++l) {
            double x[4], y[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) { x[a] = Vs[(4 * tr + a) * 65 + l]; y[a] = Tm[l * 65 + 4 * tc + a]; }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[a][c] = fma(x[a], y[c], acc[a][c]);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (4 * tc + c < b) {
#pragma unroll
                for (int a = 0; a < 4; ++a) VT[(size_t)(4 * tc + c) * ld + r0 + 4 * tr + a] = acc[a][c];
            }
    }
}

// smem: 3 stages x (V chunk 64x64 + Z chunk 64x32) and X (64 x 32); the 4 k-slice partials of X overlay the
// (drained) stages between the two phases
#define BT_STAGE_DBL (BT_NB * BT_LD + BT_CB * BT_LD)
#define BT_STAGES 3
static inline size_t bt_smem_bytes() { return (size_t)(BT_STAGES * BT_STAGE_DBL + BT_CB * BT_LD) * sizeof(double); }

// Register blocking (the m8n8k4 fragments are small, so shared-memory loads per DMMA decide the rate):
//   phase 1  X = V^T Z over a 64-row chunk: warp (ks, vh) takes the 16 rows ks of the chunk and the 32
//            reflector columns vh -> 4 x 4 tiles, 8 fragment loads per 16 DMMA; the 4 row slices are
//            summed through shared memory once per panel
//   phase 2  Z -= VT X: the whole X sits in registers as A fragments (64 doubles per lane), a warp owns
//            8 rows of every chunk -> 1 fragment load per 4 DMMA
__global__ void __launch_bounds__(256, 1) bt_apply_kernel(BtParams P) {
    extern __shared__ __align__(16) double bt_smem[];
    __shared__ int item_s;
    double* Xs = bt_smem + BT_STAGES * BT_STAGE_DBL;   // [zcol][k] , k = reflector index within the panel
    double* Xpart = bt_smem;                           // [4][zcol][k], overlays the stages (4 * 32 * 68 <= 2 stages)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fk = lane & 3, fr = lane >> 2;
    const int ks = warp & 3, vh = warp >> 2;
    for (;;) {
        __syncthreads();
        if (tid == 0) item_s = atomicAdd(P.queue, 1);
        __syncthreads();
        const int it = item_s;
        if (it >= P.nitems) break;
        const int2 item = P.items[it];
        if (P.skip && P.skip[item.x]) continue;
        const HJob jb = P.jobs[item.x];
        const int lim = P.kuser[item.x];
        const int c0 = item.y * BT_CB;
        if (c0 >= lim) continue;
        const int n = jb.n, np = jb.np, ld = jb.np;
        double* Z = ((jb.levels & 1) ? P.Qb : P.Qa) + jb.m_off + (size_t)c0 * ld;
        const int npanels = (n - 1 + BT_NB - 1) / BT_NB;
        for (int p = npanels - 1; p >= 0; --p) {
            const int j0 = p * BT_NB;
            const int NT = np >> 6;
            const double* Vt = P.A + jb.m_off + (((size_t)(j0 >> 6) * NT) << 12);   // tile column of the panel (tile-major A)
            const double* VT = P.S + jb.m_off + (size_t)j0 * ld;
            const int nch = (np - j0) / 64, b = min(BT_NB, n - 1 - j0);
            // ---------------- phase 1: X = V^T Z  (64 x 32), K = rows j0 .. np
            auto load1 = [&](int ch, int stg) {
                double* Vs = bt_smem + (size_t)stg * BT_STAGE_DBL;
                double* Zs = Vs + BT_NB * BT_LD;
                const int r0 = j0 + ch * 64;
#pragma unroll
                for (int t = 0; t < 8; ++t) {           // V chunk: 64 cols x 32 16-byte chunks
                    const int e = tid + t * 256, c = e >> 5, r2 = (e & 31) * 2;
                    cp_async16_zfill(Vs + c * BT_LD + r2, Vt + ((size_t)(r0 >> 6) << 12) + (c << 6) + r2, c < b);
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) {           // Z chunk: 32 cols x 32 chunks
                    const int e = tid + t * 256, c = e >> 5, r2 = (e & 31) * 2;
                    cp_async16(Zs + c * BT_LD + r2, Z + (size_t)c * ld + r0 + r2);
                }
            };
            double acc[4][4][2];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int z = 0; z < 4; ++z) { acc[a][z][0] = 0.0; acc[a][z][1] = 0.0; }
            __syncthreads();
            load1(0, 0);
            cp_async_commit();
            if (nch > 1) load1(1, 1);
            cp_async_commit();
            for (int ch = 0; ch < nch; ++ch) {
                cp_async_wait<1>();
                __syncthreads();                        // chunk ch has landed; everyone is done with chunk ch - 1
                if (ch + 2 < nch) load1(ch + 2, (ch + 2) % BT_STAGES);
                cp_async_commit();
                const double* Vs = bt_smem + (size_t)(ch % BT_STAGES) * BT_STAGE_DBL;
                const double* Zs = Vs + BT_NB * BT_LD;
                // D[vcol][zcol] += V[k][vcol] * Z[k][zcol] over my 16 rows of the chunk
#pragma unroll
                for (int s4 = 0; s4 < 4; ++s4) {
                    const int kr = 16 * ks + 4 * s4 + fk;
                    double af[4], bf[4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) af[a] = Vs[(32 * vh + 8 * a + fr) * BT_LD + kr];
#pragma unroll
                    for (int z = 0; z < 4; ++z) bf[z] = Zs[(8 * z + fr) * BT_LD + kr];
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int z = 0; z < 4; ++z) dmma(acc[a][z][0], acc[a][z][1], af[a], bf[z]);
                }
            }
            cp_async_wait<0>();
            __syncthreads();                            // the stages are drained: the partials may overlay them
            // partial X -> smem [ks][zcol][vcol]: lane holds D[vcol = 32 vh + 8 a + fr][zcol = 8 z + 2 fk + {0,1}]
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int z = 0; z < 4; ++z) {
                    double* xp = Xpart + ks * (BT_CB * BT_LD) + (8 * z + 2 * fk) * BT_LD + 32 * vh + 8 * a + fr;
                    xp[0] = acc[a][z][0];
                    xp[BT_LD] = acc[a][z][1];
                }
            __syncthreads();
            for (int e = tid; e < BT_CB * BT_NB; e += 256) {
                const int zc = e >> 6, v = e & 63, o = zc * BT_LD + v;
                Xs[o] = (Xpart[o] + Xpart[BT_CB * BT_LD + o]) + (Xpart[2 * BT_CB * BT_LD + o] + Xpart[3 * BT_CB * BT_LD + o]);
            }
            __syncthreads();
            // ---------------- phase 2: Z -= VT X, 64-row chunks; X as A fragments in registers
            double xa[4][16];
#pragma unroll
            for (int z = 0; z < 4; ++z)
#pragma unroll
                for (int k4 = 0; k4 < 16; ++k4) xa[z][k4] = Xs[(8 * z + fr) * BT_LD + 4 * k4 + fk];       // X[k][zcol]
            auto load2 = [&](int ch, int stg) {
                double* Ts = bt_smem + (size_t)stg * BT_STAGE_DBL;
                double* Zs = Ts + BT_NB * BT_LD;
                const int r0 = j0 + ch * 64;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int e = tid + t * 256, c = e >> 5, r2 = (e & 31) * 2;
                    cp_async16_zfill(Ts + c * BT_LD + r2, VT + (size_t)c * ld + r0 + r2, c < b);
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) {           // the Z chunk rides along: no exposed global latency in the update
                    const int e = tid + t * 256, c = e >> 5, r2 = (e & 31) * 2;
                    cp_async16(Zs + c * BT_LD + r2, Z + (size_t)c * ld + r0 + r2);
                }
            };
            load2(0, 0);
            cp_async_commit();
            if (nch > 1) load2(1, 1);
            cp_async_commit();
            for (int ch = 0; ch < nch; ++ch) {
                cp_async_wait<1>();
                __syncthreads();
                if (ch + 2 < nch) load2(ch + 2, (ch + 2) % BT_STAGES);
                cp_async_commit();
                const int r0 = j0 + ch * 64;
                double d[4][2];
#pragma unroll
                for (int z = 0; z < 4; ++z) { d[z][0] = 0.0; d[z][1] = 0.0; }
                const double* Ts = bt_smem + (size_t)(ch % BT_STAGES) * BT_STAGE_DBL;
                const double* Zs = Ts + BT_NB * BT_LD;
#pragma unroll
                for (int k4 = 0; k4 < 16; ++k4) {
                    const double bq = Ts[(4 * k4 + fk) * BT_LD + 8 * warp + fr];        // VT[row][k]
#pragma unroll
                    for (int z = 0; z < 4; ++z) dmma(d[z][0], d[z][1], xa[z][k4], bq);
                }
#pragma unroll
                for (int z = 0; z < 4; ++z) {
                    double2 zz = *(const double2*)(Zs + (8 * z + fr) * BT_LD + 8 * warp + 2 * fk);
                    zz.x -= d[z][0]; zz.y -= d[z][1];
                    *(double2*)(Z + (size_t)(8 * z + fr) * ld + r0 + 8 * warp + 2 * fk) = zz;
                }
            }
            cp_async_wait<0>();
        }
    }
}

// grid (ceil(npmax/8), njobs), block 256: warp per kept column -> sign (largest |component| positive,
// first on ties), stored as +-1 in sgn[r_off + col]
__global__ void __launch_bounds__(256) emit_sign_kernel(BtParams P, double* __restrict__ sgn, int job0) {
    const HJob jb = P.jobs[job0 + blockIdx.y];
    const int col = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int lim = P.kuser[job0 + blockIdx.y];
    if (col >= lim || col >= jb.n) return;
    const double* z = ((jb.levels & 1) ? P.Qb : P.Qa) + jb.m_off + (size_t)col * jb.np;
    double best = -1.0;
    int arg = 0;
    for (int i = lane; i < jb.n; i += 32) {
        const double v = fabs(z[i]);
        if (v > best) { best = v; arg = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    if (lane == 0) sgn[jb.r_off + col] = (z[arg] < 0.0) ? -1.0 : 1.0;
}

// grid (tiles_i * tiles_r, 1, njobs), block (32, 8): vec[i*k + r] = sgn[r] * Z[i, r]; lam[r]
__global__ void emit_vec_kernel(BtParams P, const double* __restrict__ sgn, const double* __restrict__ lamA,
                                const double* __restrict__ lamB, double* __restrict__ vec_pad,
                                double* __restrict__ lam_pad, int tiles_r, int job0) {
    __shared__ double tile[32][33];
    const HJob jb = P.jobs[job0 + blockIdx.z];
    const int n = jb.n, k = P.kuser[job0 + blockIdx.z];
    const int ti = blockIdx.x / tiles_r, tr = blockIdx.x % tiles_r;
    if (ti * 32 >= n || tr * 32 >= k) return;
    const double* Z = ((jb.levels & 1) ? P.Qb : P.Qa) + jb.m_off;
    const int ld = jb.np, tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        const int rl = ty + 8 * rr;
        const int r = tr * 32 + rl, i = ti * 32 + tx;
        tile[rl][tx] = (r < k && i < n) ? Z[(size_t)r * ld + i] * sgn[jb.r_off + r] : 0.0;
    }
    __syncthreads();
    double* vec = vec_pad + jb.vec_off;
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        const int il = ty + 8 * rr;
        const int i = ti * 32 + il, r = tr * 32 + tx;
        if (i < n && r < k) vec[(size_t)i * k + r] = tile[tx][il];
    }
    if (ti == 0 && ty == 0) {
        const int r = tr * 32 + tx;
        const double* lam = ((jb.levels & 1) ? lamB : lamA) + jb.r_off;
        if (r < k) lam_pad[jb.lam_off + r] = lam[r];
    }
}
