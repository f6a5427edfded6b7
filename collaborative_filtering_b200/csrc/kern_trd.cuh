// Large path, stage 1: blocked Householder tridiagonalisation  A = Q T Q^T  (the first half of what
// Eigen's SelfAdjointEigenSolver does at precompute_local.cpp:231), one persistent kernel.
//
// A TEAM of T co-resident CTAs (T = 1 .. all SMs) works on one user at a time and pulls users from a
// queue.  The matrix is stored full (np x np, column-major, np = 64-padded) but only the 64 x 64
// tiles on or below the diagonal are ever read or updated -- half the HBM bytes of a full symv.
// Per column j of a 64-wide panel (LAPACK dlatrd scheme, updates deferred to the end of the panel):
//
//   phase A   acol = A[:,j] - V W[j,:]^T - W V[j,:]^T           rows are owned cyclically (256-row
//             partial ||acol[j+2:]||^2, partial W^T acol, V^T acol   blocks) by the team's CTAs
//   -- team barrier 1 --
//   phase B   beta, tau, v = scale * x'  (x' = acol with x'[j+1] = alpha - beta)
//             y = A x' over this CTA's tiles: 64x64 tiles streamed by the bulk-copy engine
//             (cp.async.bulk + mbarrier, 4 stages), both A_IJ x_J and A_IJ^T x_I per tile,
//             accumulated in shared memory, written once per step as this CTA's partial
//   -- team barrier 2 --
//   phase C   p = A v - V (W^T v) - W (V^T v),  w = tau p - (tau^2/2)(p^T v) v;  then phase A of j+1
//
// and per panel  A22 -= V W^T + W V^T  on the tiles below the panel (FP64 tensor cores, m8n8k4).
// Every reduction has a fixed order, so results do not depend on timing.
#pragma once
#include "gsi_internal.cuh"
#include "ptx.cuh"

#define HH_TS 64          // tile edge
#define HH_NB 64          // panel width
#define TRD_THREADS 256
#define TRD_MAXR 4        // 256-row blocks a CTA may own: np <= 1024 * T
#define TRD_PART 136      // doubles per CTA: [0] norm^2, [1] x'Ax', [2..66) W^T acol, [66..130) V^T acol
#define TRD_STAGE_DBL (HH_TS * HH_TS + 2 * HH_TS)
#define TRD_SYR_LD 68     // k-stride of the syr2k operand panels in shared memory (conflict-free DMMA fragments)
#define TRD_SYR_DBL (4 * TRD_SYR_LD * HH_NB)

// One user of the Householder / divide-and-conquer path
struct HJob {
    int n, np;            // np = 64-padded n = leading dimension of every n x n buffer of the user
    int levels;           // D&C merge levels (0: a single leaf)
    int pad_;
    int64_t m_off;        // offset (doubles) of the user's np*np block in A / Qa / Qb / S
    int64_t r_off;        // offset of the user's np-long vectors (d, e, tau, lam, ...)
    int64_t item_off;     // into items / sig_min
    int64_t vec_off;      // into vec_pad (n * max(n,2))
    int64_t lam_off;      // into lam_pad
};

struct TrdParams {
    const HJob* jobs;
    int njobs;
    int* queue;           // job counter (zeroed before launch)
    double* A;
    double* d; double* e; double* tau;      // indexed by r_off
    int T, npmax, stages;
    double* acol;         // [teams][npmax]
    double* ypart;        // [teams][T][npmax]
    double* part;         // [teams][T][TRD_PART]
    double* tot;          // [teams][2*HH_NB]
    double* Vp; double* Wp;                 // [teams][npmax*HH_NB], leading dimension = the job's np
    unsigned* bar;        // [teams] monotonic arrival counters (zeroed before launch)
    int* slot;            // [teams] job broadcast
};

static inline size_t trd_smem_bytes(int npmax, int stages) {
    const size_t region = (size_t)std::max(stages * TRD_STAGE_DBL, TRD_SYR_DBL);
    return (region + npmax + TRD_MAXR * 256 + 4 * 64 + 512 + 16) * sizeof(double) + 8 * sizeof(uint64_t) + 128;
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sum over the 256 threads of the CTA; result valid in every thread.  `buf` holds >= 8 doubles.
__device__ __forceinline__ double cta_sum_d(double v, double* buf) {
    v = warp_sum_d(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) buf[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < TRD_THREADS / 32; ++w) s += buf[w];
    return s;
}

__device__ __forceinline__ void team_barrier(unsigned* bar, unsigned& target, int T) {
    __syncthreads();
    if (T > 1 && threadIdx.x == 0) {
        target += (unsigned)T;
        __threadfence();
        red_release_add_u32(bar, 1u);
        while ((int)(ld_acquire_u32(bar) - target) < 0) {}
        __threadfence();
    }
    __syncthreads();
}

// walks the tiles (I >= J >= J0) of the lower triangle in column-major tile order, T apart
struct TileWalk {
    int I, J, NT, T;
    __device__ __forceinline__ void norm() {
        while (J < NT && I >= NT) { const int ov = I - NT; ++J; I = J + ov; }
    }
    __device__ __forceinline__ void init(int J0, int NT_, int c, int T_) { NT = NT_; T = T_; J = J0; I = J0 + c; norm(); }
    __device__ __forceinline__ bool valid() const { return J < NT; }
    __device__ __forceinline__ void next() { I += T; norm(); }
};

__global__ void __launch_bounds__(TRD_THREADS, 1) trd_kernel(TrdParams P) {
    extern __shared__ __align__(128) unsigned char trd_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = P.T, team = blockIdx.x / T, c = blockIdx.x % T;
    const int npmax = P.npmax, stages = P.stages;

    double* stage_base = (double*)trd_smem;
    const int region = max(stages * TRD_STAGE_DBL, TRD_SYR_DBL);
    double* ysm = stage_base + region;             // [npmax]
    double* anext = ysm + npmax;                   // [TRD_MAXR*256]
    double* Wtv = anext + TRD_MAXR * 256;          // [64]
    double* Vtv = Wtv + 64;
    double* Wrow = Vtv + 64;
    double* Vrow = Wrow + 64;
    double* red = Vrow + 64;                       // [512]
    double* sc = red + 512;                        // [16] scalars
    uint64_t* full = (uint64_t*)(sc + 16);         // [8]

    double* acol = P.acol + (size_t)team * npmax;
    double* ypart = P.ypart + (size_t)team * T * npmax;
    double* part = P.part + (size_t)team * T * TRD_PART;
    double* tot = P.tot + (size_t)team * 2 * HH_NB;
    double* Vp = P.Vp + (size_t)team * npmax * HH_NB;
    double* Wp = P.Wp + (size_t)team * npmax * HH_NB;
    unsigned* bar = P.bar + team;
    unsigned bar_target = 0;
    unsigned seq = 0;                              // tiles streamed so far (mbarrier phase bookkeeping)

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    for (;;) {
        if (c == 0 && tid == 0) P.slot[team] = atomicAdd(P.queue, 1);
        team_barrier(bar, bar_target, T);
        const int job = __ldcg(P.slot + team);
        if (job >= P.njobs) break;
        const HJob jb = P.jobs[job];
        const int n = jb.n, np = jb.np, ld = jb.np, NT = jb.np >> 6;
        double* A = P.A + jb.m_off;
        double* dvec = P.d + jb.r_off;
        double* evec = P.e + jb.r_off;
        double* tvec = P.tau + jb.r_off;
        const int R = (np + 256 * T - 1) / (256 * T);
        double a_own[TRD_MAXR];
        for (int i = tid; i < np; i += TRD_THREADS) ysm[i] = 0.0;

        auto issue_tile = [&](int I, int J, unsigned sq) {     // executed by warp 0
            const int st = sq % stages;
            double* tile = stage_base + (size_t)st * TRD_STAGE_DBL;
            if (lane == 0) mbar_expect_tx(&full[st], (HH_TS * HH_TS + 2 * HH_TS) * 8);
            __syncwarp();
            const double* src = A + (size_t)(J * HH_TS) * ld + I * HH_TS;
            bulk_g2s(tile + lane * HH_TS, src + (size_t)lane * ld, HH_TS * 8, &full[st]);
            bulk_g2s(tile + (lane + 32) * HH_TS, src + (size_t)(lane + 32) * ld, HH_TS * 8, &full[st]);
            if (lane == 0) bulk_g2s(tile + HH_TS * HH_TS, acol + I * HH_TS, HH_TS * 8, &full[st]);
            if (lane == 1) bulk_g2s(tile + HH_TS * HH_TS + HH_TS, acol + J * HH_TS, HH_TS * 8, &full[st]);
        };

        for (int j0 = 0; j0 < n - 1; j0 += HH_NB) {
            const int pw = min(HH_NB, n - 1 - j0);
            // ---- phase A0: the matrix is up to date, acol = A[:, j0]
            {
                double nrm = 0.0;
#pragma unroll
                for (int s = 0; s < TRD_MAXR; ++s) {
                    const int r = (c + s * T) * 256 + tid;
                    if (s < R && r < np) {
                        const double a = (r >= j0) ? __ldcg(A + (size_t)j0 * ld + r) : 0.0;
                        a_own[s] = a;
                        acol[r] = a;
                        if (r >= j0 + 2) nrm = fma(a, a, nrm);
                    }
                }
                nrm = cta_sum_d(nrm, red);
                if (tid == 0) part[c * TRD_PART] = nrm;
            }
            for (int jj = 0; jj < pw; ++jj) {
                const int j = j0 + jj;
                team_barrier(bar, bar_target, T);                                   // ---- barrier 1
                // ---- scalars of the reflector (identical arithmetic in every CTA)
                if (warp == 0) {
                    double s = 0.0;
                    for (int cc = lane; cc < T; cc += 32) s += __ldcg(part + cc * TRD_PART);
                    s = warp_sum_d(s);
                    if (lane == 0) {
                        const double alpha = __ldcg(acol + j + 1);
                        double beta, tj, scale;
                        if (s == 0.0) { beta = alpha; tj = 0.0; scale = 0.0; }
                        else {
                            beta = -copysign(sqrt(fma(alpha, alpha, s)), alpha);
                            tj = (beta - alpha) / beta;
                            scale = 1.0 / (alpha - beta);
                        }
                        sc[0] = beta; sc[1] = tj; sc[2] = scale; sc[3] = alpha - beta;
                        if (c == 0) { dvec[j] = __ldcg(acol + j); evec[j] = beta; tvec[j] = tj; }
                    }
                }
                // ---- totals of the panel dot products, spread over the team
                for (int el = c + T * warp; el < 2 * jj; el += T * (TRD_THREADS / 32)) {
                    const int idx = (el < jj) ? 2 + el : 2 + HH_NB + (el - jj);
                    double s = 0.0;
                    for (int cc = lane; cc < T; cc += 32) s += __ldcg(part + cc * TRD_PART + idx);
                    s = warp_sum_d(s);
                    if (lane == 0) tot[idx - 2] = s;
                }
                __syncthreads();
                const double xfix = sc[3];
                // ---- phase B: y = A x' over my tiles
                const int J0 = (j + 1) >> 6;
                {
                    const int m = NT - J0, total = m * (m + 1) / 2;
                    const int mine = (c < total) ? (total - c + T - 1) / T : 0;
                    TileWalk wi, wc;
                    wi.init(J0, NT, c, T);
                    wc = wi;
                    int issued = 0;
                    if (warp == 0) fence_proxy_async();
                    for (; issued < min(stages, mine); ++issued) {
                        if (warp == 0) issue_tile(wi.I, wi.J, seq + issued);
                        wi.next();
                    }
                    double xax = 0.0;
                    const int i = tid & 63, q = tid >> 6;
                    for (int k = 0; k < mine; ++k) {
                        const unsigned sq = seq + k;
                        const int st = sq % stages;
                        mbar_wait(&full[st], (sq / stages) & 1);
                        double* tile = stage_base + (size_t)st * TRD_STAGE_DBL;
                        double* xI = tile + HH_TS * HH_TS;
                        double* xJ = xI + HH_TS;
                        const int I = wc.I, J = wc.J;
                        if (tid < 128) {
                            const int gi = (q ? J : I) * HH_TS + i;
                            double* xp = q ? xJ : xI;
                            if (gi <= j) xp[i] = 0.0;
                            else if (gi == j + 1) xp[i] = xfix;
                        }
                        __syncthreads();
                        {
                            double acc = 0.0;
#pragma unroll
                            for (int cc = 0; cc < 16; ++cc) acc = fma(tile[(q * 16 + cc) * HH_TS + i], xJ[q * 16 + cc], acc);
                            red[q * 64 + i] = acc;
                        }
                        if (I != J) {
                            double acc = 0.0;
#pragma unroll
                            for (int kk = 0; kk < 16; ++kk) {
                                const int r = q * 16 + ((kk + i) & 15);
                                acc = fma(tile[i * HH_TS + r], xI[r], acc);
                            }
                            red[256 + q * 64 + i] = acc;
                        }
                        __syncthreads();
                        if (tid < 64) {
                            const double y = (red[i] + red[64 + i]) + (red[128 + i] + red[192 + i]);
                            ysm[I * HH_TS + i] += y;
                            xax = fma(xI[i] * y, (I != J) ? 2.0 : 1.0, xax);
                        } else if (tid < 128 && I != J) {
                            const double y = (red[256 + i] + red[320 + i]) + (red[384 + i] + red[448 + i]);
                            ysm[J * HH_TS + i] += y;
                        }
                        __syncthreads();
                        if (issued < mine) {
                            if (warp == 0) { fence_proxy_async(); issue_tile(wi.I, wi.J, seq + issued); }
                            wi.next();
                            ++issued;
                        }
                        wc.next();
                    }
                    seq += mine;
                    for (int r = J0 * HH_TS + tid; r < np; r += TRD_THREADS) {
                        ypart[(size_t)c * npmax + r] = ysm[r];
                        ysm[r] = 0.0;
                    }
                    xax = cta_sum_d(xax, red);
                    if (tid == 0) part[c * TRD_PART + 1] = xax;
                }
                team_barrier(bar, bar_target, T);                                   // ---- barrier 2
                // ---- phase C
                const double beta = sc[0], tj = sc[1], scale = sc[2];
                if (tid < jj) {
                    const double wr = __ldcg(Wp + (size_t)tid * ld + j + 1), vr = __ldcg(Vp + (size_t)tid * ld + j + 1);
                    Wrow[tid] = wr; Vrow[tid] = vr;
                    Wtv[tid] = scale * (__ldcg(tot + tid) - beta * wr);
                    Vtv[tid] = scale * (__ldcg(tot + HH_NB + tid) - beta * vr);
                }
                if (warp == 2) {
                    double s = 0.0, y = 0.0;
                    for (int cc = lane; cc < T; cc += 32) {
                        s += __ldcg(part + cc * TRD_PART + 1);
                        y += __ldcg(ypart + (size_t)cc * npmax + j + 1);
                    }
                    s = warp_sum_d(s); y = warp_sum_d(y);
                    if (lane == 0) { sc[4] = s; sc[5] = y; }
                }
                __syncthreads();
                if (warp == 0) {
                    double vtw = 0.0, rowdot = 0.0;
                    for (int cc = lane; cc < jj; cc += 32) {
                        vtw = fma(Vtv[cc], Wtv[cc], vtw);
                        rowdot += Vrow[cc] * Wtv[cc] + Wrow[cc] * Vtv[cc];
                    }
                    vtw = warp_sum_d(vtw); rowdot = warp_sum_d(rowdot);
                    if (lane == 0) {
                        const double pv = scale * scale * sc[4] - 2.0 * vtw;
                        const double alpha2 = -0.5 * tj * tj * pv;
                        sc[6] = alpha2;
                        sc[7] = tj * (scale * sc[5] - rowdot) + alpha2;      // w[j+1]
                    }
                }
                __syncthreads();
                const double alpha2 = sc[6], wj1 = sc[7];
                const bool more = jj + 1 < pw;
                double nrm = 0.0;
#pragma unroll
                for (int s = 0; s < TRD_MAXR; ++s) {
                    const int r = (c + s * T) * 256 + tid;
                    if (s < R && r < np) {
                        if (r <= j) { A[(size_t)j * ld + r] = 0.0; anext[s * 256 + tid] = 0.0; }
                        else {
                            double av = 0.0;
                            for (int cc = 0; cc < T; ++cc) av += __ldcg(ypart + (size_t)cc * npmax + r);
                            av *= scale;
                            const double vr = (r == j + 1) ? 1.0 : scale * a_own[s];
                            double pdot = 0.0, adot = 0.0;
                            for (int cc = 0; cc < jj; ++cc) {
                                const double vv = __ldcg(Vp + (size_t)cc * ld + r), ww = __ldcg(Wp + (size_t)cc * ld + r);
                                pdot = fma(vv, Wtv[cc], fma(ww, Vtv[cc], pdot));
                                adot = fma(vv, Wrow[cc], fma(ww, Vrow[cc], adot));
                            }
                            const double wr = tj * (av - pdot) + alpha2 * vr;
                            Vp[(size_t)jj * ld + r] = vr;
                            Wp[(size_t)jj * ld + r] = wr;
                            A[(size_t)j * ld + r] = vr;
                            if (more) {
                                const double an = __ldcg(A + (size_t)(j + 1) * ld + r) - adot - vr * wj1 - wr;
                                a_own[s] = an;
                                acol[r] = an;
                                anext[s * 256 + tid] = an;
                                if (r >= j + 3) nrm = fma(an, an, nrm);
                            }
                        }
                    }
                }
                if (more) {
                    nrm = cta_sum_d(nrm, red);          // (its __syncthreads also publish anext and the new panel column)
                    if (tid == 0) part[c * TRD_PART] = nrm;
                    for (int qd = warp; qd < 2 * (jj + 1); qd += TRD_THREADS / 32) {
                        const int which = qd > jj, cc = which ? qd - (jj + 1) : qd;
                        const double* Pn = (which ? Vp : Wp) + (size_t)cc * ld;
                        double acc = 0.0;
                        for (int s = 0; s < R; ++s) {
                            const int base = (c + s * T) * 256;
                            for (int ii = lane; ii < 256; ii += 32) {
                                const int r = base + ii;
                                if (r < np && r >= j + 2) acc = fma(__ldcg(Pn + r), anext[s * 256 + ii], acc);
                            }
                        }
                        acc = warp_sum_d(acc);
                        if (lane == 0) part[c * TRD_PART + 2 + which * HH_NB + cc] = acc;
                    }
                }
            }
            // ---- trailing update  A22 -= V W^T + W V^T  on the tiles at or below (jn, jn)
            team_barrier(bar, bar_target, T);
            {
                const int jn = j0 + pw, kpad = (pw + 3) & ~3;
                double* S = stage_base;
                TileWalk w;
                for (w.init(jn >> 6, NT, c, T); w.valid(); w.next()) {
                    const int I = w.I, J = w.J;
                    for (int idx = tid; idx < 4 * 64 * kpad; idx += TRD_THREADS) {
                        const int which = idx / (64 * kpad), rem = idx - which * 64 * kpad, k = rem >> 6, r = rem & 63;
                        const double* src = ((which & 1) ? Wp : Vp) + (size_t)k * ld + ((which >> 1) ? J : I) * HH_TS + r;
                        S[which * (TRD_SYR_LD * HH_NB) + k * TRD_SYR_LD + r] = (k < pw) ? __ldcg(src) : 0.0;
                    }
                    __syncthreads();
                    const double* VIs = S, *WIs = S + TRD_SYR_LD * HH_NB, *VJs = S + 2 * TRD_SYR_LD * HH_NB, *WJs = S + 3 * TRD_SYR_LD * HH_NB;
                    double acc[8][2];
#pragma unroll
                    for (int rb = 0; rb < 8; ++rb) { acc[rb][0] = 0.0; acc[rb][1] = 0.0; }
                    const int fk = lane & 3, fr = lane >> 2;
                    for (int k0 = 0; k0 < kpad; k0 += 4) {
                        const double aW = WJs[(k0 + fk) * TRD_SYR_LD + 8 * warp + fr];
                        const double aV = VJs[(k0 + fk) * TRD_SYR_LD + 8 * warp + fr];
#pragma unroll
                        for (int rb = 0; rb < 8; ++rb) {
                            const double bV = VIs[(k0 + fk) * TRD_SYR_LD + 8 * rb + fr];
                            const double bW = WIs[(k0 + fk) * TRD_SYR_LD + 8 * rb + fr];
                            dmma(acc[rb][0], acc[rb][1], aW, bV);
                            dmma(acc[rb][0], acc[rb][1], aV, bW);
                        }
                    }
                    const int gc = J * HH_TS + 8 * warp + fr;
                    if (gc >= jn) {
#pragma unroll
                        for (int rb = 0; rb < 8; ++rb) {
                            const int gr = I * HH_TS + 8 * rb + 2 * fk;
                            double2* p = (double2*)(A + (size_t)gc * ld + gr);
                            double2 v = __ldcg(p);
                            if (gr >= jn) v.x -= acc[rb][0];
                            if (gr + 1 >= jn) v.y -= acc[rb][1];
                            *p = v;
                        }
                    }
                    __syncthreads();
                }
            }
            team_barrier(bar, bar_target, T);
        }
        if (c == 0 && tid == 0) {
            dvec[n - 1] = __ldcg(A + (size_t)(n - 1) * ld + (n - 1));
            if (n >= 2) { /* e[n-1], tau[n-1] unused */ }
        }
        // the next job's first barrier separates this job's scratch use from the next one's
    }
}
