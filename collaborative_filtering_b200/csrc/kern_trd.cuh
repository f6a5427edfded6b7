// Large path, stage 1: blocked Householder tridiagonalisation  A = Q T Q^T  (the first half of what
// Eigen's SelfAdjointEigenSolver does at precompute_local.cpp:231), one persistent kernel.
//
// A TEAM of T co-resident CTAs (T = 1 .. all SMs) works on one user at a time and pulls users from a
// queue.  The matrix is stored TILE-MAJOR (np x np, np = 64-padded, 64 x 64 column-major tiles, each
// 32 KB contiguous: tile (I, J) at ((J * NT + I) << 12)), so one bulk copy moves a tile; only the tiles
// on or below the diagonal are ever read or updated -- half the HBM bytes of a full symv.
// Per column j of a 64-wide panel (LAPACK dlatrd scheme, updates deferred to the end of the panel):
//
//   phase A   acol = A[:,j] - V W[j,:]^T - W V[j,:]^T           rows are owned cyclically (64-row
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
#define TRD_PART 136      // doubles per CTA: [0] norm^2, [1] x'Ax', [2..66) W^T acol, [66..130) V^T acol
#define TRD_STAGE_DBL (HH_TS * HH_TS)
#define TRD_SYR_LD 68     // k-stride of the syr2k operand panels in shared memory (conflict-free DMMA fragments)
#define TRD_SYR_KH 16     // the trailing update streams its operands in k-chunks of 16, double buffered (cp.async)
#define TRD_SYR_DBL (2 * 4 * TRD_SYR_LD * TRD_SYR_KH)
#define TRD_DSET_LD 66     // direct partials of a tile: [8 warps][64 rows], stride 66 (conflict-free 16-byte stores, 8-byte reduce loads)
#define TRD_TSET_LD 68     // transposed partials of a tile column: [4 lane groups][64 columns], stride 68
#define TRD_DSET_DBL (8 * TRD_DSET_LD)
#define TRD_TSET_DBL (4 * TRD_TSET_LD)
#define TRD_UNION_DBL (2 * TRD_DSET_DBL + 2 * TRD_TSET_DBL)   // symv: two parities of both partial sets; phase C: Wtv/Vtv/Wrow/Vrow + sum scratch
#define TRD_FIXED_DBL (16 + 16 + 16 + TRD_UNION_DBL)          // mbarriers, scalars, tile descriptors of the stages (+ pad), the union above; multiple of 16
#define TRD_MAX_STAGES 6

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

// element (r, c) of a tile-major np x np matrix with NT = np / 64 tiles per dimension
__host__ __device__ __forceinline__ size_t hh_tidx(int r, int c, int NT) {
#ifdef __CUDA_ARCH__
    GSI_BOUNDS(r >= 0 && c >= 0 && (r >> 6) < NT && (c >> 6) < NT);
#endif
    return (((size_t)(c >> 6) * NT + (r >> 6)) << 12) + ((c & 63) << 6) + (r & 63);
}

// One launch walks a few LEVELS: level l works on the users of one size class with teams of T_l CTAs,
// T_0 > T_1 > ... (the last level usually has T = 1).  The teams of level l+1 are subdivisions of the
// teams of level l, so a team that finds its level's queue empty splits and moves on at once -- the
// big users start first on big teams and the small ones fill every SM behind them (no tail per class).
#define TRD_MAX_LEVELS 6
#define TRD_PROF_LVSTAT 200   // trace buffer: [0..12) CTA 0 phase cycles, [15] start, [16..16+SMs) finish, then 3 values per (CTA, level)
struct TrdLevel {
    int T, job0, njobs, npmax;              // team size, job range [job0, job0 + njobs), largest padded size
    int stages, own;                        // tile stages that fit beside this level's vectors; 64-row blocks a CTA can own
    int nb, pad_;                           // panel width of the level (<= HH_NB): small users pay less per column with narrow panels
    double* acol;         // [teams][npmax]
    double* ypart;        // [teams][T][npmax]
    double* part;         // [teams][T][TRD_PART]
    double* tot;          // [teams][2*HH_NB]
    double* Vp; double* Wp;                 // [teams][npmax*HH_NB], leading dimension = the job's np
    unsigned* bar;        // [teams] monotonic arrival counters (zeroed before launch)
    int* slot;            // [teams] job broadcast
    int* queue;           // job counter of the level (zeroed before launch)
};
struct TrdParams {
    const HJob* jobs;
    double* A;
    double* d; double* e; double* tau;      // indexed by r_off
    int nlevels;
    TrdLevel lv[TRD_MAX_LEVELS];
    long long* prof;      // optional [16] cycle counters of CTA 0 (GSI_TRACE), else nullptr
};

// shared memory of a level: fixed part, own rows, y and x' vectors, tile stages (the trailing update's
// operand staging overlays the stages)
static inline size_t trd_smem_bytes(int npmax, int stages, int own) {
    const size_t region = (size_t)std::max(stages * TRD_STAGE_DBL, TRD_SYR_DBL);
    return (TRD_FIXED_DBL + (size_t)own * 64 + 2 * (size_t)npmax + region) * sizeof(double) + 128;
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

#define TRD_PROF(slot_)                                                             \
    do {                                                                            \
        if (prof_on) { const long long t_ = clock64(); prof_acc[slot_] += t_ - prof_t; prof_t = t_; } \
    } while (0)

__global__ void __launch_bounds__(TRD_THREADS, 1) trd_kernel(TrdParams P) {
    extern __shared__ __align__(128) unsigned char trd_smem[];
    const bool prof_on = P.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    if (prof_on) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); P.prof[15] = (long long)t; }
    long long prof_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long prof_t = prof_on ? clock64() : 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // fixed part of the carve; the per-level part follows (sizes differ per level)
    double* sm = (double*)trd_smem;
    uint64_t* full = (uint64_t*)sm;                // [8] tile stages, [8] = the x vector
    uint64_t* xbar = full + 8;
    double* sc = sm + 16;                          // [16] scalars
    int2* desc = (int2*)(sc + 16);                 // [8] (I, J) of the tile in each stage, written by the issuing thread
    double* Wtv = sc + 32;                         // phase C view of the union: 4 x [64] ...
    double* Vtv = Wtv + 64;
    double* Wrow = Vtv + 64;
    double* Vrow = Wrow + 64;
    double* red = Vrow + 64;                       // ... and the scratch of cta_sum_d
    double* dset = Wtv;                            // symv view: [2][8][66] direct partials, [2][4][68] transposed partials
    double* tset = dset + 2 * TRD_DSET_DBL;
    double* lvl_base = Wtv + TRD_UNION_DBL;
    const uint32_t full_u32 = smem_u32(full), xbar_u32 = smem_u32(xbar);
    unsigned xphase = 0, phbits = 0;               // parity to wait for next, per mbarrier
    int c_st = 0, p_st = 0, pre_issued = 0;        // stage ring: next stage to consume / to fill; tiles requested ahead

    if (tid == 0) {
        for (int s = 0; s < TRD_MAX_STAGES; ++s) mbar_init(&full[s], 1);
        mbar_init(xbar, 1);
        mbar_fence_init();
    }
    __syncthreads();

    int rank = blockIdx.x, team = 0, size_prev = gridDim.x;
    bool nested = true;                            // still inside the nested team structure
    for (int lvi = 0; lvi < P.nlevels; ++lvi) {
    const TrdLevel& LV = P.lv[lvi];
    const int T = LV.T;
    if (T == 1) { team = blockIdx.x; rank = 0; }   // every CTA is its own team
    else {
        if (!nested) continue;
        const int nsub = size_prev / T, sub = rank / T;
        if (sub >= nsub) { nested = false; continue; }     // left over by the subdivision: joins again at T = 1
        team = team * nsub + sub; rank = rank % T; size_prev = T;
    }
    const int c = rank, lnp = LV.npmax, stages = LV.stages;
    double* aown = lvl_base;                       // [own*64] acol of the rows this CTA owns
    double* ysm = aown + LV.own * 64;              // [lnp]
    double* xsm = ysm + lnp;                       // [lnp] acol of the step (bulk-copied once per step)
    double* stage_base = xsm + lnp;                // 128-byte aligned: every size above is a multiple of 16 doubles
    c_st = 0; p_st = 0;                            // every stage is drained between levels
    double* acol = LV.acol + (size_t)team * lnp;
    double* ypart = LV.ypart + (size_t)team * T * lnp;
    double* part = LV.part + (size_t)team * T * TRD_PART;
    double* tot = LV.tot + (size_t)team * 2 * HH_NB;
    double* Vp = LV.Vp + (size_t)team * lnp * HH_NB;
    double* Wp = LV.Wp + (size_t)team * lnp * HH_NB;
    unsigned* bar = LV.bar + team;
    unsigned bar_target = 0;
    long long* lvstat = (long long*)(sc + 10);     // trace only: [0] symv cycles, [1] tiles, [2] level start
    if (P.prof != nullptr && tid == 0) { lvstat[0] = 0; lvstat[1] = 0; lvstat[2] = clock64(); }
    for (;;) {
        if (c == 0 && tid == 0) LV.slot[team] = atomicAdd(LV.queue, 1);
        team_barrier(bar, bar_target, T);
        const int job = __ldcg(LV.slot + team);
        if (job >= LV.njobs) break;
        const HJob jb = P.jobs[LV.job0 + job];
        const int n = jb.n, np = jb.np, ld = jb.np, NT = jb.np >> 6;
        double* A = P.A + jb.m_off;
        double* dvec = P.d + jb.r_off;
        double* evec = P.e + jb.r_off;
        double* tvec = P.tau + jb.r_off;
        const int nown = (NT > c) ? (NT - c + T - 1) / T : 0;       // 64-row blocks c, c+T, ... owned by this CTA
        for (int i = tid; i < np; i += TRD_THREADS) ysm[i] = 0.0;
        if (n < 2) team_barrier(bar, bar_target, T);                // keeps the slot broadcast race-free

        auto issue_tile = [&](int I, int J, int st) {          // one thread; ONE request per tile (the copy engine
            double* tile = stage_base + (size_t)st * TRD_STAGE_DBL;    // serialises requests, ~150 ns each)
            desc[st] = make_int2(I, J);
            mbar_expect_tx_u32(full_u32 + 8 * st, HH_TS * HH_TS * 8);
            bulk_g2s_u32(smem_u32(tile), A + (((size_t)J * NT + I) << 12), HH_TS * HH_TS * 8, full_u32 + 8 * st);
        };
        TileWalk wi;                                           // meaningful in thread 0 only
        wi.init(0, NT, c, T);

        for (int j0 = 0; j0 < n - 1; j0 += LV.nb) {
            const int pw = min(LV.nb, n - 1 - j0);
            const int JP = j0 >> 6;                                 // the panel's diagonal block
            // first owned slot at or below the panel's diagonal block, number of live owned rows
            const int s0 = (JP > c) ? (JP - c + T - 1) / T : 0;
            const int rows_live = max(0, nown - s0) * 64;
            // threads per row: short dependent chains when the CTA owns few rows
            const int G = (rows_live <= 64) ? 4 : (rows_live <= 128 ? 2 : 1);
            const int sub = tid & (G - 1), rs0 = tid / G, rstep = TRD_THREADS / G;
            // ---- phase A0: the matrix is up to date, acol = A[:, j0]
            {
                double nrm = 0.0;
                for (int rs = tid; rs < rows_live; rs += TRD_THREADS) {
                    const int slot = s0 + (rs >> 6), r = (c + slot * T) * 64 + (rs & 63);
                    const double a = (r >= j0) ? __ldcg(A + hh_tidx(r, j0, NT)) : 0.0;
                    aown[slot * 64 + (rs & 63)] = a;
                    acol[r] = a;
                    if (r >= j0 + 2) nrm = fma(a, a, nrm);
                }
                nrm = cta_sum_d(nrm, red);
                if (tid == 0) part[c * TRD_PART] = nrm;
            }
            for (int jj = 0; jj < pw; ++jj) {
                const int j = j0 + jj;
                TRD_PROF(0);
                team_barrier(bar, bar_target, T);                                   // ---- barrier 1
                TRD_PROF(1);
                // x' of this step (rows J0*64 .. np of acol, ONE request) is requested at once: it travels while the
                // scalars and the panel totals are computed
                const int J0 = (j + 1) >> 6;
                const int mine = [&]() { const int m = NT - J0, total = m * (m + 1) / 2; return (c < total) ? (total - c + T - 1) / T : 0; }();
                if (tid == 0 && mine > 0) {
                    const unsigned xb = (unsigned)(np - J0 * HH_TS) * 8;
                    mbar_expect_tx_u32(xbar_u32, xb);
                    bulk_g2s_u32(smem_u32(xsm + J0 * HH_TS), acol + J0 * HH_TS, xb, xbar_u32);
                }
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
                if (T <= 32) {                      // thread per element, T loads each
                    for (int el = c + T * tid; el < 2 * jj; el += T * TRD_THREADS) {
                        const int idx = (el < jj) ? 2 + el : 2 + HH_NB + (el - jj);
                        double s = 0.0;
                        for (int cc = 0; cc < T; ++cc) s += __ldcg(part + cc * TRD_PART + idx);
                        tot[idx - 2] = s;
                    }
                } else {                            // warp per element, lanes over the team
                    for (int el = c + T * warp; el < 2 * jj; el += T * (TRD_THREADS / 32)) {
                        const int idx = (el < jj) ? 2 + el : 2 + HH_NB + (el - jj);
                        double s = 0.0;
                        for (int cc = lane; cc < T; cc += 32) s += __ldcg(part + cc * TRD_PART + idx);
                        s = warp_sum_d(s);
                        if (lane == 0) tot[idx - 2] = s;
                    }
                }
                __syncthreads();
                TRD_PROF(2);
                const double xfix = sc[3];
                // ---- phase B: y = A x' over my tiles
                {
                    // The tile walk lives in the issuing thread only; everybody else reads (I, J) of a stage from `desc`.
                    int issued = pre_issued;                        // tiles already requested at the end of the last step
                    if (tid == 0) {
                        if (pre_issued == 0) wi.init(J0, NT, c, T); // (else wi continues behind the tiles requested ahead)
                        for (; issued < min(stages, mine); ++issued) {
                            issue_tile(wi.I, wi.J, p_st);
                            p_st = (p_st + 1 == stages) ? 0 : p_st + 1;
                            wi.next();
                        }
                    }
                    pre_issued = 0;
                    if (mine > 0) {
                        mbar_wait_u32(xbar_u32, xphase & 1);
                        ++xphase;
                        // x' = acol with the rows <= j zeroed and x'[j+1] = alpha - beta: only block J0 is touched
                        if (tid < 64) {
                            const int gi = J0 * HH_TS + tid;
                            if (gi <= j) xsm[gi] = 0.0;
                            else if (gi == j + 1) xsm[gi] = xfix;
                        }
                        // the first tile sums an all-zero "previous" direct set into block J0
                        for (int i = tid; i < TRD_DSET_DBL; i += TRD_THREADS) dset[TRD_DSET_DBL + i] = 0.0;
                        fence_proxy_async();                        // this generic write precedes the next step's bulk copy
                        __syncthreads();
                    }
                    // Per tile: warp w owns the tile columns 8w..8w+7, a lane the rows 2 lane, 2 lane + 1 (16-byte loads).
                    //   direct      A_IJ x_J: 8 warp partials per row -> dset[k & 1]; they are summed into y_I one tile LATER,
                    //               by all threads, in the shadow of the next tile's loads (no warp is late at the barrier)
                    //   transposed  A_IJ^T x_I: column sums stay in registers while J does not change (the walk is column-major);
                    //               on a change, 3 shuffle stages leave 4 partials per column in tset, summed one tile later too
                    // x'^T A x' is x'^T y of this CTA's partial y, taken once per column below.
                    const int rrow = 8 * warp + (lane & 7);         // reduce role: lanes 0..7 of a warp own 8 entries of a 64-block (the others shadow them)
                    double tacc[8], xj[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) { tacc[q] = 0.0; xj[q] = 0.0; }
                    int Jacc = -1, Iprev = J0, Jfl = -1;            // running column block; blocks whose partial sets wait
                    unsigned fl_par = 0;
                    auto flush_tacc = [&](double* ts) {             // 8 column sums over 32 lanes -> 4 partials per column
                        const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4;
                        double u[4], v2[2];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const double send = h4 ? tacc[q] : tacc[q + 4];
                            const double keep = h4 ? tacc[q + 4] : tacc[q];
                            u[q] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                        }
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const double send = h3 ? u[q] : u[q + 2];
                            const double keep = h3 ? u[q + 2] : u[q];
                            v2[q] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                        }
                        const double send = h2 ? v2[0] : v2[1];
                        const double keep = h2 ? v2[1] : v2[0];
                        ts[(lane & 3) * TRD_TSET_LD + 8 * warp + (lane >> 2)] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
#pragma unroll
                        for (int q = 0; q < 8; ++q) tacc[q] = 0.0;
                    };
                    // shuffle-free and branch-free on purpose: one basic block with the tile's FMAs, so the loads and the
                    // add tree hide behind them (every lane computes and stores; lanes l, l + 8, l + 16, l + 24 are identical)
                    auto sum_dset = [&](const double* ds, int blk) {        // y_blk += sum of the 8 warp partials
                        const double* dp = ds + rrow;
                        const double p0 = dp[0], p1 = dp[TRD_DSET_LD], p2 = dp[2 * TRD_DSET_LD], p3 = dp[3 * TRD_DSET_LD];
                        const double p4 = dp[4 * TRD_DSET_LD], p5 = dp[5 * TRD_DSET_LD], p6 = dp[6 * TRD_DSET_LD], p7 = dp[7 * TRD_DSET_LD];
                        double* yp = ysm + blk * HH_TS + rrow;
                        const double yo = *yp;
                        *yp = yo + (((p0 + p1) + (p2 + p3)) + ((p4 + p5) + (p6 + p7)));     // lanes l, l + 8, .. store the same value
                    };
                    auto sum_tset = [&](const double* ts, int blk) {        // y_blk += sum of the 4 lane-group partials
                        const double* tp2 = ts + rrow;
                        const double p0 = tp2[0], p1 = tp2[TRD_TSET_LD], p2 = tp2[2 * TRD_TSET_LD], p3 = tp2[3 * TRD_TSET_LD];
                        double* yp = ysm + blk * HH_TS + rrow;
                        *yp = *yp + ((p0 + p1) + (p2 + p3));
                    };
                    if (P.prof != nullptr && tid == 0) { lvstat[0] -= clock64(); lvstat[1] += mine; }
                    for (int k = 0; k < mine; ++k) {
                        const int2 dsc = desc[c_st];
                        mbar_wait_u32(full_u32 + 8 * c_st, (phbits >> c_st) & 1u);
                        phbits ^= 1u << c_st;
                        const double* tile = stage_base + (size_t)c_st * TRD_STAGE_DBL;
                        if (++c_st == stages) c_st = 0;
                        const int I = dsc.x, J = dsc.y;
                        const double2* tp = (const double2*)tile + (8 * warp) * 32 + lane;
                        double2 a[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) a[q] = tp[q * 32];
                        if (J != Jacc) {                            // x_J of my 8 columns: reloaded only when the tile column changes
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const double2 v = ((const double2*)(xsm + J * HH_TS))[4 * warp + q];
                                xj[2 * q] = v.x; xj[2 * q + 1] = v.y;
                            }
                        }
                        double2 xi = ((const double2*)(xsm + I * HH_TS))[lane];
                        if (I == J) xi = make_double2(0.0, 0.0);    // diagonal tile: the transposed product would count it twice
                        // the partial sets of the previous tile (written before its barrier)
                        sum_dset(dset + ((k + 1) & 1) * TRD_DSET_DBL, Iprev);
                        if (Jfl >= 0) sum_tset(tset + (fl_par ^ 1) * TRD_TSET_DBL, Jfl);
                        Jfl = -1;
                        double2 de = make_double2(0.0, 0.0), dodd = make_double2(0.0, 0.0);
#pragma unroll
                        for (int q = 0; q < 8; q += 2) {
                            de.x = fma(a[q].x, xj[q], de.x); de.y = fma(a[q].y, xj[q], de.y);
                            dodd.x = fma(a[q + 1].x, xj[q + 1], dodd.x); dodd.y = fma(a[q + 1].y, xj[q + 1], dodd.y);
                        }
                        *(double2*)(dset + (k & 1) * TRD_DSET_DBL + warp * TRD_DSET_LD + 2 * lane) = make_double2(de.x + dodd.x, de.y + dodd.y);
                        if (Jacc >= 0 && J != Jacc) {               // CTA-uniform
                            flush_tacc(tset + fl_par * TRD_TSET_DBL);
                            fl_par ^= 1;
                            Jfl = Jacc;
                        }
                        Jacc = J;
#pragma unroll
                        for (int q = 0; q < 8; ++q) tacc[q] = fma(a[q].x, xi.x, fma(a[q].y, xi.y, tacc[q]));
                        Iprev = I;
                        __syncthreads();
                        if (tid == 0 && issued < mine) {            // every warp has the tile in registers: refill the stage
                            issue_tile(wi.I, wi.J, p_st);
                            p_st = (p_st + 1 == stages) ? 0 : p_st + 1;
                            wi.next();
                            ++issued;
                        }
                    }
                    if (mine > 0) {                                 // the sets still waiting, then the running column sums
                        flush_tacc(tset + fl_par * TRD_TSET_DBL);   // (the last tile's barrier covers the reads of this parity)
                        __syncthreads();
                        sum_dset(dset + ((mine + 1) & 1) * TRD_DSET_DBL, Iprev);
                        if (Jfl >= 0) sum_tset(tset + (fl_par ^ 1) * TRD_TSET_DBL, Jfl);
                        sum_tset(tset + fl_par * TRD_TSET_DBL, Jacc);
                    }
                    __syncthreads();
                    if (P.prof != nullptr && tid == 0) lvstat[0] += clock64();
                    // the matrix does not change inside a panel: request the first tiles of the next step now, so
                    // that they stream in while the team synchronises
                    if (jj + 1 < pw && tid == 0) {
                        const int J0n = (j + 2) >> 6, mn = NT - J0n, totn = mn * (mn + 1) / 2;
                        const int minen = (c < totn) ? (totn - c + T - 1) / T : 0;
                        wi.init(J0n, NT, c, T);
                        for (; pre_issued < min(stages, minen); ++pre_issued) {
                            issue_tile(wi.I, wi.J, p_st);
                            p_st = (p_st + 1 == stages) ? 0 : p_st + 1;
                            wi.next();
                        }
                    }
                    TRD_PROF(3);
                    double xax = 0.0;
                    for (int r = J0 * HH_TS + tid; r < np; r += TRD_THREADS) {
                        const double yv = ysm[r];
                        ypart[(size_t)c * lnp + r] = yv;
                        if (mine > 0) xax = fma(xsm[r], yv, xax);
                        ysm[r] = 0.0;
                    }
                    xax = cta_sum_d(xax, red);
                    if (tid == 0) part[c * TRD_PART + 1] = xax;
                }
                TRD_PROF(4);
                // row j+1 of the panel (written in earlier steps) is fetched across the barrier
                double wr_pre = 0.0, vr_pre = 0.0;
                if (tid < jj) { wr_pre = __ldcg(Wp + (size_t)tid * ld + j + 1); vr_pre = __ldcg(Vp + (size_t)tid * ld + j + 1); }
                team_barrier(bar, bar_target, T);                                   // ---- barrier 2
                TRD_PROF(5);
                // ---- phase C
                const double beta = sc[0], tj = sc[1], scale = sc[2];
                if (tid < jj) {
                    const double wr = wr_pre, vr = vr_pre;
                    Wrow[tid] = wr; Vrow[tid] = vr;
                    Wtv[tid] = scale * (__ldcg(tot + tid) - beta * wr);
                    Vtv[tid] = scale * (__ldcg(tot + HH_NB + tid) - beta * vr);
                }
                if (warp == 2) {
                    double s = 0.0, y = 0.0;
                    for (int cc = lane; cc < T; cc += 32) {
                        s += __ldcg(part + cc * TRD_PART + 1);
                        y += __ldcg(ypart + (size_t)cc * lnp + j + 1);
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
                TRD_PROF(6);
                double nrm = 0.0;
                for (int base = 0; base < rows_live; base += rstep) {          // CTA-uniform trip count (shuffles inside)
                    const int rs = base + rs0;
                    const bool valid = rs < rows_live;
                    const int slot = s0 + (rs >> 6), ii = rs & 63, r = (c + slot * T) * 64 + ii;
                    const bool act = valid && r > j;
                    // rows of the panel's diagonal block above the reflector
                    if (valid && !act && sub == 0) A[hh_tidx(r, j, NT)] = 0.0;      // (live rows start at the diagonal block)
                    double av = 0.0, pdot = 0.0, adot = 0.0;
                    // next column of the (unchanged inside a panel) matrix: requested with the panel entries, not after them
                    const double anext = (act && sub == 0 && more) ? __ldcg(A + hh_tidx(r, j + 1, NT)) : 0.0;
                    if (act) {
                        // L2-latency bound: 8 (y partials) / 16 (panel entries) independent loads in flight per thread
                        int cc = sub;
                        for (; cc + 7 * G < T; cc += 8 * G) {
                            double t8[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) t8[u] = __ldcg(ypart + (size_t)(cc + u * G) * lnp + r);
#pragma unroll
                            for (int u = 0; u < 8; ++u) av += t8[u];
                        }
                        for (; cc < T; cc += G) av += __ldcg(ypart + (size_t)cc * lnp + r);
                        cc = sub;
                        for (; cc + 7 * G < jj; cc += 8 * G) {
                            double v8[8], w8[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                v8[u] = __ldcg(Vp + (size_t)(cc + u * G) * ld + r);
                                w8[u] = __ldcg(Wp + (size_t)(cc + u * G) * ld + r);
                            }
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                const int c2 = cc + u * G;
                                pdot = fma(v8[u], Wtv[c2], fma(w8[u], Vtv[c2], pdot));
                                adot = fma(v8[u], Wrow[c2], fma(w8[u], Vrow[c2], adot));
                            }
                        }
                        for (; cc < jj; cc += G) {
                            const double vv = __ldcg(Vp + (size_t)cc * ld + r), ww = __ldcg(Wp + (size_t)cc * ld + r);
                            pdot = fma(vv, Wtv[cc], fma(ww, Vtv[cc], pdot));
                            adot = fma(vv, Wrow[cc], fma(ww, Vrow[cc], adot));
                        }
                    }
                    for (int o = 1; o < G; o <<= 1) {              // lanes of a row are adjacent
                        av += __shfl_xor_sync(0xffffffffu, av, o);
                        pdot += __shfl_xor_sync(0xffffffffu, pdot, o);
                        adot += __shfl_xor_sync(0xffffffffu, adot, o);
                    }
                    if (act && sub == 0) {
                        av *= scale;
                        const double vr = (r == j + 1) ? 1.0 : scale * aown[slot * 64 + ii];
                        const double wr = tj * (av - pdot) + alpha2 * vr;
                        Vp[(size_t)jj * ld + r] = vr;
                        Wp[(size_t)jj * ld + r] = wr;
                        A[hh_tidx(r, j, NT)] = vr;
                        if (more) {
                            const double an = anext - adot - vr * wj1 - wr;
                            aown[slot * 64 + ii] = an;
                            acol[r] = an;
                            if (r >= j + 3) nrm = fma(an, an, nrm);
                        }
                    }
                }
                TRD_PROF(7);
                if (more) {
                    nrm = cta_sum_d(nrm, red);          // (its __syncthreads also publish aown and the new panel column)
                    if (tid == 0) part[c * TRD_PART] = nrm;
                    // partial W^T acol, V^T acol over my live rows: a warp takes 4 columns at a time so that
                    // their (L2-latency bound) loads overlap
                    const int ncols = 2 * (jj + 1);
                    for (int q0 = 4 * warp; q0 < ncols; q0 += 4 * (TRD_THREADS / 32)) {
                        const double* Pn[4];
                        double acc[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int qd = min(q0 + u, ncols - 1);
                            const int which = qd > jj, cc = which ? qd - (jj + 1) : qd;
                            Pn[u] = (which ? Vp : Wp) + (size_t)cc * ld;
                            acc[u] = 0.0;
                        }
                        // 4 row groups x 4 columns = 16 independent (predicated) loads in flight per lane; panel entries of
                        // rows above the reflector are never written, so they are not read either
                        for (int rs = lane; rs < rows_live; rs += 128) {
                            double pv[4][4], avv[4];
                            int rr[4];
                            bool ok[4];
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                const int rg = min(rs + 32 * g, rows_live - 1);
                                const int slot = s0 + (rg >> 6);
                                rr[g] = (c + slot * T) * 64 + (rg & 63);
                                ok[g] = (rs + 32 * g < rows_live) && (rr[g] >= j + 2);
                                avv[g] = aown[slot * 64 + (rg & 63)];
                            }
#pragma unroll
                            for (int g = 0; g < 4; ++g)
#pragma unroll
                                for (int u = 0; u < 4; ++u) pv[g][u] = ok[g] ? __ldcg(Pn[u] + rr[g]) : 0.0;
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                const double a1 = ok[g] ? avv[g] : 0.0;
#pragma unroll
                                for (int u = 0; u < 4; ++u) acc[u] = fma(pv[g][u], a1, acc[u]);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const double v = warp_sum_d(acc[u]);
                            const int qd = q0 + u;
                            if (lane == 0 && qd < ncols) {
                                const int which = qd > jj, cc = which ? qd - (jj + 1) : qd;
                                part[c * TRD_PART + 2 + which * HH_NB + cc] = v;
                            }
                        }
                    }
                }
            }
            // ---- trailing update  A22 -= V W^T + W V^T  on the tiles at or below (jn, jn)
            TRD_PROF(8);
            team_barrier(bar, bar_target, T);
            TRD_PROF(9);
            {
                const int jn = j0 + pw, kpad = (pw + 3) & ~3, nq = (kpad + TRD_SYR_KH - 1) / TRD_SYR_KH;
                // operand chunk q of tile (I, J): rows of V_I, W_I, V_J, W_J x 16 panel columns -> S[buf][which][k][68]
                auto stage_unit = [&](int I, int J, int q, int buf) {
                    double* dst = stage_base + buf * (4 * TRD_SYR_LD * TRD_SYR_KH);
                    const int kh = q * TRD_SYR_KH;
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        const int e = tid + t * TRD_THREADS;            // 4 blocks x 16 k x 32 16-byte chunks
                        const int which = e >> 9, k = (e >> 5) & 15, r2 = (e & 31) * 2;
                        const double* src = ((which & 1) ? Wp : Vp) + (size_t)(kh + k) * ld + ((which >> 1) ? J : I) * HH_TS + r2;
                        cp_async16_zfill(dst + which * (TRD_SYR_LD * TRD_SYR_KH) + k * TRD_SYR_LD + r2, src, kh + k < pw);
                    }
                };
                TileWalk w, wn;
                w.init(jn >> 6, NT, c, T);
                wn = w;
                int buf = 0;
                if (w.valid()) stage_unit(w.I, w.J, 0, 0);
                cp_async_commit();
                const int fk = lane & 3, fr = lane >> 2;
                for (; w.valid(); w.next()) {
                    const int I = w.I, J = w.J;
                    wn.next();
                    const int lc = 8 * warp + fr, gc = J * HH_TS + lc;
                    double* ctile = A + (((size_t)J * NT + I) << 12) + lc * HH_TS;
                    double2 cv[8];                                  // the C values of this lane, requested early
                    if (gc >= jn) {
#pragma unroll
                        for (int rb = 0; rb < 8; ++rb) cv[rb] = __ldcg((const double2*)(ctile + 8 * rb + 2 * fk));
                    }
                    double acc[8][2];
#pragma unroll
                    for (int rb = 0; rb < 8; ++rb) { acc[rb][0] = 0.0; acc[rb][1] = 0.0; }
                    for (int q = 0; q < nq; ++q) {
                        // request the next chunk (of this tile or of my next tile) into the other buffer
                        if (q + 1 < nq) stage_unit(I, J, q + 1, buf ^ 1);
                        else if (wn.valid()) stage_unit(wn.I, wn.J, 0, buf ^ 1);
                        cp_async_commit();
                        cp_async_wait<1>();
                        __syncthreads();
                        const double* S = stage_base + buf * (4 * TRD_SYR_LD * TRD_SYR_KH);
                        const double* VIs = S, *WIs = S + TRD_SYR_LD * TRD_SYR_KH, *VJs = S + 2 * TRD_SYR_LD * TRD_SYR_KH,
                                     *WJs = S + 3 * TRD_SYR_LD * TRD_SYR_KH;
                        const int kc = min(TRD_SYR_KH, kpad - q * TRD_SYR_KH);
                        for (int k0 = 0; k0 < kc; k0 += 4) {
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
                        __syncthreads();
                        buf ^= 1;
                    }
                    if (gc >= jn) {
#pragma unroll
                        for (int rb = 0; rb < 8; ++rb) {
                            const int lr = 8 * rb + 2 * fk, gr = I * HH_TS + lr;
                            double2 v = cv[rb];
                            if (gr >= jn) v.x -= acc[rb][0];
                            if (gr + 1 >= jn) v.y -= acc[rb][1];
                            *(double2*)(ctile + lr) = v;
                        }
                    }
                }
                cp_async_wait<0>();
                fence_proxy_async();                   // S staging (generic proxy) precedes the next bulk copies
            }
            TRD_PROF(10);
            team_barrier(bar, bar_target, T);
            TRD_PROF(11);
        }
        if (c == 0 && tid == 0) dvec[n - 1] = __ldcg(A + hh_tidx(n - 1, n - 1, NT));
        // the next job's first barrier separates this job's scratch use from the next one's
    }
    if (P.prof != nullptr && tid == 0) {                   // per-CTA, per-level stream statistics (trace only)
        long long* o = P.prof + TRD_PROF_LVSTAT + ((size_t)blockIdx.x * TRD_MAX_LEVELS + lvi) * 3;
        o[0] = lvstat[0]; o[1] = lvstat[1]; o[2] = clock64() - lvstat[2];
    }
    }   // levels
    if (prof_on)
        for (int i = 0; i < 12; ++i) P.prof[i] = prof_acc[i];
    if (P.prof != nullptr && threadIdx.x == 0) {           // finish time of every CTA (trace only)
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        P.prof[16 + blockIdx.x] = (long long)t;
    }
}

// ---------------------------------------------------------------------------------------------------
// Tridiagonalisation of the users whose whole matrix fits in shared memory (n <= TRD_SMALL_MAX): one CTA per user,
// unblocked Householder (LAPACK dsytd2 scheme) on a full symmetric copy in shared memory -- no L2 round trips and no
// team barriers between the columns, which is what the persistent kernel above spends most of its time on for
// matrices of two or three tiles.  Same outputs: d, e, tau and the reflectors in the columns of A (tile-major, 1 at
// row j+1, zeros above inside the diagonal tile), same sign conventions, fixed reduction order.
// ---------------------------------------------------------------------------------------------------
#define TRD_SMALL_MAX 160
static inline size_t trd_small_smem_bytes(int nmax) { return ((size_t)nmax * (nmax | 1) + 3 * (size_t)nmax + 64) * sizeof(double); }

__global__ void __launch_bounds__(TRD_THREADS, 1) trd_small_kernel(const HJob* __restrict__ jobs, int job0, double* __restrict__ Aall,
                                                                   double* __restrict__ dall, double* __restrict__ eall,
                                                                   double* __restrict__ tauall) {
    extern __shared__ __align__(16) double tsm[];
    const HJob jb = jobs[job0 + blockIdx.x];
    const int n = jb.n, NT = jb.np >> 6, ld = n | 1, tid = threadIdx.x;
    double* As = tsm;                       // [n][ld] column-major, full symmetric
    double* v = As + (size_t)n * ld;        // [n]
    double* p = v + n;
    double* w = p + n;
    double* red = w + n;                    // [64]
    double* A = Aall + jb.m_off;
    double* dvec = dall + jb.r_off;
    double* evec = eall + jb.r_off;
    double* tvec = tauall + jb.r_off;
    // the tiles on / below the diagonal hold sym(lower(L)); mirror into the upper triangle
    for (int e = tid; e < n * n; e += TRD_THREADS) {
        const int c = e / n, r = e - c * n;
        if (r >= c) {
            const double a = __ldcg(A + hh_tidx(r, c, NT));
            As[r + (size_t)c * ld] = a;
            As[c + (size_t)r * ld] = a;
        }
    }
    __syncthreads();
    for (int j = 0; j < n - 1; ++j) {
        // ---- reflector of column j: alpha = A[j+1][j], s = sum_{r >= j+2} A[r][j]^2
        double s = 0.0;
        for (int r = j + 2 + tid; r < n; r += TRD_THREADS) { const double a = As[r + (size_t)j * ld]; s = fma(a, a, s); }
        s = cta_sum_d(s, red);
        const double alpha = As[j + 1 + (size_t)j * ld];
        double beta, tj, scale;
        if (s == 0.0) { beta = alpha; tj = 0.0; scale = 0.0; }
        else {
            beta = -copysign(sqrt(fma(alpha, alpha, s)), alpha);
            tj = (beta - alpha) / beta;
            scale = 1.0 / (alpha - beta);
        }
        if (tid == 0) { dvec[j] = As[j + (size_t)j * ld]; evec[j] = beta; tvec[j] = tj; }
        for (int r = tid; r < n; r += TRD_THREADS) {
            const double vr = (r <= j) ? 0.0 : (r == j + 1 ? 1.0 : scale * As[r + (size_t)j * ld]);
            v[r] = vr;
            if (r >= (j & ~63)) A[hh_tidx(r, j, NT)] = vr;        // reflector column (zeros above, inside the diagonal tile)
        }
        __syncthreads();
        // ---- p = A v over the trailing block (rows / columns > j); thread per row, four partial sums
        for (int i = j + 1 + tid; i < n; i += TRD_THREADS) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int k = j + 1;
            for (; k + 3 < n; k += 4) {
                a0 = fma(As[i + (size_t)k * ld], v[k], a0);
                a1 = fma(As[i + (size_t)(k + 1) * ld], v[k + 1], a1);
                a2 = fma(As[i + (size_t)(k + 2) * ld], v[k + 2], a2);
                a3 = fma(As[i + (size_t)(k + 3) * ld], v[k + 3], a3);
            }
            for (; k < n; ++k) a0 = fma(As[i + (size_t)k * ld], v[k], a0);
            p[i] = (a0 + a1) + (a2 + a3);
        }
        __syncthreads();
        double pv = 0.0;
        for (int i = j + 1 + tid; i < n; i += TRD_THREADS) pv = fma(p[i], v[i], pv);
        pv = cta_sum_d(pv, red);
        // w = tau p - (tau^2 / 2)(p^T v) v
        const double alpha2 = -0.5 * tj * tj * pv;
        for (int i = j + 1 + tid; i < n; i += TRD_THREADS) w[i] = fma(tj, p[i], alpha2 * v[i]);
        __syncthreads();
        // ---- A22 -= v w^T + w v^T (both triangles)
        for (int i = j + 1 + tid; i < n; i += TRD_THREADS) {
            const double vi = v[i], wi = w[i];
            for (int k = j + 1; k < n; ++k) {
                double* a = As + i + (size_t)k * ld;
                *a = *a - (vi * w[k] + wi * v[k]);
            }
        }
        __syncthreads();
    }
    if (tid == 0) dvec[n - 1] = As[n - 1 + (size_t)(n - 1) * ld];
}
