// Host side of the large path (n > GSI_S_MAX_N): Householder tridiagonalisation -> divide & conquer
// -> back-transformation.  Included by gsi.cu after its helpers (DevBuf, Workspace, MetaBuilder,
// upload_meta, finish_chunk, Job, RunOut).
#pragma once

struct HhPlan {
    int nj = 0;
    std::vector<HJob> jobs;
    std::vector<DcLeaf> leaves;
    std::vector<DcNode> nodes;            // ordered by level, then job
    std::vector<int> lvl_begin;           // [Lmax + 2]; level l nodes are [lvl_begin[l], lvl_begin[l+1])
    std::vector<int> lvl_mmax;            // [Lmax + 1]
    std::vector<int2> bt_items;
    int64_t mtot = 0, rtot = 0, vtot = 0, ltot = 0, max_nk = 0;
    int Lmax = 0, npmax = 0, nmax = 0;
};


static void hh_cuts(int n, int& L, std::vector<int>& cuts) {
    L = 0;
    while (((n + (1 << L) - 1) >> L) > 24) ++L;
    const int nl = 1 << L;
    cuts.assign(nl + 1, 0);
    for (int i = 1; i < nl; ++i) cuts[i] = (int)((((int64_t)n * i / nl) + 4) / 8 * 8);
    cuts[nl] = n;
}

static void hh_build_plan(const Job* jobs, int nj, HhPlan& pl) {
    pl = HhPlan();
    pl.nj = nj;
    pl.jobs.resize(nj);
    std::vector<std::vector<int>> cuts(nj);
    for (int j = 0; j < nj; ++j) {
        HJob& h = pl.jobs[j];
        h.n = jobs[j].n; h.np = hh_np(h.n); h.pad_ = 0;
        h.m_off = pl.mtot; h.r_off = pl.rtot; h.item_off = jobs[j].item_off; h.vec_off = pl.vtot; h.lam_off = pl.ltot;
        hh_cuts(h.n, h.levels, cuts[j]);
        pl.mtot += (int64_t)h.np * h.np; pl.rtot += h.np;
        pl.vtot += (int64_t)h.n * std::max(h.n, 2); pl.ltot += std::max(h.n, 2);
        pl.max_nk = std::max<int64_t>(pl.max_nk, (int64_t)h.n * std::max(h.n, 2));
        pl.Lmax = std::max(pl.Lmax, h.levels); pl.npmax = std::max(pl.npmax, h.np); pl.nmax = std::max(pl.nmax, h.n);
        const int nl = 1 << h.levels;
        for (int i = 0; i < nl; ++i) pl.leaves.push_back({j, cuts[j][i], cuts[j][i + 1] - cuts[j][i]});
    }
    pl.lvl_begin.assign(pl.Lmax + 2, 0);
    pl.lvl_mmax.assign(pl.Lmax + 1, 0);
    for (int l = 1; l <= pl.Lmax; ++l) {
        pl.lvl_begin[l] = (int)pl.nodes.size();
        for (int j = 0; j < nj; ++j) {
            const int L = pl.jobs[j].levels;
            if (L < l) continue;
            const int step = 1 << l, half = step >> 1, cnt = 1 << (L - l);
            for (int i = 0; i < cnt; ++i) {
                const int off = cuts[j][i * step], mid = cuts[j][i * step + half], end = cuts[j][(i + 1) * step];
                pl.nodes.push_back({j, off, mid - off, end - mid, l == L ? 1 : 0, 0});
                pl.lvl_mmax[l] = std::max(pl.lvl_mmax[l], end - off);
            }
        }
    }
    pl.lvl_begin[pl.Lmax + 1] = (int)pl.nodes.size();
    // back-transform work items, largest users first (jobs are sorted by n descending)
    for (int j = 0; j < nj; ++j)
        for (int cb = 0; cb * BT_CB < pl.jobs[j].n; ++cb) pl.bt_items.push_back(make_int2(j, cb));
}

struct HhDev {           // device views of one chunk
    const HJob* jobs; const DcLeaf* leaves; const DcNode* nodes; const int2* bt_items;
    DcState* state; int32_t* tile_off; unsigned int* sigmax; int32_t* kuser; int* ctl; uint8_t* bt_skip;
    double *A, *Qa, *Qb, *S;
    double *d, *e, *tau, *lamA, *lamB, *dk, *zk, *zhat, *dctau, *lamk, *dval, *ds, *zs, *sgn, *deg, *scale, *rot;
    int32_t *orig, *gmap, *colsrc, *dsrc, *pos_nd, *pos_df, *src;
    int64_t* user; int64_t* vec_dst; int64_t* lam_dst; int32_t* n_arr; int32_t* ld_arr; int64_t* moff_arr; int64_t* ioff_arr;
    int64_t* roff_arr; int64_t* voff_arr; int64_t* loff_arr;
};


#define HH_CTL_INTS 12288      // [0..8) trd level queues, [8] bt queue, [64 + 256 l ..) slots, [2048 + 256 l ..) barriers, [3840..) trace

static int hh_alloc(gsi_ctx* ctx, const HhPlan& pl, const Job* jobs, HhDev& D) {
    Workspace& ws = WS(ctx);
    int rc;
    const int nj = pl.nj;
    if ((rc = ws.hhA.ensure(ctx, pl.mtot * 8)) != GSI_OK) return rc;
    if ((rc = ws.hhQa.ensure(ctx, pl.mtot * 8)) != GSI_OK) return rc;
    if ((rc = ws.hhQb.ensure(ctx, pl.mtot * 8)) != GSI_OK) return rc;
    if ((rc = ws.hhS.ensure(ctx, pl.mtot * 8)) != GSI_OK) return rc;
    if ((rc = ws.hhvec.ensure(ctx, (size_t)pl.rtot * 8 * 21)) != GSI_OK) return rc;
    if ((rc = ws.hhivec.ensure(ctx, (size_t)pl.rtot * 4 * 7)) != GSI_OK) return rc;
    if ((rc = ws.vec_pad.ensure(ctx, pl.vtot * 8)) != GSI_OK) return rc;
    if ((rc = ws.lam_pad.ensure(ctx, pl.ltot * 8)) != GSI_OK) return rc;
    std::vector<int64_t> user(nj), moff(nj), ioff(nj), roff(nj), voff(nj), loff(nj);
    std::vector<int32_t> n(nj), ld(nj);
    for (int j = 0; j < nj; ++j) {
        user[j] = jobs[j].user; moff[j] = pl.jobs[j].m_off; ioff[j] = pl.jobs[j].item_off; roff[j] = pl.jobs[j].r_off;
        voff[j] = pl.jobs[j].vec_off; loff[j] = pl.jobs[j].lam_off; n[j] = pl.jobs[j].n; ld[j] = pl.jobs[j].np;
    }
    MetaBuilder mb;
    const size_t o_jobs = mb.add(pl.jobs), o_leaves = mb.add(pl.leaves), o_nodes = mb.add(pl.nodes), o_items = mb.add(pl.bt_items);
    const size_t o_user = mb.add(user), o_moff = mb.add(moff), o_ioff = mb.add(ioff), o_roff = mb.add(roff), o_voff = mb.add(voff),
                 o_loff = mb.add(loff), o_n = mb.add(n), o_ld = mb.add(ld);
    const size_t o_state = mb.reserve(std::max<size_t>(1, pl.nodes.size()) * sizeof(DcState));
    const size_t o_tile = mb.reserve((2 * pl.nodes.size() + 2) * 4);
    const size_t o_sig = mb.reserve((size_t)nj * 4), o_ku = mb.reserve((size_t)nj * 4), o_ctl = mb.reserve(HH_CTL_INTS * 4);
    const size_t o_vd = mb.reserve((size_t)nj * 8), o_ldst = mb.reserve((size_t)nj * 8);
    const size_t o_skip = mb.reserve((size_t)nj);
    char* base;
    if ((rc = upload_meta(ctx, mb, &base)) != GSI_OK) return rc;
    D.jobs = (const HJob*)(base + o_jobs); D.leaves = (const DcLeaf*)(base + o_leaves); D.nodes = (const DcNode*)(base + o_nodes);
    D.bt_items = (const int2*)(base + o_items); D.state = (DcState*)(base + o_state); D.tile_off = (int32_t*)(base + o_tile);
    D.sigmax = (unsigned int*)(base + o_sig); D.kuser = (int32_t*)(base + o_ku); D.ctl = (int*)(base + o_ctl); D.bt_skip = (uint8_t*)(base + o_skip);
    D.user = (int64_t*)(base + o_user); D.vec_dst = (int64_t*)(base + o_vd); D.lam_dst = (int64_t*)(base + o_ldst);
    D.n_arr = (int32_t*)(base + o_n); D.ld_arr = (int32_t*)(base + o_ld); D.moff_arr = (int64_t*)(base + o_moff);
    D.ioff_arr = (int64_t*)(base + o_ioff); D.roff_arr = (int64_t*)(base + o_roff); D.voff_arr = (int64_t*)(base + o_voff);
    D.loff_arr = (int64_t*)(base + o_loff);
    D.A = ws.hhA.as<double>(); D.Qa = ws.hhQa.as<double>(); D.Qb = ws.hhQb.as<double>(); D.S = ws.hhS.as<double>();
    double* v = ws.hhvec.as<double>();
    const int64_t r = pl.rtot;
    D.d = v; D.e = v + r; D.tau = v + 2 * r; D.lamA = v + 3 * r; D.lamB = v + 4 * r; D.dk = v + 5 * r; D.zk = v + 6 * r;
    D.zhat = v + 7 * r; D.dctau = v + 8 * r; D.lamk = v + 9 * r; D.dval = v + 10 * r; D.ds = v + 11 * r; D.zs = v + 12 * r;
    D.sgn = v + 13 * r; D.deg = v + 14 * r; D.scale = v + 15 * r; D.rot = v + 16 * r;      // rot: 4 r
    int32_t* iv = ws.hhivec.as<int32_t>();
    D.orig = iv; D.gmap = iv + r; D.colsrc = iv + 2 * r; D.dsrc = iv + 3 * r; D.pos_nd = iv + 4 * r; D.pos_df = iv + 5 * r; D.src = iv + 6 * r;
    cudaStream_t st = ctx->stream;
    GSI_CUDA(ctx, cudaMemsetAsync(D.A, 0, pl.mtot * 8, st));
    GSI_CUDA(ctx, cudaMemsetAsync(D.Qa, 0, pl.mtot * 8, st));
    GSI_CUDA(ctx, cudaMemsetAsync(D.Qb, 0, pl.mtot * 8, st));
    return GSI_OK;
}

// GSI_TRACE: synchronising stopwatch around a group of launches (diagnostics only)
struct HhTrace {
    gsi_ctx* ctx; const char* name; cudaEvent_t a = nullptr, b = nullptr;
    HhTrace(gsi_ctx* c, const char* n) : ctx(c), name(n) {
        if (ctx->trace) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, ctx->stream); }
    }
    ~HhTrace() {
        if (!a) return;
        cudaEventRecord(b, ctx->stream); cudaEventSynchronize(b);
        float ms = 0.f; cudaEventElapsedTime(&ms, a, b);
        fprintf(stderr, "[gsi trace] %s %.3f ms\n", name, ms);
        cudaEventDestroy(a); cudaEventDestroy(b);
    }
};

#include "sbr_host.cuh"

struct HhSbrState { SbrDev dev; SbrStage2 st2; int nu = 0; };     // two-stage users of the chunk in flight (a prefix of its jobs)

// size class of a user on the tridiagonalisation levels.  Every CTA of a team pays the per-column
// synchronisation latency, so teams are kept as small as the tail allows: the biggest users start first
// and everything smaller fills the other SMs behind them.
static int hh_level_of(int np) { return np > 4096 ? 0 : (np > 2048 ? 1 : (np > 1024 ? 2 : 3)); }

// tridiagonalisation only (uses D.jobs, D.A, D.d, D.e, D.tau, D.ctl): A = Q T Q^T, T in (d, e), reflectors in A / tau
static int hh_trd(gsi_ctx* ctx, const HhPlan& pl, const HhDev& D, int forced_team, HhSbrState& SB) {
    Workspace& ws = WS(ctx);
    cudaStream_t st = ctx->stream;
    const int nj = pl.nj;
    int rc;
    // ---------------- the big users: two-stage reduction (kern_sbr.cuh); jobs are sorted by n descending
    SB.nu = 0;
    if (forced_team <= 0 && sbr_min_n() > 0) while (SB.nu < nj && pl.jobs[SB.nu].n >= sbr_min_n()) ++SB.nu;
    if (SB.nu > 0) {
        if ((rc = sbr_alloc(ctx, pl, SB.nu, SB.dev)) != GSI_OK) return rc;
        { HhTrace tr(ctx, "sbr stage 1 (dense -> band)"); if ((rc = sbr_stage1(ctx, pl, D, SB.dev)) != GSI_OK) return rc; }
        if ((rc = sbr_stage2(ctx, pl, D, SB.dev, SB.st2)) != GSI_OK) return rc;
    }
    const int first_one_stage = SB.nu;
    // ---------------- tridiagonalisation: ONE persistent launch, levels of nested teams (kern_trd.cuh)
    {
        const int sms = ctx->sm_count;
        int level_T[4] = {37, 6, 2, 1};
        int level_nb[4] = {64, 64, 32, 16};
        if (const char* lt = getenv("GSI_TRD_TEAMS")) sscanf(lt, "%d,%d,%d,%d", &level_T[0], &level_T[1], &level_T[2], &level_T[3]);
        if (const char* lt = getenv("GSI_TRD_NB")) sscanf(lt, "%d,%d,%d,%d", &level_nb[0], &level_nb[1], &level_nb[2], &level_nb[3]);
        TrdParams P;
        memset(&P, 0, sizeof P);
        P.jobs = D.jobs; P.A = D.A; P.d = D.d; P.e = D.e; P.tau = D.tau;
        P.nlevels = 0;
        size_t smem = 0;
        int b = first_one_stage;
        // users whose matrix fits in shared memory go to the CTA-resident kernel (jobs are sorted by n descending: the tail)
        static const int small_cap = getenv("GSI_TRD_SMALL_MAX") ? std::min(TRD_SMALL_MAX, atoi(getenv("GSI_TRD_SMALL_MAX"))) : TRD_SMALL_MAX;
        int nj_big = nj;
        if (forced_team <= 0) while (nj_big > first_one_stage && pl.jobs[nj_big - 1].n <= small_cap) --nj_big;
        // (The route of a user depends on its own n only -- never on what else is in the chunk -- so that a user's record
        // is bit-identical however the users are sharded over chunks, ranks or GPUs: SURVEY.md section 7 test (h).)
        if (nj_big < nj) {
            const size_t ssm = trd_small_smem_bytes(pl.jobs[nj_big].n);
            GSI_CUDA(ctx, cudaFuncSetAttribute(trd_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm));
            GsiSpan sp(ctx, GSI_T_TRD, 1);
            HhTrace tr(ctx, "trd_small");
            trd_small_kernel<<<nj - nj_big, TRD_THREADS, ssm, st>>>(D.jobs, nj_big, D.A, D.d, D.e, D.tau);
            sp.end();
            GSI_CUDA(ctx, cudaGetLastError());
        }
        while (b < nj_big) {                                       // jobs are sorted by n descending
            const int lv = hh_level_of(pl.jobs[b].np);
            int e = b;
            while (e < nj_big && hh_level_of(pl.jobs[e].np) == lv) ++e;
            TrdLevel& L = P.lv[P.nlevels++];
            L.T = forced_team > 0 ? std::min(sms, forced_team) : level_T[lv];
            L.job0 = b; L.njobs = e - b; L.npmax = pl.jobs[b].np;
            L.nb = std::min(HH_NB, std::max(4, level_nb[lv] & ~3));
            L.own = (L.npmax / 64 + L.T - 1) / L.T;
            L.stages = TRD_MAX_STAGES;
            if (const char* ms = getenv("GSI_TRD_STAGES")) L.stages = std::max(2, std::min(TRD_MAX_STAGES, atoi(ms)));   // probe
            while (L.stages > 2 && trd_smem_bytes(L.npmax, L.stages, L.own) > 227 * 1024) --L.stages;   // (a refilled stage is consumed >= 2 tiles later)
            if (trd_smem_bytes(L.npmax, L.stages, L.own) > 227 * 1024)
                return gsi_fail(ctx, GSI_ERR_INVALID, "user with n = %d does not fit the tridiagonalisation kernel", pl.jobs[b].n);
            smem = std::max(smem, trd_smem_bytes(L.npmax, L.stages, L.own));
            b = e;
        }
        // teams per level (the kernel derives the same nesting), scratch
        size_t o_acol = 0, o_ypart = 0, o_part = 0, o_pan = 0;
        int size_prev = sms, teams = 1;
        bool nested = true;
        std::vector<size_t> off_acol(P.nlevels), off_ypart(P.nlevels), off_part(P.nlevels), off_pan(P.nlevels);
        std::vector<int> nteams(P.nlevels);
        for (int l = 0; l < P.nlevels; ++l) {
            TrdLevel& L = P.lv[l];
            if (L.T == 1) teams = sms;
            else if (nested) { const int nsub = size_prev / L.T; teams *= nsub; size_prev = L.T; }
            nteams[l] = teams;
            if (teams > 240) return gsi_fail(ctx, GSI_ERR_INVALID, "internal: too many teams");
            off_acol[l] = o_acol; off_ypart[l] = o_ypart; off_part[l] = o_part; off_pan[l] = o_pan;
            o_acol += (size_t)teams * L.npmax;
            o_ypart += (size_t)teams * L.T * L.npmax;
            o_part += (size_t)teams * L.T * TRD_PART + (size_t)teams * 2 * HH_NB;
            o_pan += (size_t)2 * teams * L.npmax * HH_NB;
        }
        if ((rc = ws.trd_acol.ensure(ctx, o_acol * 8)) != GSI_OK) return rc;
        if ((rc = ws.trd_ypart.ensure(ctx, o_ypart * 8)) != GSI_OK) return rc;
        if ((rc = ws.trd_part.ensure(ctx, o_part * 8)) != GSI_OK) return rc;
        if ((rc = ws.trd_panels.ensure(ctx, o_pan * 8)) != GSI_OK) return rc;
        for (int l = 0; l < P.nlevels; ++l) {
            TrdLevel& L = P.lv[l];
            L.acol = ws.trd_acol.as<double>() + off_acol[l];
            L.ypart = ws.trd_ypart.as<double>() + off_ypart[l];
            L.part = ws.trd_part.as<double>() + off_part[l];
            L.tot = L.part + (size_t)nteams[l] * L.T * TRD_PART;
            L.Vp = ws.trd_panels.as<double>() + off_pan[l];
            L.Wp = L.Vp + (size_t)nteams[l] * L.npmax * HH_NB;
            L.queue = D.ctl + l; L.slot = D.ctl + 64 + 256 * l; L.bar = (unsigned*)(D.ctl + 2048 + 256 * l);
        }
        P.prof = ctx->trace ? (long long*)(D.ctl + 3840) : nullptr;
        GSI_CUDA(ctx, cudaMemsetAsync(D.ctl, 0, HH_CTL_INTS * 4, st));
        if (P.nlevels > 0) GSI_CUDA(ctx, cudaFuncSetAttribute(trd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GsiSpan sp(ctx, GSI_T_TRD, 1);
        void* args[] = {&P};
        cudaEvent_t ta = nullptr, tb = nullptr;
        if (ctx->trace) { cudaEventCreate(&ta); cudaEventCreate(&tb); cudaEventRecord(ta, st); }
        if (P.nlevels > 0) GSI_CUDA(ctx, cudaLaunchCooperativeKernel((void*)trd_kernel, dim3(sms), dim3(TRD_THREADS), args, smem, st));
        sp.end();
        if (ctx->trace) {
            cudaEventRecord(tb, st); cudaEventSynchronize(tb);
            float ms = 0.f; cudaEventElapsedTime(&ms, ta, tb);
            double n3 = 0; for (int j = 0; j < nj; ++j) n3 += (double)pl.jobs[j].n * pl.jobs[j].n * pl.jobs[j].n;
            fprintf(stderr, "[gsi trace] trd: %d users n=%d..%d  %.2f ms  (sum n^3 = %.3g, symv bytes %.3g -> %.1f GB/s); levels:",
                    nj, pl.jobs[0].n, pl.jobs[nj - 1].n, ms, n3, n3 * 8 / 6, n3 * 8 / 6 / (ms * 1e-3) / 1e9);
            for (int l = 0; l < P.nlevels; ++l)
                fprintf(stderr, " [T=%d x%d teams, %d users, np<=%d, %d stages]", P.lv[l].T, nteams[l], P.lv[l].njobs, P.lv[l].npmax, P.lv[l].stages);
            long long pr[12];
            cudaMemcpy(pr, D.ctl + 3840, sizeof pr, cudaMemcpyDeviceToHost);
            static const char* nm[12] = {"dots", "bar1", "scal", "symv", "ywr", "bar2", "preC", "rows", "dots+A0", "bar3", "syr2k", "bar4"};
            long long tot = 0; for (int i = 0; i < 12; ++i) tot += pr[i];
            fprintf(stderr, "\n[gsi trace]   CTA0 cycles %%:");
            for (int i = 0; i < 12; ++i) fprintf(stderr, " %s %.1f", nm[i], 100.0 * pr[i] / std::max<long long>(tot, 1));
            fprintf(stderr, "  (total %.1f ms @1.965GHz)\n", tot / 1.965e6);
            std::vector<long long> fin(16 + sms);
            cudaMemcpy(fin.data(), D.ctl + 3840, fin.size() * 8, cudaMemcpyDeviceToHost);
            std::vector<double> ft;
            for (int i = 0; i < sms; ++i) ft.push_back((fin[16 + i] - fin[15]) * 1e-6);
            std::sort(ft.begin(), ft.end());
            fprintf(stderr, "[gsi trace]   CTA finish times (ms): min %.1f  p25 %.1f  median %.1f  p75 %.1f  p90 %.1f  max %.1f\n",
                    ft[0], ft[sms / 4], ft[sms / 2], ft[3 * sms / 4], ft[(9 * sms) / 10], ft[sms - 1]);
            std::vector<long long> lst((size_t)sms * TRD_MAX_LEVELS * 3);
            cudaMemcpy(lst.data(), (long long*)(D.ctl + 3840) + TRD_PROF_LVSTAT, lst.size() * 8, cudaMemcpyDeviceToHost);
            for (int l = 0; l < P.nlevels; ++l) {          // what a CTA streams while it is in the symv phase, per level
                double cyc = 0, tiles = 0, tot = 0, tmax = 0; int nc = 0;
                for (int b2 = 0; b2 < sms; ++b2) {
                    const long long* o = &lst[((size_t)b2 * TRD_MAX_LEVELS + l) * 3];
                    if (o[2] == 0) continue;
                    cyc += (double)o[0]; tiles += (double)o[1]; tot += (double)o[2]; tmax = std::max(tmax, (double)o[2]); ++nc;
                }
                fprintf(stderr, "[gsi trace]   level %d (T=%d, %d stages): %d CTAs, %.0f tiles, symv %.1f %% of the level's CTA time, %.1f GB/s per CTA in symv, "
                        "level time mean %.1f ms max %.1f ms\n", l, P.lv[l].T, P.lv[l].stages, nc, tiles, 100.0 * cyc / std::max(tot, 1.0),
                        tiles * 32768.0 / std::max(cyc / 1.965e9, 1e-12) / 1e9, tot / std::max(nc, 1) / 1.965e6, tmax / 1.965e6);
            }
            cudaEventDestroy(ta); cudaEventDestroy(tb);
        }
    }
    return GSI_OK;
}

// ---- tensor-core merges: host plan ---------------------------------------------------------------------------------------
struct HhTcBatch { int level, first, ntasks, Mmax, Nmax, Kmax; };
struct HhTcPlan {
    std::vector<HhTcBatch> batches;
    const uint8_t* d_flags = nullptr; const DcTcTask* d_tasks = nullptr; TcTask* d_tc = nullptr;
    int8_t* d_planes = nullptr; int32_t* d_expo = nullptr;
    int S = 8;
};
// merged size from which a node's GEMMs go to the tcgen05 engine (GSI_TC_MIN; 0 = never).  Below ~2,000 the slicing passes and
// the tile quantisation (128 x 64 tiles, one CTA per SM) cost more than the engine saves over the DMMA kernel.
static int hh_tc_min() { const char* e = getenv("GSI_TC_MIN"); return e ? atoi(e) : 2048; }

static int hh_tc_plan(gsi_ctx* ctx, const HhPlan& pl, HhTcPlan& TC) {
    Workspace& ws = WS(ctx);
    const int tc_min = hh_tc_min();
    if (const char* e = getenv("GSI_TC_SLICES")) TC.S = std::min(8, std::max(6, atoi(e)));
    if (tc_min <= 0 || pl.nodes.empty()) return GSI_OK;
    const int64_t budget = std::min<int64_t>((int64_t)8 << 30, std::max<int64_t>(ctx->ws_limit / 8, (int64_t)256 << 20));
    std::vector<uint8_t> flags(pl.nodes.size(), 0);
    std::vector<DcTcTask> tasks;
    int64_t planes_max = 0, expo_max = 0;
    int ntasks_max = 0;
    for (int l = 1; l <= pl.Lmax; ++l) {
        HhTcBatch cur{l, (int)tasks.size(), 0, 0, 0, 0};
        int64_t used = 0, eused = 0;
        auto flush = [&]() {
            if (cur.ntasks > 0) {
                TC.batches.push_back(cur);
                planes_max = std::max(planes_max, used); expo_max = std::max(expo_max, eused); ntasks_max = std::max(ntasks_max, cur.ntasks);
            }
            cur = HhTcBatch{l, (int)tasks.size(), 0, 0, 0, 0};
            used = 0; eused = 0;
        };
        for (int nd = pl.lvl_begin[l]; nd < pl.lvl_begin[l + 1]; ++nd) {
            const DcNode& N = pl.nodes[nd];
            const int m = N.n1 + N.n2;
            if (m < tc_min) continue;
            int64_t need = 0;
            for (int h = 0; h < 2; ++h)
                need += (int64_t)tc_gemm_plane_bytes_a(h ? N.n2 : N.n1, m, TC.S) + (int64_t)tc_gemm_plane_bytes_b(m, m, TC.S);
            if (need > budget) continue;                                    // stays on the DMMA kernel
            if (used + need > budget) flush();
            flags[nd] = 1;
            for (int h = 0; h < 2; ++h) {
                const int M = h ? N.n2 : N.n1;
                DcTcTask t;
                t.node = nd; t.bottom = h; t.ap_off = used; used += (int64_t)tc_gemm_plane_bytes_a(M, m, TC.S);
                t.bp_off = used; used += (int64_t)tc_gemm_plane_bytes_b(m, m, TC.S);
                t.e_off = eused; eused += M + m;
                tasks.push_back(t);
                ++cur.ntasks;
                cur.Mmax = std::max(cur.Mmax, M); cur.Nmax = std::max(cur.Nmax, m); cur.Kmax = std::max(cur.Kmax, m);
            }
        }
        flush();
    }
    if (tasks.empty()) return GSI_OK;
    int rc;
    MetaBuilder mb;
    const size_t o_f = mb.add(flags), o_t = mb.add(tasks);
    if ((rc = ws.tc_meta.ensure(ctx, mb.host.size())) != GSI_OK) return rc;
    if ((rc = ws.tc_tasks.ensure(ctx, (size_t)(ntasks_max + 1) * sizeof(TcTask))) != GSI_OK) return rc;
    if ((rc = ws.tc_planes.ensure(ctx, (size_t)planes_max)) != GSI_OK) return rc;
    if ((rc = ws.tc_expo.ensure(ctx, (size_t)expo_max * 4)) != GSI_OK) return rc;
    GSI_CUDA(ctx, cudaMemcpyAsync(ws.tc_meta.p, mb.host.data(), mb.host.size(), cudaMemcpyHostToDevice, ctx->stream));
    GSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));                       // mb.host is a local
    TC.d_flags = (const uint8_t*)(ws.tc_meta.as<char>() + o_f);
    TC.d_tasks = (const DcTcTask*)(ws.tc_meta.as<char>() + o_t);
    TC.d_tc = ws.tc_tasks.as<TcTask>(); TC.d_planes = ws.tc_planes.as<int8_t>(); TC.d_expo = ws.tc_expo.as<int32_t>();
    if (ctx->trace) fprintf(stderr, "[gsi trace] tcgen05 merges: %zu tasks in %zu batches, planes %.2f GB, %d slices, nodes >= %d\n", tasks.size(),
                            TC.batches.size(), planes_max / 1073741824.0, TC.S, tc_min);
    return GSI_OK;
}

// ---- back-transformation of the big users on the tensor-core engine (kern_bt_tc.cuh) ------------------------------------------
// users with n >= this take it (GSI_BT_TC_MIN; 0 = never, the default).  Jobs are sorted by n descending: they are a prefix of the
// chunk.  Measured on the ML-10M shape (profiles/r02n_bt_tc.md): with GSI_BT_TC_MIN=3072 the 49 biggest users' back-transform takes
// 399 ms on the engine against ~460 ms on the DMMA kernel (bt class 935 -> 875 ms per step, step -1.0 %), but the host path loses
// the overlap of their records' copy with the back-transform of the next group (e2e 12,146 -> 11,858 users/s).  Not a default yet:
// the K = 512 update GEMM re-reads 768 KB of INT8 planes per 128 x 64 tile and sits on the L2 -> SM bandwidth, not on the MMA rate.
static int hh_bt_tc_min() { const char* e = getenv("GSI_BT_TC_MIN"); return e ? atoi(e) : 0; }

static int hh_bt_tc(gsi_ctx* ctx, const HhPlan& pl, const HhDev& D) {
    Workspace& ws = WS(ctx);
    cudaStream_t st = ctx->stream;
    const int tmin = hh_bt_tc_min();
    int nu = 0;
    if (tmin > 0) while (nu < pl.nj && pl.jobs[nu].n >= tmin) ++nu;
    GSI_CUDA(ctx, cudaMemsetAsync(D.bt_skip, 0, (size_t)pl.nj, st));
    if (nu == 0) return GSI_OK;
    int S = 8;
    if (const char* e = getenv("GSI_TC_SLICES")) S = std::min(8, std::max(6, atoi(e)));
    const int64_t budget = std::min<int64_t>((int64_t)8 << 30, std::max<int64_t>(ctx->ws_limit / 8, (int64_t)256 << 20));
    // ---- super-panels
    std::vector<BttSp> sps;
    std::vector<int> sp_first(nu + 1, 0);
    int max_nsp = 0;
    for (int u = 0; u < nu; ++u) {
        const int nrefl = pl.jobs[u].n - 1;
        sp_first[u] = (int)sps.size();
        for (int j0 = 0; j0 < nrefl; j0 += BTT_NB) sps.push_back(BttSp{u, j0, std::min(BTT_NB, nrefl - j0), 0, (int64_t)sps.size() * BTT_NB * BTT_NB});
        max_nsp = std::max(max_nsp, (int)sps.size() - sp_first[u]);
    }
    sp_first[nu] = (int)sps.size();
    const int nsp = (int)sps.size();
    int rc;
    if ((rc = ws.btt_gt.ensure(ctx, (size_t)nsp * BTT_NB * BTT_NB * 8)) != GSI_OK) return rc;
    std::vector<int64_t> x_off(nu + 1, 0);
    for (int u = 0; u < nu; ++u) x_off[u + 1] = x_off[u] + (int64_t)BTT_NB * pl.jobs[u].np;
    if ((rc = ws.btt_x.ensure(ctx, (size_t)x_off[nu] * 8)) != GSI_OK) return rc;
    double* GT = ws.btt_gt.as<double>();
    double* X = ws.btt_x.as<double>();
    // ---- task list: batches of (tasks..., sentinel); planes / exponents are sub-allocated per batch from shared buffers
    struct Batch { int first, ntasks, Mmax, Nmax, Kmax; };
    std::vector<TcTask> tasks;
    std::vector<int32_t> task_job;
    std::vector<Batch> b_g, b_vt;
    std::vector<std::vector<Batch>> b_x(max_nsp), b_u(max_nsp);
    int64_t planes_max = 0, expo_max = 0;
    int8_t* planes = nullptr;       // filled in after the buffers exist: offsets are kept in the pointer fields until then
    struct Open { Batch b; int64_t used, eused; };
    auto begin = [&]() { Open o; o.b = Batch{(int)tasks.size(), 0, 0, 0, 0}; o.used = 0; o.eused = 0; return o; };
    auto close = [&](Open& o, std::vector<Batch>& dst) {
        if (o.b.ntasks == 0) return;
        TcTask z; memset(&z, 0, sizeof z);
        tasks.push_back(z); task_job.push_back(-1);                       // sentinel
        dst.push_back(o.b);
        planes_max = std::max(planes_max, o.used); expo_max = std::max(expo_max, o.eused);
        o = begin();
    };
    auto add = [&](Open& o, std::vector<Batch>& dst, TcTask t, int job_for_n, int Nbound) {
        const int64_t need = (int64_t)tc_gemm_plane_bytes_a(t.M, t.K, S) + (int64_t)tc_gemm_plane_bytes_b(t.K, Nbound, S);
        if (o.b.ntasks > 0 && o.used + need > budget) close(o, dst);
        t.Ap = (int8_t*)(intptr_t)o.used; o.used += (int64_t)tc_gemm_plane_bytes_a(t.M, t.K, S);
        t.Bp = (int8_t*)(intptr_t)o.used; o.used += (int64_t)tc_gemm_plane_bytes_b(t.K, Nbound, S);
        t.ea = (int32_t*)(intptr_t)(o.eused * 4); t.eb = (int32_t*)(intptr_t)((o.eused + t.M) * 4); o.eused += t.M + Nbound;
        t.N = Nbound;                                                      // replaced by kuser on the device where job_for_n >= 0
        tasks.push_back(t); task_job.push_back(job_for_n);
        ++o.b.ntasks;
        o.b.Mmax = std::max(o.b.Mmax, t.M); o.b.Nmax = std::max(o.b.Nmax, Nbound); o.b.Kmax = std::max(o.b.Kmax, t.K);
    };
    auto vfull = [&](int u) { return ((pl.jobs[u].levels & 1) ? D.Qa : D.Qb) + pl.jobs[u].m_off; };
    auto zbuf = [&](int u) { return ((pl.jobs[u].levels & 1) ? D.Qb : D.Qa) + pl.jobs[u].m_off; };
    {   // G = V^T V and VT = V T of every super-panel
        Open og = begin();
        for (const BttSp& sp : sps) {
            const HJob& jb = pl.jobs[sp.job];
            TcTask t; memset(&t, 0, sizeof t);
            const double* Vb = vfull(sp.job) + (size_t)sp.j0 * jb.np + sp.j0;
            t.A = Vb; t.lda = jb.np; t.flags = TC_TRANS_A | TC_UNIT_A | TC_UNIT_B; t.B = Vb; t.ldb = jb.np; t.C = GT + sp.g_off; t.ldc = BTT_NB;
            t.M = sp.nb; t.K = jb.n - sp.j0;
            add(og, b_g, t, -1, sp.nb);
        }
        close(og, b_g);
        Open ov = begin();
        for (const BttSp& sp : sps) {
            const HJob& jb = pl.jobs[sp.job];
            TcTask t; memset(&t, 0, sizeof t);
            t.A = vfull(sp.job) + (size_t)sp.j0 * jb.np + sp.j0; t.lda = jb.np; t.flags = TC_UNIT_A; t.B = GT + sp.g_off; t.ldb = BTT_NB;
            t.C = D.S + jb.m_off + (size_t)sp.j0 * jb.np + sp.j0; t.ldc = jb.np;
            t.M = jb.n - sp.j0; t.K = sp.nb;
            add(ov, b_vt, t, -1, sp.nb);
        }
        close(ov, b_vt);
    }
    for (int s = 0; s < max_nsp; ++s) {   // step s: every user's s-th super-panel from the end
        Open ox = begin();
        for (int u = 0; u < nu; ++u) {
            const int cnt = sp_first[u + 1] - sp_first[u];
            if (s >= cnt) continue;
            const BttSp& sp = sps[sp_first[u] + cnt - 1 - s];
            const HJob& jb = pl.jobs[u];
            TcTask t; memset(&t, 0, sizeof t);
            t.A = vfull(u) + (size_t)sp.j0 * jb.np + sp.j0; t.lda = jb.np; t.flags = TC_TRANS_A | TC_UNIT_A | TC_UNIT_B;
            t.B = zbuf(u) + sp.j0; t.ldb = jb.np; t.C = X + x_off[u]; t.ldc = BTT_NB;
            t.M = sp.nb; t.K = jb.n - sp.j0;
            add(ox, b_x[s], t, u, jb.n);
        }
        close(ox, b_x[s]);
        Open ou = begin();
        for (int u = 0; u < nu; ++u) {
            const int cnt = sp_first[u + 1] - sp_first[u];
            if (s >= cnt) continue;
            const BttSp& sp = sps[sp_first[u] + cnt - 1 - s];
            const HJob& jb = pl.jobs[u];
            TcTask t; memset(&t, 0, sizeof t);
            t.A = D.S + jb.m_off + (size_t)sp.j0 * jb.np + sp.j0; t.lda = jb.np;
            t.B = X + x_off[u]; t.ldb = BTT_NB; t.C = zbuf(u) + sp.j0; t.ldc = jb.np; t.flags = TC_SUB_C;
            t.M = jb.n - sp.j0; t.K = sp.nb;
            add(ou, b_u[s], t, u, jb.n);
        }
        close(ou, b_u[s]);
    }
    if ((rc = ws.btt_planes.ensure(ctx, (size_t)planes_max)) != GSI_OK) return rc;
    if ((rc = ws.btt_expo.ensure(ctx, (size_t)expo_max * 4)) != GSI_OK) return rc;
    planes = ws.btt_planes.as<int8_t>();
    int32_t* expo = ws.btt_expo.as<int32_t>();
    for (size_t t = 0; t < tasks.size(); ++t) {
        if (!tasks[t].A) continue;                                          // sentinel
        tasks[t].Ap = planes + (intptr_t)tasks[t].Ap; tasks[t].Bp = planes + (intptr_t)tasks[t].Bp;
        tasks[t].ea = (int32_t*)((char*)expo + (intptr_t)tasks[t].ea); tasks[t].eb = (int32_t*)((char*)expo + (intptr_t)tasks[t].eb);
    }
    MetaBuilder mb;
    const size_t o_t = mb.add(tasks), o_j = mb.add(task_job), o_sp = mb.add(sps);
    if ((rc = ws.btt_meta.ensure(ctx, mb.host.size())) != GSI_OK) return rc;
    GSI_CUDA(ctx, cudaMemcpyAsync(ws.btt_meta.p, mb.host.data(), mb.host.size(), cudaMemcpyHostToDevice, st));
    GSI_CUDA(ctx, cudaStreamSynchronize(st));                                 // mb.host is a local
    TcTask* d_tasks = (TcTask*)(ws.btt_meta.as<char>() + o_t);
    const int32_t* d_tjob = (const int32_t*)(ws.btt_meta.as<char>() + o_j);
    const BttSp* d_sps = (const BttSp*)(ws.btt_meta.as<char>() + o_sp);
    auto run = [&](const Batch& b) -> int {
        TcBatch B;
        B.tasks = d_tasks + b.first; B.ntasks = b.ntasks; B.Mmax = b.Mmax; B.Nmax = b.Nmax; B.Kmax = b.Kmax; B.S = S;
        GSI_CUDA(ctx, tc_gemm_batch(B, st, ctx->sm_count));
        return GSI_OK;
    };
    size_t nbatches = b_g.size() + b_vt.size();
    for (int s = 0; s < max_nsp; ++s) nbatches += b_x[s].size() + b_u[s].size();
    GsiSpan span(ctx, GSI_T_BT, 3 + 7 * (int64_t)nbatches);
    HhTrace tr(ctx, "bt on the tcgen05 engine");
    if (ctx->trace) fprintf(stderr, "[gsi trace] tcgen05 back-transform: %d users n >= %d, %d super-panels, %zu tasks in %zu batches, planes %.2f GB\n", nu,
                            tmin, nsp, tasks.size(), nbatches, planes_max / 1073741824.0);
    GSI_CUDA(ctx, cudaMemsetAsync(D.bt_skip, 1, (size_t)nu, st));
    btt_set_n_kernel<<<((int)tasks.size() + 255) / 256, 256, 0, st>>>(d_tasks, d_tjob, (int)tasks.size(), D.kuser);
    const int NTmax = pl.jobs[0].np >> 6;
    btt_extract_kernel<<<dim3(NTmax * NTmax, nu), 256, 0, st>>>(D.jobs, nu, NTmax, D.A, D.Qa, D.Qb);
    GSI_CUDA(ctx, cudaMemsetAsync(GT, 0, (size_t)nsp * BTT_NB * BTT_NB * 8, st));
    for (const Batch& b : b_g) if ((rc = run(b)) != GSI_OK) return rc;
    GSI_CUDA(ctx, cudaFuncSetAttribute(btt_formt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)btt_formt_smem_bytes()));
    btt_formt_kernel<<<nsp, 256, btt_formt_smem_bytes(), st>>>(D.jobs, d_sps, D.tau, GT);
    for (const Batch& b : b_vt) if ((rc = run(b)) != GSI_OK) return rc;
    for (int s = 0; s < max_nsp; ++s) {
        for (const Batch& b : b_x[s]) if ((rc = run(b)) != GSI_OK) return rc;
        for (const Batch& b : b_u[s]) if ((rc = run(b)) != GSI_OK) return rc;
    }
    span.end();
    GSI_CUDA(ctx, cudaGetLastError());
    return GSI_OK;
}

static int hh_solve(gsi_ctx* ctx, const HhPlan& pl, const HhDev& D, int forced_team, bool defer_bt_apply = false) {
    cudaStream_t st = ctx->stream;
    const int nj = pl.nj;
    int rc;
    HhSbrState SB;
    if ((rc = hh_trd(ctx, pl, D, forced_team, SB)) != GSI_OK) return rc;
    // ---------------- divide & conquer
    DcParams P;
    P.jobs = D.jobs; P.nodes = D.nodes; P.state = D.state; P.Qa = D.Qa; P.Qb = D.Qb; P.S = D.S; P.lamA = D.lamA; P.lamB = D.lamB;
    P.e = D.e; P.dk = D.dk; P.zk = D.zk; P.zhat = D.zhat; P.tau = D.dctau; P.lamk = D.lamk; P.dval = D.dval;
    P.orig = D.orig; P.gmap = D.gmap; P.colsrc = D.colsrc; P.dsrc = D.dsrc; P.pos_nd = D.pos_nd; P.pos_df = D.pos_df;
    P.ds = D.ds; P.zs = D.zs; P.src = D.src; P.rot = D.rot; P.sigmax = D.sigmax; P.kuser = D.kuser;
    {
        GsiSpan sp(ctx, GSI_T_DC, 1);
        const int nl = (int)pl.leaves.size();
        HhTrace tr(ctx, "dc_leaf");
        dc_leaf_kernel<<<(nl + 3) / 4, 128, 0, st>>>(D.jobs, D.leaves, nl, D.d, D.e, D.lamA, D.Qa);
        sp.end();
        GSI_CUDA(ctx, cudaGetLastError());
    }
    const size_t gemm_smem = (size_t)DCG_STAGES * DCG_STAGE_DBL * sizeof(double);
    GSI_CUDA(ctx, cudaFuncSetAttribute(dc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem));
    // ---- big merges on the tcgen05 engine (tc_gemm.cu): planned here from the node sizes (upper bounds), in sub-batches that fit
    // the plane budget; everything is uploaded once, the level loop only launches
    HhTcPlan TC;
    if ((rc = hh_tc_plan(ctx, pl, TC)) != GSI_OK) return rc;
    P.tc_node = TC.d_flags;
    for (int l = 1; l <= pl.Lmax; ++l) {
        P.node0 = pl.lvl_begin[l]; P.nnodes = pl.lvl_begin[l + 1] - pl.lvl_begin[l]; P.in_b = (l - 1) & 1;
        const int mmax = pl.lvl_mmax[l];
        {
            GsiSpan sp(ctx, GSI_T_DC, 6);
            if (ctx->trace) fprintf(stderr, "[gsi trace] dc level %d: %d nodes, mmax %d\n", l, P.nnodes, mmax);
            {
                HhTrace tr(ctx, "  deflate");
                const int in_global = mmax > DC_DEFLATE_SMEM_MAX_M;
                const size_t dsm = in_global ? 64 : dc_deflate_smem_bytes(mmax);
                GSI_CUDA(ctx, cudaFuncSetAttribute(dc_deflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(dsm, 48 * 1024)));
                dc_deflate_kernel<<<P.nnodes, 256, dsm, st>>>(P, in_global);
            }
            {
                HhTrace tr(ctx, "  secular");       // lanes per root grow with the merge size (few, big merges near the top)
                if (mmax <= 256) dc_secular_kernel<1><<<dim3(P.nnodes, (mmax + 127) / 128), 128, 0, st>>>(P);
                else if (mmax <= 1024) dc_secular_kernel<8><<<dim3(P.nnodes, (mmax + 15) / 16), 128, 0, st>>>(P);
                else dc_secular_kernel<32><<<dim3(P.nnodes, (mmax + 3) / 4), 128, 0, st>>>(P);
            }
            { HhTrace tr(ctx, "  rank"); dc_rank_kernel<<<P.nnodes, 256, 0, st>>>(P); }
            { HhTrace tr(ctx, "  zhat"); dc_zhat_kernel<<<dim3(P.nnodes, (mmax + 7) / 8), 256, 0, st>>>(P); }
            { HhTrace tr(ctx, "  vectors"); dc_vectors_kernel<<<dim3(P.nnodes, (mmax + 7) / 8), 256, 0, st>>>(P); }
            dc_plan_kernel<<<1, 1024, 0, st>>>(P, D.tile_off);
            sp.end();
        }
        {
            int ntcb = 0;
            for (const HhTcBatch& tb : TC.batches) ntcb += (tb.level == l);
            GsiSpan sp(ctx, GSI_T_DC_GEMM, 2 + 8 * ntcb);          // per tcgen05 batch: resolve, 5 slicing kernels, tile scan, GEMM
            { HhTrace tr(ctx, "  gemm"); dc_gemm_kernel<<<2 * ctx->sm_count, 256, gemm_smem, st>>>(P, D.tile_off); }
            for (const HhTcBatch& tb : TC.batches) {
                if (tb.level != l) continue;
                HhTrace tr(ctx, "  gemm (tcgen05 batch)");
                dc_tc_resolve_kernel<<<(tb.ntasks + 1 + 127) / 128, 128, 0, st>>>(P, TC.d_tasks + tb.first, tb.ntasks, TC.d_planes, TC.d_expo, TC.d_tc);
                TcBatch B;
                B.tasks = TC.d_tc; B.ntasks = tb.ntasks; B.Mmax = tb.Mmax; B.Nmax = tb.Nmax; B.Kmax = tb.Kmax; B.S = TC.S;
                GSI_CUDA(ctx, tc_gemm_batch(B, st, ctx->sm_count));
            }
            { HhTrace tr(ctx, "  copy"); dc_copy_kernel<<<dim3(P.nnodes, (mmax + 7) / 8), 256, 0, st>>>(P); }
            sp.end();
        }
        GSI_CUDA(ctx, cudaGetLastError());
    }
    // ---------------- back-transformation: Q2 (bulge-chasing reflectors) for the two-stage users first, then Q1 / Q for everybody
    if (SB.nu > 0 && (rc = sbr_bt2(ctx, D, SB.st2)) != GSI_OK) return rc;
    BtParams B;
    B.jobs = D.jobs; B.njobs = nj; B.A = D.A; B.tau = D.tau; B.S = D.S; B.Qa = D.Qa; B.Qb = D.Qb; B.kuser = D.kuser;
    B.items = D.bt_items; B.nitems = (int)pl.bt_items.size(); B.queue = D.ctl + 8; B.skip = D.bt_skip;
    if ((rc = hh_bt_tc(ctx, pl, D)) != GSI_OK) return rc;            // the big users: aggregated panels on the tcgen05 engine
    {
        GsiSpan sp(ctx, GSI_T_BT, 2);
        GSI_CUDA(ctx, cudaFuncSetAttribute(bt_formt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bt_formt_smem_bytes()));
        GSI_CUDA(ctx, cudaFuncSetAttribute(bt_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bt_smem_bytes()));
        { HhTrace tr(ctx, "bt_formt"); bt_formt_kernel<<<dim3((pl.nmax - 1 + BT_NB - 1) / BT_NB, nj), 256, bt_formt_smem_bytes(), st>>>(B); }
        GSI_CUDA(ctx, cudaMemsetAsync(D.ctl + 8, 0, 4, st));
        if (!defer_bt_apply) { HhTrace tr(ctx, "bt_apply"); bt_apply_kernel<<<ctx->sm_count, 256, bt_smem_bytes(), st>>>(B); }
        sp.end();
        GSI_CUDA(ctx, cudaGetLastError());
    }
    return GSI_OK;
}

static BtParams hh_bt_params(const HhPlan& pl, const HhDev& D) {
    BtParams B;
    B.jobs = D.jobs; B.njobs = pl.nj; B.A = D.A; B.tau = D.tau; B.S = D.S; B.Qa = D.Qa; B.Qb = D.Qb; B.kuser = D.kuser;
    B.items = D.bt_items; B.nitems = (int)pl.bt_items.size(); B.queue = D.ctl + 8; B.skip = D.bt_skip;
    return B;
}

// ---- large chunk through the Householder path --------------------------------------------------------
static int run_hh_chunk(gsi_ctx* ctx, const Job* jobs, int nj, const int32_t* d_items, const RunOut& out) {
    Workspace& ws = WS(ctx);
    cudaStream_t st = ctx->stream;
    HhPlan pl;
    hh_build_plan(jobs, nj, pl);
    HhDev D;
    int rc;
    if ((rc = hh_alloc(ctx, pl, jobs, D)) != GSI_OK) return rc;
    {   // Laplacian stage (shared with the block-Jacobi path): sym(lower(L)) without the +1 shift.  Jobs are
        // sorted by n descending; launch per group of similar size so that the grids match the matrices.
        int ngroups = 0;
        for (int b1 = 0; b1 < nj; ++ngroups) { const int nm = pl.jobs[b1].n; while (b1 < nj && 2 * pl.jobs[b1].n > nm) ++b1; }
        GsiSpan sp(ctx, GSI_T_LAP, (getenv("GSI_LAP_UNFUSED") ? 5 : 2) * ngroups);
        HhTrace tr(ctx, "lap");
        int b0 = 0;
        while (b0 < nj) {
            const int nmax = pl.jobs[b0].n;
            int e0 = b0;
            while (e0 < nj && 2 * pl.jobs[e0].n > nmax) ++e0;
            const int cnt = e0 - b0, tiles = (nmax + 31) / 32;
            LChunk C;
            memset(&C, 0, sizeof C);
            C.nu = cnt; C.n = D.n_arr + b0; C.ld = D.ld_arr + b0; C.g_off = D.moff_arr + b0; C.item_off = D.ioff_arr + b0;
            C.row_off = D.roff_arr + b0;
            C.G = D.A; C.deg = D.deg; C.scale = D.scale; C.sigmax = D.sigmax + b0; C.tiled = 1;
            if (getenv("GSI_LAP_UNFUSED")) {                 // the five-kernel version (shared with the block-Jacobi path), for comparison
                lap_gather_kernel<<<dim3(tiles * tiles, 1, cnt), dim3(32, 8), 0, st>>>(C, ctx->d_w, ctx->w_rows, d_items, tiles);
                lap_degree_kernel<<<dim3((nmax + 127) / 128, 1, cnt), 128, 0, st>>>(C);
                lap_transform_kernel<<<dim3((nmax + 127) / 128, (nmax + 7) / 8, cnt), 128, 0, st>>>(C);
                lap_sigmin_kernel<<<dim3((nmax + 127) / 128, 1, cnt), 128, 0, st>>>(C, out.d_sig_min);
                lap_symmetrize_kernel<<<dim3(tiles * tiles, 1, cnt), dim3(32, 8), 0, st>>>(C, tiles, 0.0);
            } else {
                lap_fused_gather_kernel<<<dim3((nmax + 63) / 64, 1, cnt), 256, 0, st>>>(C, ctx->d_w, ctx->w_rows, d_items);
                lap_fused_transform_kernel<<<dim3((nmax + 63) / 64, 1, cnt), 256, 0, st>>>(C, out.d_sig_min);
            }
            b0 = e0;
        }
        sp.end();
        GSI_CUDA(ctx, cudaGetLastError());
    }
    if ((rc = hh_solve(ctx, pl, D, 0, /*defer_bt_apply=*/true)) != GSI_OK) return rc;
    OutJobs J;
    J.nj = nj; J.n = D.n_arr; J.k = D.kuser; J.user = D.user; J.vec_pad = D.voff_arr; J.lam_pad = D.loff_arr;
    J.vec_dst = D.vec_dst; J.lam_dst = D.lam_dst;
    std::vector<int32_t> h_n(nj);
    for (int j = 0; j < nj; ++j) h_n[j] = pl.jobs[j].n;
    // The kept-eigenpair counts are final after the last merge, so the record offsets are known BEFORE the
    // back-transform.  Device path: one back-transform launch over all users (biggest first: best balance).
    // Host path: back-transform, emit and compaction run per group of users, SMALLEST users first (they hold half
    // of the eigenvector volume but a seventh of the back-transform work; groups cut at 50 / 75 / 90 % of the n^2
    // volume), and the eigenvector block of a group is copied to pinned memory on a second stream while the
    // bigger users are still being back-transformed.
    if ((rc = finish_chunk(ctx, J, out, pl.max_nk, h_n.data(), /*do_scan=*/true, 0, 0)) != GSI_OK) return rc;
    const bool overlap = out.h_vec != nullptr && out.vec_copied != nullptr;
    int gb[5] = {0, 0, 0, 0, nj};                          // job ranges [gb[g], gb[g+1]) in job order (n descending)
    if (overlap) {
        double tot = 0, run = 0;
        for (int j = 0; j < nj; ++j) tot += (double)h_n[j] * h_n[j];
        const double cut[3] = {0.5, 0.75, 0.9};
        int g = 0;
        gb[1] = gb[2] = gb[3] = 0;
        for (int j = nj - 1; j >= 0 && g < 3; --j) {       // from the small end
            run += (double)h_n[j] * h_n[j];
            if (run >= cut[g] * tot) { gb[3 - g] = j; ++g; }
        }
    }
    std::vector<int> item0(nj + 1, 0);                     // bt work items are laid out job by job (hh_build_plan)
    for (int j = 0; j < nj; ++j) item0[j + 1] = item0[j] + (pl.jobs[j].n + BT_CB - 1) / BT_CB;
    cudaEvent_t ev_scan = nullptr, ev_grp[4] = {nullptr, nullptr, nullptr, nullptr};
    if (overlap) {
        if (!ctx->copy_stream) GSI_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (int g = 0; g < 4; ++g)                        // offset of the first record of every group, then the end
            GSI_CUDA(ctx, cudaMemcpyAsync(out.h_bounds + g, gb[g] < nj ? (const void*)(D.vec_dst + gb[g]) : (const void*)(out.d_totals + 1), 8,
                                          cudaMemcpyDeviceToHost, st));
        GSI_CUDA(ctx, cudaMemcpyAsync(out.h_bounds + 4, out.d_totals + 1, 8, cudaMemcpyDeviceToHost, st));
        GSI_CUDA(ctx, cudaEventCreateWithFlags(&ev_scan, cudaEventDisableTiming));
        GSI_CUDA(ctx, cudaEventRecord(ev_scan, st));
    }
    BtParams B = hh_bt_params(pl, D);
    for (int g = 3; g >= 0; --g) {                         // smallest users first
        const int jb = gb[g], je = gb[g + 1];
        if (jb >= je) continue;
        {
            GsiSpan sp(ctx, GSI_T_BT, 1);
            BtParams Bg = B;
            Bg.items = B.items + item0[jb]; Bg.nitems = item0[je] - item0[jb];
            GSI_CUDA(ctx, cudaMemsetAsync(D.ctl + 8, 0, 4, st));
            HhTrace tr(ctx, "bt_apply");
            bt_apply_kernel<<<ctx->sm_count, 256, bt_smem_bytes(), st>>>(Bg);
            sp.end();
            GSI_CUDA(ctx, cudaGetLastError());
        }
        {
            GsiSpan sp(ctx, GSI_T_BT, 2);
            HhTrace tr(ctx, "emit");
            for (int b0 = jb; b0 < je;) {                 // sub-groups of similar size: grids that match the matrices
                const int nmax = pl.jobs[b0].n;
                int e0 = b0;
                while (e0 < je && 2 * pl.jobs[e0].n > nmax) ++e0;
                const int tiles = (nmax + 31) / 32;
                emit_sign_kernel<<<dim3((nmax + 7) / 8, e0 - b0), 256, 0, st>>>(B, D.sgn, b0);
                emit_vec_kernel<<<dim3(tiles * tiles, 1, e0 - b0), dim3(32, 8), 0, st>>>(B, D.sgn, D.lamA, D.lamB, ws.vec_pad.as<double>(),
                                                                                   ws.lam_pad.as<double>(), tiles, b0);
                b0 = e0;
            }
            sp.end();
            GSI_CUDA(ctx, cudaGetLastError());
        }
        if ((rc = finish_chunk(ctx, J, out, pl.max_nk, h_n.data(), /*do_scan=*/false, jb, je)) != GSI_OK) return rc;
        if (overlap) {
            GSI_CUDA(ctx, cudaEventCreateWithFlags(&ev_grp[g], cudaEventDisableTiming));
            GSI_CUDA(ctx, cudaEventRecord(ev_grp[g], st));
        }
    }
    if (overlap) {
        GSI_CUDA(ctx, cudaEventSynchronize(ev_scan));      // the offsets are on the host; the groups above are already queued
        for (int g = 3; g >= 0; --g) {
            if (!ev_grp[g]) continue;
            const int64_t b0 = out.h_bounds[g], e0 = out.h_bounds[g + 1];
            GSI_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ev_grp[g], 0));
            if (e0 > b0 && e0 <= out.vec_cap)
                GSI_CUDA(ctx, cudaMemcpyAsync(out.h_vec + b0, out.d_vec + b0, (size_t)(e0 - b0) * 8, cudaMemcpyDeviceToHost, ctx->copy_stream));
        }
        *out.vec_copied = true;                            // the caller synchronises copy_stream
        cudaEventDestroy(ev_scan);
        for (int g = 0; g < 4; ++g) if (ev_grp[g]) cudaEventDestroy(ev_grp[g]);
    }
    return GSI_OK;
}

// ---- stage-wise test hook ---------------------------------------------------------------------------
extern "C" int gsi_debug_eigh(gsi_ctx* ctx, int n, const double* a, float thr, int team, double* d, double* e, double* tau,
                              double* v, double* lam, double* u, int32_t* k) {
    if (!ctx) return gsi_fail(nullptr, GSI_ERR_INVALID, "null context");
    if (!a || n < 33 || n > 17664) return gsi_fail(ctx, GSI_ERR_INVALID, "gsi_debug_eigh: need a matrix with 33 <= n <= 17664");
    GSI_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    Job job{0, n, 0};
    HhPlan pl;
    hh_build_plan(&job, 1, pl);
    HhDev D;
    int rc;
    if ((rc = hh_alloc(ctx, pl, &job, D)) != GSI_OK) return rc;
    const int np = pl.jobs[0].np;
    const int NT = np >> 6;
    std::vector<double> tiled((size_t)np * np, 0.0);
    for (int cc = 0; cc < n; ++cc)
        for (int r = 0; r < n; ++r) tiled[hh_tidx(r, cc, NT)] = a[(size_t)cc * n + r];
    GSI_CUDA(ctx, cudaMemcpyAsync(D.A, tiled.data(), tiled.size() * 8, cudaMemcpyHostToDevice, st));
    // the cutoff kernel computes (float)(sigmax + 0.01); hand it sigmax = thr - 0.01 (test hook only)
    const float sm = thr - 0.01f;
    GSI_CUDA(ctx, cudaMemcpyAsync(D.sigmax, &sm, 4, cudaMemcpyHostToDevice, st));
    GSI_CUDA(ctx, cudaStreamSynchronize(st));
    if ((rc = hh_solve(ctx, pl, D, team)) != GSI_OK) return rc;
    GSI_CUDA(ctx, cudaStreamSynchronize(st));
    int32_t kk = 0;
    GSI_CUDA(ctx, cudaMemcpy(&kk, D.kuser, 4, cudaMemcpyDeviceToHost));
    if (k) *k = kk;
    if (d) GSI_CUDA(ctx, cudaMemcpy(d, D.d, (size_t)n * 8, cudaMemcpyDeviceToHost));
    if (e) GSI_CUDA(ctx, cudaMemcpy(e, D.e, (size_t)(n - 1) * 8, cudaMemcpyDeviceToHost));
    if (tau) GSI_CUDA(ctx, cudaMemcpy(tau, D.tau, (size_t)(n - 1) * 8, cudaMemcpyDeviceToHost));
    if (v) {
        GSI_CUDA(ctx, cudaMemcpy(tiled.data(), D.A, tiled.size() * 8, cudaMemcpyDeviceToHost));
        for (int cc = 0; cc < n; ++cc)
            for (int r = 0; r < n; ++r) v[(size_t)cc * n + r] = tiled[hh_tidx(r, cc, NT)];
    }
    const bool in_b = pl.jobs[0].levels & 1;
    if (lam) GSI_CUDA(ctx, cudaMemcpy(lam, in_b ? D.lamB : D.lamA, (size_t)n * 8, cudaMemcpyDeviceToHost));
    if (u && kk > 0) GSI_CUDA(ctx, cudaMemcpy2D(u, (size_t)n * 8, in_b ? D.Qb : D.Qa, (size_t)np * 8, (size_t)n * 8, kk, cudaMemcpyDeviceToHost));
    return GSI_OK;
}

// ---- stage-wise test hook of the two-stage path: the band after stage 1 ------------------------------------------------
extern "C" int gsi_debug_band(gsi_ctx* ctx, int n, const double* a, double* ab /* [n * 128] */, double* d, double* e) {
    if (!ctx) return gsi_fail(nullptr, GSI_ERR_INVALID, "null context");
    if (!a || !ab || n < 130 || n > 17664) return gsi_fail(ctx, GSI_ERR_INVALID, "gsi_debug_band: need a matrix with 130 <= n <= 17664");
    GSI_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    Job job{0, n, 0};
    HhPlan pl;
    hh_build_plan(&job, 1, pl);
    HhDev D;
    int rc;
    if ((rc = hh_alloc(ctx, pl, &job, D)) != GSI_OK) return rc;
    const int np = pl.jobs[0].np, NT = np >> 6;
    std::vector<double> tiled((size_t)np * np, 0.0);
    for (int cc = 0; cc < n; ++cc)
        for (int r = 0; r < n; ++r) tiled[hh_tidx(r, cc, NT)] = a[(size_t)cc * n + r];
    GSI_CUDA(ctx, cudaMemcpyAsync(D.A, tiled.data(), tiled.size() * 8, cudaMemcpyHostToDevice, st));
    GSI_CUDA(ctx, cudaStreamSynchronize(st));
    HhSbrState SB;
    SB.nu = 1;
    if ((rc = sbr_alloc(ctx, pl, 1, SB.dev)) != GSI_OK) return rc;
    if ((rc = sbr_stage1(ctx, pl, D, SB.dev)) != GSI_OK) return rc;
    GSI_CUDA(ctx, cudaMemcpyAsync(ab, SB.dev.AB, (size_t)n * SBR_LDB * 8, cudaMemcpyDeviceToHost, st));
    GSI_CUDA(ctx, cudaStreamSynchronize(st));
    if (d || e) {
        if ((rc = sbr_stage2(ctx, pl, D, SB.dev, SB.st2)) != GSI_OK) return rc;
        if (d) GSI_CUDA(ctx, cudaMemcpyAsync(d, D.d, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        if (e) GSI_CUDA(ctx, cudaMemcpyAsync(e, D.e, (size_t)(n - 1) * 8, cudaMemcpyDeviceToHost, st));
        GSI_CUDA(ctx, cudaStreamSynchronize(st));
    }
    return GSI_OK;
}
