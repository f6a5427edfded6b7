// Host side of the two-stage tridiagonalisation (kern_sbr.cuh).  Included by hh_host.cuh.
#pragma once

#define SBR_SEG_MAX 4

struct SbrDev {                 // device buffers of the SBR users of a chunk (a prefix of the chunk's jobs: sorted by n descending)
    int nu = 0;                 // users on the two-stage path
    int64_t rtot = 0;           // rows (np) of those users = extent of the r_off-indexed buffers
    int* users = nullptr;       // [nu] job indices 0 .. nu-1
    double *Vp, *Wp, *Xp, *Yp, *Sp, *T1, *TS, *qr_part, *AB;
    unsigned* qr_bar;
    int parts_cap = 0;
};

// users with n >= this take the two-stage path (GSI_SBR_MIN overrides; 0 = never).  Default: the users the one-stage kernel
// cannot take (n > GSI_HH_MAX_N) -- measured on the ML-10M shape (profiles/r02c_sbr_threshold.md) the two-stage path does not
// yet beat the one-stage kernel below that: its bulge chase is bound by the 3-task lag between consecutive sweeps and the
// second back-transform costs 1.8 n^3 more flop on an FP64 pipe that is only ~5 flop/byte away from the HBM roofline.
static int sbr_min_n() {
    const char* e = getenv("GSI_SBR_MIN");               // read on every call (tests switch it)
    const int v = e ? atoi(e) : GSI_HH_MAX_N + 1;
    return v <= 0 ? 0 : std::max(v, 130);
}

static int sbr_alloc(gsi_ctx* ctx, const HhPlan& pl, int nu, SbrDev& S) {
    Workspace& ws = WS(ctx);
    int rc;
    S.nu = nu;
    S.rtot = pl.jobs[nu - 1].r_off + pl.jobs[nu - 1].np;
    S.parts_cap = (pl.jobs[0].np + SBR_RB - 1) / SBR_RB + 1;
    const size_t panel = (size_t)S.rtot * 64;
    if ((rc = ws.sbr_panels.ensure(ctx, panel * (4 + SBR_SEG_MAX) * 8)) != GSI_OK) return rc;
    if ((rc = ws.sbr_small.ensure(ctx, ((size_t)nu * 2 * 4096 + (size_t)nu * S.parts_cap * SBR_QP) * 8 + (size_t)nu * 8)) != GSI_OK) return rc;
    if ((rc = ws.sbr_band.ensure(ctx, (size_t)S.rtot * SBR_LDB * 8)) != GSI_OK) return rc;
    double* b = ws.sbr_panels.as<double>();
    S.Vp = b; S.Wp = b + panel; S.Xp = b + 2 * panel; S.Sp = b + 3 * panel; S.Yp = b + 4 * panel;
    double* s = ws.sbr_small.as<double>();
    S.T1 = s; S.TS = s + (size_t)nu * 4096; S.qr_part = s + (size_t)nu * 2 * 4096;
    S.qr_bar = (unsigned*)(S.qr_part + (size_t)nu * S.parts_cap * SBR_QP);
    S.users = (int*)(S.qr_bar + nu);
    S.AB = ws.sbr_band.as<double>();
    std::vector<int> ids(nu);
    for (int i = 0; i < nu; ++i) ids[i] = i;
    GSI_CUDA(ctx, cudaMemcpyAsync(S.users, ids.data(), (size_t)nu * 4, cudaMemcpyHostToDevice, ctx->stream));
    GSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));         // ids is a stack-lifetime host buffer
    return GSI_OK;
}

// GSI_TRACE: per-kernel-kind stopwatch of stage 1 (synchronising; diagnostics only)
struct SbrKindTimer {
    gsi_ctx* ctx; bool on; int cur = -1; cudaEvent_t a = nullptr, b = nullptr; double ms[4] = {0, 0, 0, 0}; int waves = 0; int64_t panels = 0;
    explicit SbrKindTimer(gsi_ctx* c) : ctx(c), on(c->trace) { if (on) { cudaEventCreate(&a); cudaEventCreate(&b); } }
    void go(int kind) {
        if (!on) return;
        if (cur >= 0) { cudaEventRecord(b, ctx->stream); cudaEventSynchronize(b); float t = 0.f; cudaEventElapsedTime(&t, a, b); ms[cur] += t; }
        cur = kind;
        if (kind >= 0) cudaEventRecord(a, ctx->stream);
    }
    ~SbrKindTimer() {
        if (!on) return;
        fprintf(stderr, "[gsi trace] sbr stage 1 kinds: waves %d panel-steps %lld  qr %.1f  symm %.1f  w123 %.1f  syr2k %.1f ms\n", waves,
                (long long)panels, ms[0], ms[1], ms[2], ms[3]);
        cudaEventDestroy(a); cudaEventDestroy(b);
    }
};

// stage 1 for the users [0, nu) of the chunk: dense -> band (in A) + reflectors (in A, tau); AB = compact band
static int sbr_stage1(gsi_ctx* ctx, const HhPlan& pl, const HhDev& D, const SbrDev& S) {
    cudaStream_t st = ctx->stream;
    const int sms = ctx->sm_count;
    GSI_CUDA(ctx, cudaFuncSetAttribute(sbr_panel_qr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbr_qr_smem_bytes()));
    GSI_CUDA(ctx, cudaFuncSetAttribute(sbr_symm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbr_symm_smem_bytes()));
    GSI_CUDA(ctx, cudaFuncSetAttribute(sbr_syr2k_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbr_syr2k_smem_bytes()));
    GSI_CUDA(ctx, cudaFuncSetAttribute(sbr_w1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbr_w_smem_bytes()));
    GSI_CUDA(ctx, cudaFuncSetAttribute(sbr_w2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbr_w_smem_bytes()));
    GSI_CUDA(ctx, cudaFuncSetAttribute(sbr_w3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbr_w_smem_bytes()));
    SbrParams P;
    memset(&P, 0, sizeof P);
    P.jobs = D.jobs; P.A = D.A; P.tau = D.tau; P.Vp = S.Vp; P.Wp = S.Wp; P.Xp = S.Xp; P.Yp = S.Yp; P.Sp = S.Sp; P.T1 = S.T1; P.TS = S.TS;
    P.qr_part = S.qr_part; P.qr_bar = S.qr_bar; P.ystride = S.rtot * 64; P.qr_parts_max = S.parts_cap;
    // tau of the users: zero (panels without a reflector keep 0)
    GSI_CUDA(ctx, cudaMemsetAsync(D.tau, 0, (size_t)S.rtot * 8, st));
    SbrKindTimer tk(ctx);
    int b = 0;
    while (b < S.nu) {
        ++tk.waves;
        // a wave: users whose QR teams are co-resident (one CTA per SM)
        int e = b, parts = 0;
        while (e < S.nu) {
            const int c = (pl.jobs[e].n - 64 + SBR_RB - 1) / SBR_RB;
            if (e > b && parts + c > sms) break;
            parts += std::max(c, 1); ++e;
        }
        const int nmax = pl.jobs[b].n, npmax = pl.jobs[b].np, NTmax = npmax >> 6;
        P.users = S.users + b;
        for (int p = 0; nmax - (p + 1) * 64 >= 2; ++p) {
            int nw = 0;                                         // active users of the wave (a prefix: sorted by n descending)
            int64_t rows_blocks = 0;
            while (b + nw < e && pl.jobs[b + nw].n - (p + 1) * 64 >= 2) { rows_blocks += (pl.jobs[b + nw].np >> 6) - (p + 1); ++nw; }
            P.p = p; ++tk.panels;
            const int m = nmax - (p + 1) * 64, qparts = (m + SBR_RB - 1) / SBR_RB, nbmax = NTmax - (p + 1);
            P.seg = (int)std::min<int64_t>(std::min(SBR_SEG_MAX, nbmax), std::max<int64_t>(1, (2 * sms + rows_blocks - 1) / rows_blocks));
            GsiSpan sp(ctx, GSI_T_SBR, 6);
            GSI_CUDA(ctx, cudaMemsetAsync(S.qr_bar + b, 0, (size_t)nw * 4, st));
            const int64_t tiles = (int64_t)nbmax * (nbmax + 1) / 2;
            const int ctas = (int)std::max<int64_t>(1, std::min<int64_t>(tiles, (4 * sms + nw - 1) / nw));
            tk.go(0); sbr_panel_qr_kernel<<<dim3(qparts, nw), 256, sbr_qr_smem_bytes(), st>>>(P);
            tk.go(1); sbr_symm_kernel<<<dim3(nbmax * P.seg, nw), 256, sbr_symm_smem_bytes(), st>>>(P);
            tk.go(2); sbr_w1_kernel<<<dim3(nbmax, nw), 256, sbr_w_smem_bytes(), st>>>(P);
            sbr_w2_kernel<<<nw, 256, sbr_w_smem_bytes(), st>>>(P);
            sbr_w3_kernel<<<dim3(nbmax, nw), 256, sbr_w_smem_bytes(), st>>>(P);
            tk.go(3); sbr_syr2k_kernel<<<dim3(ctas, nw), 256, sbr_syr2k_smem_bytes(), st>>>(P);
            tk.go(-1);
            sp.end();
            GSI_CUDA(ctx, cudaGetLastError());
        }
        b = e;
    }
    {
        GsiSpan sp(ctx, GSI_T_SBR, 1);
        P.users = S.users;
        sbr_band_kernel<<<dim3(pl.jobs[0].np >> 6, S.nu), 256, 0, st>>>(P, S.AB);
        sp.end();
        GSI_CUDA(ctx, cudaGetLastError());
    }
    return GSI_OK;
}

// ---- stage 2 -----------------------------------------------------------------------------------------------------
struct SbrStage2 {              // device views built by sbr_stage2, used again by sbr_bt2
    double* V2 = nullptr; int64_t nblocks = 0;
    int64_t* v2_off = nullptr; int* goff = nullptr; int* goff_off = nullptr;
    int2* bt_items = nullptr; int n_bt_items = 0;
    int* queue = nullptr;       // [2]: chase, bt2
};

static inline int sbr_ntasks(int n, int G) { const int r = n - 3 - 32 * G; return r < 0 ? 0 : r / 64 + 1; }

static int sbr_stage2(gsi_ctx* ctx, const HhPlan& pl, const HhDev& D, const SbrDev& S, SbrStage2& R) {
    Workspace& ws = WS(ctx);
    cudaStream_t st = ctx->stream;
    const int nu = S.nu;
    int rc;
    // reflector blocks: user u, group G, task t -> block v2_off[u] / BLK + goff[u][G] + t
    std::vector<int64_t> v2_off(nu);
    std::vector<int> goff, goff_off(nu);
    int64_t nblocks = 0;
    for (int u = 0; u < nu; ++u) {
        const int n = pl.jobs[u].n, ng = (n - 2 + 31) / 32;
        v2_off[u] = nblocks * SBR_BLK_DBL;
        goff_off[u] = (int)goff.size();
        int run = 0;
        for (int G = 0; G <= ng; ++G) { goff.push_back(run); run += sbr_ntasks(n, G); }
        nblocks += run;
    }
    // sweeps, sweep-major (users are sorted by n descending: the users that still have sweep s are a prefix)
    std::vector<int2> list;
    for (int s = 0; s <= pl.jobs[0].n - 3; ++s)
        for (int u = 0; u < nu && pl.jobs[u].n - 3 >= s; ++u) list.push_back(make_int2(u, s));
    // back-transform items (job, 32-column block), biggest users first
    std::vector<int2> items;
    for (int u = 0; u < nu; ++u)
        for (int cb = 0; cb * 32 < pl.jobs[u].n; ++cb) items.push_back(make_int2(u, cb));
    MetaBuilder mb;
    const size_t o_v2 = mb.add(v2_off), o_goff = mb.add(goff), o_go = mb.add(goff_off), o_list = mb.add(list), o_items = mb.add(items);
    const size_t o_q = mb.reserve(64);
    if ((rc = ws.sbr_list.ensure(ctx, mb.host.size())) != GSI_OK) return rc;
    if ((rc = ws.sbr_v2.ensure(ctx, (size_t)std::max<int64_t>(nblocks, 1) * SBR_BLK_DBL * 8)) != GSI_OK) return rc;
    if ((rc = ws.sbr_prog.ensure(ctx, (size_t)S.rtot * 4)) != GSI_OK) return rc;
    GSI_CUDA(ctx, cudaMemcpyAsync(ws.sbr_list.p, mb.host.data(), mb.host.size(), cudaMemcpyHostToDevice, st));
    GSI_CUDA(ctx, cudaStreamSynchronize(st));                     // mb.host is a local
    char* base = ws.sbr_list.as<char>();
    R.V2 = ws.sbr_v2.as<double>(); R.nblocks = nblocks;
    R.v2_off = (int64_t*)(base + o_v2); R.goff = (int*)(base + o_goff); R.goff_off = (int*)(base + o_go);
    R.bt_items = (int2*)(base + o_items); R.n_bt_items = (int)items.size(); R.queue = (int*)(base + o_q);
    GSI_CUDA(ctx, cudaMemsetAsync(R.V2, 0, (size_t)nblocks * SBR_BLK_DBL * 8, st));
    GSI_CUDA(ctx, cudaMemsetAsync(ws.sbr_prog.p, 0, (size_t)S.rtot * 4, st));
    ChaseParams C;
    memset(&C, 0, sizeof C);
    C.jobs = D.jobs; C.list = (const int2*)(base + o_list); C.nitems = (int)list.size(); C.queue = R.queue;
    C.AB = S.AB; C.prog = ws.sbr_prog.as<int>(); C.V2 = R.V2; C.v2_off = R.v2_off; C.goff = R.goff; C.goff_off = R.goff_off;
    GSI_CUDA(ctx, cudaFuncSetAttribute(sbr_chase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbr_chase_smem_bytes()));
    {
        GsiSpan sp(ctx, GSI_T_SBR, 3);
        HhTrace tr(ctx, "sbr chase + t2");
        sbr_chase_kernel<<<2 * ctx->sm_count, 256, sbr_chase_smem_bytes(), st>>>(C);
        SbrParams P;
        memset(&P, 0, sizeof P);
        P.jobs = D.jobs; P.users = S.users;
        sbr_de_kernel<<<dim3((pl.jobs[0].n + 255) / 256, nu), 256, 0, st>>>(P, S.AB, D.d, D.e);
        if (nblocks > 0) sbr_t2_kernel<<<(unsigned)nblocks, 256, 0, st>>>(R.V2, nblocks);
        sp.end();
        GSI_CUDA(ctx, cudaGetLastError());
    }
    return GSI_OK;
}

// Z <- Q2 Z for the SBR users (after the last D&C merge, before bt_apply)
static int sbr_bt2(gsi_ctx* ctx, const HhDev& D, const SbrStage2& R) {
    Bt2Params B;
    memset(&B, 0, sizeof B);
    B.jobs = D.jobs; B.items = R.bt_items; B.nitems = R.n_bt_items; B.queue = R.queue + 1;
    B.V2 = R.V2; B.v2_off = R.v2_off; B.goff = R.goff; B.goff_off = R.goff_off; B.Qa = D.Qa; B.Qb = D.Qb; B.kuser = D.kuser;
    GSI_CUDA(ctx, cudaFuncSetAttribute(sbr_bt2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbr_bt2_smem_bytes()));
    GsiSpan sp(ctx, GSI_T_BT2, 1);
    HhTrace tr(ctx, "sbr bt2");
    sbr_bt2_kernel<<<2 * ctx->sm_count, 256, sbr_bt2_smem_bytes(), ctx->stream>>>(B);
    sp.end();
    GSI_CUDA(ctx, cudaGetLastError());
    return GSI_OK;
}
