// Host side of gsi_local_calc_host: the per-movie variant (local_calc.cpp:262-526).
// Included by gsi.cu after hh_host.cuh (uses Job, HhPlan, HhDev, hh_build_plan, hh_alloc, hh_solve).
//
// Per chunk of movies (biggest local graphs first, sized by the workspace limit):
//   lc_laplacian -> L per movie            lc_fill + hh_solve -> eigenpairs of sym(lower(L)) up to max row norm + 0.01
//   lc_gram      -> P = L L^T per movie
// and per chunk of (movie, user) pairs of those movies (most unrated nodes first):
//   fast path   lc_lanczos: smallest Ritz value of P[unrated, unrated] through the movie's P        -> w_lim
//   exact path  lc_fill (gather P[unrated, unrated]) + hh_trd + lc_tmin (pairs the fast path gives up on, GSI_LC_EXACT=1,
//               local graphs too big for the fast path's shared memory)                              -> w_lim
//   lc_predict -> err / pred
#pragma once

struct LcArena {              // device allocations that live until the end of the call / of a chunk
    std::vector<void*> ptrs;
    ~LcArena() { release(); }
    void release() { for (void* p : ptrs) cudaFree(p); ptrs.clear(); }
    template <class T> int alloc(gsi_ctx* ctx, T** out, size_t count) {
        void* p = nullptr;
        const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) { cudaGetLastError(); return gsi_fail(ctx, GSI_ERR_NOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); }
        ptrs.push_back(p);
        *out = (T*)p;
        return GSI_OK;
    }
    template <class T> int upload(gsi_ctx* ctx, T** out, const std::vector<T>& v) {
        int rc = alloc(ctx, out, v.size());
        if (rc != GSI_OK) return rc;
        if (!v.empty()) GSI_CUDA(ctx, cudaMemcpyAsync(*out, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        return GSI_OK;
    }
};

struct LcLight { int movie; int32_t user; int64_t t; int kk; int n_unr; };     // a pair before its index lists are built

// pipeline input for `fills.size()` jobs (already sorted by n descending), full solve with the cutoff at
// (float)(sigmax + 0.01) per job, results left in D
static int lc_solve(gsi_ctx* ctx, const std::vector<Job>& jobs, const std::vector<LcFill>& fills, const unsigned int* d_sigmax,
                    LcArena& ar, HhPlan& pl, HhDev& D) {
    cudaStream_t st = ctx->stream;
    const int nj = (int)jobs.size();
    int rc;
    hh_build_plan(jobs.data(), nj, pl);
    if ((rc = hh_alloc(ctx, pl, jobs.data(), D)) != GSI_OK) return rc;
    LcFill* d_fills;
    if ((rc = ar.upload(ctx, &d_fills, fills)) != GSI_OK) return rc;
    const int NT = pl.npmax >> 6;
    lc_fill_kernel<<<dim3(nj, NT * NT), 256, 0, st>>>(D.jobs, d_fills, D.A);
    GSI_CUDA(ctx, cudaGetLastError());
    GSI_CUDA(ctx, cudaMemcpyAsync(D.sigmax, d_sigmax, (size_t)nj * 4, cudaMemcpyDeviceToDevice, st));
    return hh_solve(ctx, pl, D, 0);
}

extern "C" int gsi_local_calc_host(gsi_ctx* ctx, int64_t nu, const int64_t* offsets, const int32_t* items, const double* ratings,
                                   const uint8_t* pair_mask, float* err, int32_t* kk, double* pred, int32_t* status, int32_t* cols,
                                   double* w_lim) {
    if (!ctx) return gsi_fail(nullptr, GSI_ERR_INVALID, "null context");
    if (!ctx->d_w) return gsi_fail(ctx, GSI_ERR_STATE, "gsi_local_calc: no weight table set (call gsi_set_weights_*)");
    if (nu < 0 || !offsets || (nu > 0 && offsets[nu] > 0 && (!items || !ratings || !err || !kk || !pred || !status || !cols)))
        return gsi_fail(ctx, GSI_ERR_INVALID, "gsi_local_calc_host: null argument");
    if (offsets[0] != 0) return gsi_fail(ctx, GSI_ERR_INVALID, "offsets[0] must be 0");
    for (int64_t u = 0; u < nu; ++u)
        if (offsets[u + 1] < offsets[u]) return gsi_fail(ctx, GSI_ERR_INVALID, "offsets must be non-decreasing");
    const int64_t nnz = offsets[nu];
    for (int64_t t = 0; t < nnz; ++t)
        if (items[t] < 0) return gsi_fail(ctx, GSI_ERR_INVALID, "negative movie id");
    GSI_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int rows = ctx->w_rows;
    int rc;
    for (int64_t t = 0; t < nnz; ++t) {
        err[t] = 0.f; kk[t] = 0; pred[t] = 0.0; status[t] = GSI_PRED_SKIPPED; cols[t] = 0;
        if (w_lim) w_lim[t] = 0.0;
    }
    if (nnz == 0) return GSI_OK;

    // ---- item graph: out-neighbour lists of the thresholded table (graph_loader :102-117, neigh_program :166-200)
    LcArena call;
    int32_t* d_cnt; int64_t* d_noff; int32_t* d_nbr;
    if ((rc = call.alloc(ctx, &d_cnt, (size_t)rows)) != GSI_OK) return rc;
    lc_nbr_count_kernel<<<rows, 128, 0, st>>>(ctx->d_w, rows, d_cnt);
    GSI_CUDA(ctx, cudaGetLastError());
    std::vector<int32_t> h_cnt(rows);
    GSI_CUDA(ctx, cudaMemcpyAsync(h_cnt.data(), d_cnt, (size_t)rows * 4, cudaMemcpyDeviceToHost, st));
    GSI_CUDA(ctx, cudaStreamSynchronize(st));
    std::vector<int64_t> noff(rows + 1, 0);
    for (int m = 0; m < rows; ++m) noff[m + 1] = noff[m] + h_cnt[m];
    if ((rc = call.upload(ctx, &d_noff, noff)) != GSI_OK) return rc;
    if ((rc = call.alloc(ctx, &d_nbr, (size_t)noff[rows])) != GSI_OK) return rc;
    lc_nbr_fill_kernel<<<(rows + 3) / 4, 128, 0, st>>>(ctx->d_w, rows, d_noff, d_nbr);
    GSI_CUDA(ctx, cudaGetLastError());
    std::vector<int32_t> nbr((size_t)noff[rows]);
    if (!nbr.empty()) GSI_CUDA(ctx, cudaMemcpyAsync(nbr.data(), d_nbr, nbr.size() * 4, cudaMemcpyDeviceToHost, st));
    GSI_CUDA(ctx, cudaStreamSynchronize(st));

    // ---- test ratings by movie (graph_test_loader :119-141): pairs (user, position in the caller's arrays)
    std::vector<int64_t> mcount(rows + 1, 0);
    for (int64_t t = 0; t < nnz; ++t)
        if (items[t] < rows && (!pair_mask || pair_mask[t])) ++mcount[items[t] + 1];
    for (int m = 0; m < rows; ++m) mcount[m + 1] += mcount[m];
    std::vector<int32_t> by_user((size_t)mcount[rows]);
    std::vector<int64_t> by_pos((size_t)mcount[rows]);
    {
        std::vector<int64_t> cur(mcount.begin(), mcount.end() - 1);
        for (int64_t u = 0; u < nu; ++u)
            for (int64_t t = offsets[u]; t < offsets[u + 1]; ++t)
                if (items[t] < rows && (!pair_mask || pair_mask[t])) { by_user[cur[items[t]]] = (int32_t)u; by_pos[cur[items[t]]++] = t; }
    }
    // active movies: at least one requested pair and a local graph of >= 3 nodes (:271-272), biggest first
    std::vector<int> active;
    for (int m = 0; m < rows; ++m)
        if (mcount[m + 1] > mcount[m] && h_cnt[m] + 1 >= 3) active.push_back(m);
    // local graphs above GSI_HH_MAX_N nodes take the two-stage tridiagonalisation like the big users of the precompute path
    for (int m : active)
        if (h_cnt[m] + 1 > 46000)
            return gsi_fail(ctx, GSI_ERR_INVALID, "movie %d: local graph of %d nodes exceeds the eigensolver's limit %d", m, h_cnt[m] + 1, 46000);
    std::stable_sort(active.begin(), active.end(), [&](int a, int b) { return h_cnt[a] > h_cnt[b]; });
    const auto npad = [](int n) { const int nn = std::max(n, LC_PAD_N); return (int64_t)hh_np(nn); };
    const int64_t budget = std::max<int64_t>(ctx->ws_limit / 64, (int64_t)GSI_HH_MAX_N * GSI_HH_MAX_N);   // doubles of sum(np^2) per chunk

    std::vector<int32_t> pos(rows, -1);                   // local index of a movie in the current local graph
    size_t mb = 0;
    while (mb < active.size()) {
        size_t me = mb;
        int64_t used = 0;
        while (me < active.size() && me - mb < 4096) {
            const int64_t c = npad(h_cnt[active[me]] + 1) * npad(h_cnt[active[me]] + 1);
            if (me > mb && used + c > budget) break;
            used += c; ++me;
        }
        const int nm = (int)(me - mb);
        LcArena ch;
        // ---- local graphs of the chunk
        std::vector<LcMovie> movies(nm);
        std::vector<int32_t> nodes;
        int64_t ltot = 0;
        int nmax = 0;
        for (int j = 0; j < nm; ++j) {
            const int m = active[mb + j], n = h_cnt[m] + 1;
            movies[j] = LcMovie{n, 0, (int64_t)nodes.size(), ltot, 0};
            nodes.push_back(m);
            nodes.insert(nodes.end(), nbr.begin() + noff[m], nbr.begin() + noff[m + 1]);
            ltot += (int64_t)n * n;
            nmax = std::max(nmax, n);
        }
        LcMovie* d_movies; int32_t* d_nodes; double *d_deg, *d_scale, *d_L, *d_P, *d_lam, *d_vec; unsigned int* d_sig;
        if ((rc = ch.upload(ctx, &d_movies, movies)) != GSI_OK) return rc;
        if ((rc = ch.upload(ctx, &d_nodes, nodes)) != GSI_OK) return rc;
        if ((rc = ch.alloc(ctx, &d_deg, nodes.size())) != GSI_OK) return rc;
        if ((rc = ch.alloc(ctx, &d_scale, nodes.size())) != GSI_OK) return rc;
        if ((rc = ch.alloc(ctx, &d_lam, nodes.size())) != GSI_OK) return rc;
        if ((rc = ch.alloc(ctx, &d_L, (size_t)ltot)) != GSI_OK) return rc;
        if ((rc = ch.alloc(ctx, &d_P, (size_t)ltot)) != GSI_OK) return rc;
        if ((rc = ch.alloc(ctx, &d_sig, (size_t)nm)) != GSI_OK) return rc;
        lc_laplacian_kernel<<<nm, 256, 0, st>>>(d_movies, d_nodes, ctx->d_w, rows, d_deg, d_scale, d_L, d_sig);
        GSI_CUDA(ctx, cudaGetLastError());
        // ---- eigenpairs of sym(lower(L)) per movie (:378)
        {
            std::vector<Job> jobs(nm);
            std::vector<LcFill> fills(nm);
            for (int j = 0; j < nm; ++j) {
                jobs[j] = Job{j, std::max(movies[j].n, LC_PAD_N), 0};
                fills[j] = LcFill{d_L + movies[j].l_off, nullptr, movies[j].n, movies[j].n};
            }
            HhPlan pl; HhDev D;
            if ((rc = lc_solve(ctx, jobs, fills, d_sig, ch, pl, D)) != GSI_OK) return rc;
            std::vector<int32_t> h_k(nm);
            GSI_CUDA(ctx, cudaMemcpyAsync(h_k.data(), D.kuser, (size_t)nm * 4, cudaMemcpyDeviceToHost, st));
            GSI_CUDA(ctx, cudaStreamSynchronize(st));
            int64_t vtot = 0;
            for (int j = 0; j < nm; ++j) {
                if (h_k[j] < 2 || h_k[j] > movies[j].n)
                    return gsi_fail(ctx, GSI_ERR_CUDA, "internal: movie %d kept %d of %d eigenpairs", active[mb + j], h_k[j], movies[j].n);
                movies[j].k = h_k[j]; movies[j].vec_off = vtot; vtot += (int64_t)movies[j].n * h_k[j];
            }
            GSI_CUDA(ctx, cudaMemcpyAsync(d_movies, movies.data(), (size_t)nm * sizeof(LcMovie), cudaMemcpyHostToDevice, st));
            if ((rc = ch.alloc(ctx, &d_vec, (size_t)vtot)) != GSI_OK) return rc;
            lc_take_kernel<<<nm, 256, 0, st>>>(D.jobs, d_movies, D.Qa, D.Qb, D.lamA, D.lamB, d_lam, d_vec);
            GSI_CUDA(ctx, cudaGetLastError());
        }
        {
            const int T = (nmax + LC_GT - 1) / LC_GT;
            for (int z0 = 0; z0 < nm; z0 += 32768) {
                const int zc = std::min(32768, nm - z0);
                lc_gram_kernel<<<dim3(T, T, zc), 256, 0, st>>>(d_movies + z0, d_L, d_P);
            }
            GSI_CUDA(ctx, cudaGetLastError());
        }
        // ---- pairs of the chunk: known / unrated node sets per (movie, user) (:393-413)
        std::vector<LcLight> light;
        for (int j = 0; j < nm; ++j) {
            const int m = active[mb + j], n = movies[j].n;
            for (int i = 0; i < n; ++i) pos[nodes[movies[j].node_off + i]] = i;
            for (int64_t q = mcount[m]; q < mcount[m + 1]; ++q) {
                const int32_t u = by_user[q];
                int known = 0;
                for (int64_t t = offsets[u]; t < offsets[u + 1]; ++t)
                    if (items[t] < rows && pos[items[t]] > 0 && ratings[t] != 0.0) ++known;
                const int64_t t0 = by_pos[q];
                kk[t0] = known;
                if (known == 0) {                         // nothing known: 0/0 (:487); no cutoff is needed
                    status[t0] = GSI_PRED_EMPTY;
                    pred[t0] = nan(""); err[t0] = nanf(""); cols[t0] = 0;
                    if (w_lim) w_lim[t0] = nan("");
                    continue;
                }
                light.push_back(LcLight{j, u, t0, known, n - known});
            }
            for (int i = 0; i < n; ++i) pos[nodes[movies[j].node_off + i]] = -1;
        }
        // one batch of pairs: index lists, cutoff (Lanczos fast path or exact tridiagonalisation), prediction, results
        const size_t lz_smem = (size_t)nmax * 25 + 64;
        auto run_pairs = [&](const LcLight* lp, int np_, bool exact, std::vector<LcLight>* retry) -> int {
            LcArena pc;
            std::vector<LcPair> pairs(np_);
            std::vector<int32_t> kidx, uidx;
            std::vector<double> krat;
            std::vector<int64_t> u_off(np_);
            size_t capA = 1, capM = 1;
            for (int p = 0; p < np_; ++p) {
                const LcLight& Lp = lp[p];
                const LcMovie& M = movies[Lp.movie];
                for (int i = 0; i < M.n; ++i) pos[nodes[M.node_off + i]] = i;
                pairs[p] = LcPair{Lp.movie, Lp.kk, (int64_t)kidx.size(), ratings[Lp.t]};
                std::vector<std::pair<int32_t, double>> known;
                for (int64_t t = offsets[Lp.user]; t < offsets[Lp.user + 1]; ++t)
                    if (items[t] < rows && pos[items[t]] > 0 && ratings[t] != 0.0) known.emplace_back(pos[items[t]], ratings[t]);
                std::sort(known.begin(), known.end());
                u_off[p] = (int64_t)uidx.size();
                size_t q = 0;
                for (int i = 0; i < M.n; ++i) {
                    if (q < known.size() && known[q].first == i) { kidx.push_back(i); krat.push_back(known[q].second); ++q; }
                    else if (exact) uidx.push_back(i);
                }
                for (int i = 0; i < M.n; ++i) pos[nodes[M.node_off + i]] = -1;
                const size_t lm = (size_t)std::min(Lp.kk, M.k);
                capA = std::max(capA, (size_t)Lp.kk * lm);
                capM = std::max(capM, lm * lm + 2 * lm);
            }
            LcPair* d_pairs; int32_t *d_kidx, *d_uidx; double *d_krat, *d_wl, *d_pred, *d_scr; float* d_err; int32_t *d_status, *d_cols;
            if ((rc = pc.upload(ctx, &d_pairs, pairs)) != GSI_OK) return rc;
            if ((rc = pc.upload(ctx, &d_kidx, kidx)) != GSI_OK) return rc;
            if ((rc = pc.upload(ctx, &d_uidx, uidx)) != GSI_OK) return rc;
            if ((rc = pc.upload(ctx, &d_krat, krat)) != GSI_OK) return rc;
            if ((rc = pc.alloc(ctx, &d_wl, (size_t)np_)) != GSI_OK) return rc;
            if ((rc = pc.alloc(ctx, &d_pred, (size_t)np_)) != GSI_OK) return rc;
            if ((rc = pc.alloc(ctx, &d_err, (size_t)np_)) != GSI_OK) return rc;
            if ((rc = pc.alloc(ctx, &d_status, (size_t)np_)) != GSI_OK) return rc;
            if ((rc = pc.alloc(ctx, &d_cols, (size_t)np_)) != GSI_OK) return rc;
            int32_t* d_conv = nullptr;
            if (!exact) {   // fast path: Lanczos on P[unrated, unrated] through the movie's P, no per-pair matrix (kern_lc.cuh)
                if ((rc = pc.alloc(ctx, &d_conv, (size_t)np_)) != GSI_OK) return rc;
                GSI_CUDA(ctx, cudaFuncSetAttribute(lc_lanczos_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lz_smem));
                const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / lz_smem));
                const char* g = getenv("GSI_LC_GUARD");                      // 0: single run (the r01 behaviour); default: two start vectors
                const int ntrials = (g && atoi(g) == 0) ? 1 : 2;
                lc_lanczos_kernel<<<std::min(np_, per_sm * ctx->sm_count), 256, lz_smem, st>>>(d_pairs, np_, d_movies, d_P, d_kidx, nmax, ntrials, d_wl, d_conv);
                GSI_CUDA(ctx, cudaGetLastError());
            }
            if (exact) {   // exact cutoff: smallest eigenvalue of L_h L_h^T = P[unrated, unrated] (:417-436).  Only that one value is
                // needed, so a pair job is tridiagonalised (the input matrix is its only np^2 buffer) and the smallest
                // eigenvalue of T is bracketed on the Sturm count; divide & conquer and back-transform are skipped.
                std::vector<Job> jobs(np_);
                std::vector<LcFill> fills(np_);
                for (int p = 0; p < np_; ++p) {
                    const LcMovie& M = movies[pairs[p].movie];
                    jobs[p] = Job{p, std::max(lp[p].n_unr, LC_PAD_N), 0};
                    fills[p] = LcFill{d_P + M.l_off, d_uidx + u_off[p], M.n, lp[p].n_unr};
                }
                HhPlan pl;
                hh_build_plan(jobs.data(), np_, pl);
                HhDev D;
                memset(&D, 0, sizeof D);
                HJob* d_jobs; LcFill* d_fills; double *d_A, *d_tv; int* d_ctl;
                if ((rc = pc.upload(ctx, &d_jobs, pl.jobs)) != GSI_OK) return rc;
                if ((rc = pc.upload(ctx, &d_fills, fills)) != GSI_OK) return rc;
                if ((rc = pc.alloc(ctx, &d_A, (size_t)pl.mtot)) != GSI_OK) return rc;
                if ((rc = pc.alloc(ctx, &d_tv, (size_t)pl.rtot * 3)) != GSI_OK) return rc;
                if ((rc = pc.alloc(ctx, &d_ctl, (size_t)HH_CTL_INTS)) != GSI_OK) return rc;
                GSI_CUDA(ctx, cudaMemsetAsync(d_A, 0, (size_t)pl.mtot * 8, st));
                D.jobs = d_jobs; D.A = d_A; D.d = d_tv; D.e = d_tv + pl.rtot; D.tau = d_tv + 2 * pl.rtot; D.ctl = d_ctl;
                const int NT = pl.npmax >> 6;
                lc_fill_kernel<<<dim3(np_, NT * NT), 256, 0, st>>>(D.jobs, d_fills, D.A);
                GSI_CUDA(ctx, cudaGetLastError());
                { HhSbrState sb; if ((rc = hh_trd(ctx, pl, D, 0, sb)) != GSI_OK) return rc; }
                lc_tmin_kernel<<<(np_ + 3) / 4, 128, 0, st>>>(D.jobs, np_, D.d, D.e, d_wl);
                GSI_CUDA(ctx, cudaGetLastError());
            }
            const size_t per_cta = capA + capM;
            int grid = std::min(np_, 2 * ctx->sm_count);
            while (grid > 1 && (size_t)grid * per_cta * 8 > ((size_t)2 << 30)) grid = (grid + 1) / 2;
            if ((rc = pc.alloc(ctx, &d_scr, (size_t)grid * per_cta)) != GSI_OK) return rc;
            lc_predict_kernel<<<grid, 256, 0, st>>>(d_pairs, np_, d_movies, d_lam, d_vec, d_kidx, d_krat, d_wl, d_scr, per_cta,
                                                    d_err, d_pred, d_status, d_cols);
            GSI_CUDA(ctx, cudaGetLastError());
            std::vector<float> h_err(np_);
            std::vector<double> h_pred(np_), h_wl(np_);
            std::vector<int32_t> h_status(np_), h_cols(np_);
            GSI_CUDA(ctx, cudaMemcpyAsync(h_err.data(), d_err, (size_t)np_ * 4, cudaMemcpyDeviceToHost, st));
            GSI_CUDA(ctx, cudaMemcpyAsync(h_pred.data(), d_pred, (size_t)np_ * 8, cudaMemcpyDeviceToHost, st));
            GSI_CUDA(ctx, cudaMemcpyAsync(h_wl.data(), d_wl, (size_t)np_ * 8, cudaMemcpyDeviceToHost, st));
            GSI_CUDA(ctx, cudaMemcpyAsync(h_status.data(), d_status, (size_t)np_ * 4, cudaMemcpyDeviceToHost, st));
            GSI_CUDA(ctx, cudaMemcpyAsync(h_cols.data(), d_cols, (size_t)np_ * 4, cudaMemcpyDeviceToHost, st));
            std::vector<int32_t> h_conv(np_, 1);
            if (d_conv) GSI_CUDA(ctx, cudaMemcpyAsync(h_conv.data(), d_conv, (size_t)np_ * 4, cudaMemcpyDeviceToHost, st));
            GSI_CUDA(ctx, cudaStreamSynchronize(st));
            for (int p = 0; p < np_; ++p) {
                if (!h_conv[p]) { retry->push_back(lp[p]); continue; }       // not converged: the exact path decides
                const int64_t t0 = lp[p].t;
                err[t0] = h_err[p]; pred[t0] = h_pred[p]; status[t0] = h_status[p]; cols[t0] = h_cols[p];
                if (w_lim) w_lim[t0] = h_wl[p];
            }
            return GSI_OK;
        };
        const char* fe = getenv("GSI_LC_EXACT");               // GSI_LC_EXACT=1: every pair takes the exact path (tests, comparison)
        const bool lanczos = !(fe && atoi(fe) != 0) && lz_smem <= 200 * 1024;
        std::vector<LcLight> retry;
        if (lanczos) {                                          // movie order: the pairs of a movie share its P in L2
            for (size_t pb = 0; pb < light.size(); pb += 32768)
                if ((rc = run_pairs(light.data() + pb, (int)std::min<size_t>(32768, light.size() - pb), false, &retry)) != GSI_OK) return rc;
        } else {
            retry.swap(light);
        }
        std::stable_sort(retry.begin(), retry.end(), [](const LcLight& a, const LcLight& b) { return a.n_unr > b.n_unr; });
        size_t pb = 0;
        while (pb < retry.size()) {
            size_t pe = pb;
            int64_t pused = 0;
            while (pe < retry.size() && pe - pb < 16384) {
                const int64_t c = npad(retry[pe].n_unr) * npad(retry[pe].n_unr);
                if (pe > pb && pused + c > 4 * budget) break;
                pused += c; ++pe;
            }
            if ((rc = run_pairs(retry.data() + pb, (int)(pe - pb), true, nullptr)) != GSI_OK) return rc;
            pb = pe;
        }
        GSI_CUDA(ctx, cudaStreamSynchronize(st));
        mb = me;
    }
    return GSI_OK;
}
