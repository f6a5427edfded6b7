// libgsi.so -- C ABI (include/gsi.h) over the sm_100a kernels.  Host side: context, planner
// (bucketing users by n, chunking by workspace), launch sequencing, staging.  No CPU fallback:
// every entry point needs a CUDA device.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <numeric>

#include "gsi_internal.cuh"
#include "kern_bj.cuh"
#include "kern_eig_cta.cuh"
#include "kern_out.cuh"
#include "kern_predict.cuh"
#include "kern_knn.cuh"
#include "kern_trd.cuh"
#include "kern_dc.cuh"
#include "kern_bt.cuh"
#include "kern_sbr.cuh"
#include "kern_cheby.cuh"
#include "kern_lc.cuh"
#include "tc_gemm.cuh"
#include "kern_bt_tc.cuh"

static thread_local std::string g_tls_err;

int gsi_fail(gsi_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    g_tls_err = buf;
    return code;
}

// ---- timing spans --------------------------------------------------------------------------
static cudaEvent_t take_event(gsi_ctx* ctx) {
    if (!ctx->ev_pool.empty()) { cudaEvent_t e = ctx->ev_pool.back(); ctx->ev_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
GsiSpan::GsiSpan(gsi_ctx* c, int cls_, int64_t n_launches, int64_t n_timed) : ctx(c), cls(cls_), launches(n_launches) {
    ctx->t_launch[cls] += n_launches;
    const int64_t timed = n_timed < 0 ? n_launches : n_timed;
    if (ctx->timing && timed > 0) {
        ctx->t_samples[cls] += timed;
        a = take_event(ctx); b = take_event(ctx); cudaEventRecord(a, ctx->stream);
    }
}
void GsiSpan::end() {
    if (a) { cudaEventRecord(b, ctx->stream); ctx->spans.push_back({cls, a, b}); a = nullptr; }
}
void gsi_count_launch(gsi_ctx* ctx, int cls, int64_t n) { ctx->t_launch[cls] += n; }

static void drain_spans(gsi_ctx* ctx) {
    for (auto& s : ctx->spans) {
        float ms = 0.f;
        if (cudaEventSynchronize(s.b) == cudaSuccess && cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess)
            ctx->t_ms[s.cls] += ms;
        ctx->ev_pool.push_back(s.a);
        ctx->ev_pool.push_back(s.b);
    }
    ctx->spans.clear();
}

// ---- growable device / pinned buffers ---------------------------------------------------------
struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    int ensure(gsi_ctx* ctx, size_t bytes) {
        if (bytes <= cap) return GSI_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
        if (e != cudaSuccess) { cudaGetLastError(); return gsi_fail(ctx, GSI_ERR_NOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); }
        cap = want;
        return GSI_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() { return (T*)p; }
};
struct PinBuf {
    void* p = nullptr; size_t cap = 0;
    int ensure(gsi_ctx* ctx, size_t bytes) {
        if (bytes <= cap) return GSI_OK;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess) { cudaGetLastError(); return gsi_fail(ctx, GSI_ERR_NOMEM, "cudaMallocHost(%zu): %s", want, cudaGetErrorString(e)); }
        cap = want;
        return GSI_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() { return (T*)p; }
};

struct Workspace {
    DevBuf k_off, k_items, k_rat, k_cnt, k_num, k_S, k_ecnt, k_eoff, k_ea, k_eb, k_ew, k_co, k_err, k_mcnt, k_has;
    DevBuf k_coff, k_cur, k_cuser, k_crat, k_useg;
    DevBuf tc_meta, tc_tasks, tc_planes, tc_expo;
    DevBuf btt_gt, btt_x, btt_planes, btt_expo, btt_meta;   // tensor-core back-transform (hh_bt_tc)      // tensor-core merges of the divide & conquer (hh_tc_plan)     // item-major transpose of the knn CSR (knn_row_kernel)
    DevBuf c_off, c_col, c_w, c_wn, c_vec;          // Chebyshev filter: CSR, normalised weights, 5 vertex vectors
    int64_t knn_edges = 0;
    DevBuf pred_meta, pred_work, p_off, p_items, p_wlim, p_rat, p_k, p_lamoff, p_vecoff, p_lam, p_vec, p_err, p_kk, p_pred, p_status, p_cols, p_gsum, p_hsum, p_sumoff, p_exact, p_lim, p_mask;
    DevBuf meta, vec_pad, lam_pad, G, rows, cols, perm, hpart, q, totals, items, sig, outk, outlam, outvec, stage_vec, stage_lam, probe;
    DevBuf hhA, hhQa, hhQb, hhS, hhvec, hhivec, trd_acol, trd_ypart, trd_part, trd_panels;
    DevBuf sbr_panels, sbr_small, sbr_band, sbr_v2, sbr_prog, sbr_list;
    PinBuf h_meta, h_stage_vec, h_stage_lam, h_small, h_k, h_lamoff, h_vecoff, h_sig;
};

struct gsi_ctx_full : gsi_ctx { Workspace ws; };
static Workspace& WS(gsi_ctx* c) { return static_cast<gsi_ctx_full*>(c)->ws; }

// ---- context ----------------------------------------------------------------------------------
template <int EPT> static cudaError_t set_smem_attr() {
    return cudaFuncSetAttribute(eig_cta_kernel<EPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)eig_cta_smem_bytes(32 * EPT));
}

extern "C" int gsi_create(gsi_ctx** out, int device, void* stream) {
    if (!out) return gsi_fail(nullptr, GSI_ERR_INVALID, "gsi_create: out is null");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return gsi_fail(nullptr, GSI_ERR_CUDA, "gsi_create: no CUDA device (%s); libgsi has no CPU fallback",
                        cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return gsi_fail(nullptr, GSI_ERR_INVALID, "gsi_create: device %d of %d", device, ndev);
    gsi_ctx_full* ctx = new gsi_ctx_full();
    // a failure below must not leak the half-built context (and its message has to outlive it: thread-local, gsi_last_error(NULL))
#define CREATE_CUDA(call)                                                                                              \
    do {                                                                                                               \
        cudaError_t e_ = (call);                                                                                       \
        if (e_ != cudaSuccess) {                                                                                       \
            const int rc_ = gsi_fail(nullptr, GSI_ERR_CUDA, "gsi_create: %s: %s", #call, cudaGetErrorString(e_));     \
            if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);                                        \
            delete ctx;                                                                                                \
            return rc_;                                                                                                \
        }                                                                                                              \
    } while (0)
    ctx->device = device;
    CREATE_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    CREATE_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        int rc = gsi_fail(nullptr, GSI_ERR_CUDA, "gsi_create: device %d is sm_%d%d; libgsi is built for sm_100a only", device, prop.major, prop.minor);
        delete ctx;
        return rc;
    }
    ctx->sm_count = prop.multiProcessorCount;
    if (stream) ctx->stream = (cudaStream_t)stream;
    else { CREATE_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->own_stream = true; }
    CREATE_CUDA(set_smem_attr<1>());
    CREATE_CUDA(set_smem_attr<2>());
    CREATE_CUDA(set_smem_attr<3>());
    CREATE_CUDA(set_smem_attr<4>());
    CREATE_CUDA(set_smem_attr<5>());
    const char* bm = getenv("GSI_BJ_M");
    ctx->bj_m = (bm && atoi(bm) == 32) ? 32 : 64;
    const char* lp = getenv("GSI_LARGE");          // "bj": one-sided block Jacobi (kept for comparison); default Householder + D&C
    ctx->large_bj = lp && strcmp(lp, "bj") == 0;
    if (const char* sm = getenv("GSI_SMALL_MAX")) ctx->small_max = std::min(GSI_S_MAX_N, std::max(32, atoi(sm)));
    const char* tr = getenv("GSI_TRACE");
    ctx->trace = tr && atoi(tr) != 0;
    CREATE_CUDA(cudaFuncSetAttribute(bj_inner_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 64 * 65 * 8));
#undef CREATE_CUDA
    *out = ctx;
    return GSI_OK;
}

extern "C" int gsi_destroy(gsi_ctx* c) {
    if (!c) return GSI_OK;
    gsi_ctx_full* ctx = static_cast<gsi_ctx_full*>(c);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); ctx->copy_stream = nullptr; }
    drain_spans(ctx);
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    Workspace& w = ctx->ws;
    DevBuf* d[] = {&w.meta, &w.vec_pad, &w.lam_pad, &w.G, &w.rows, &w.cols, &w.perm, &w.hpart, &w.q, &w.totals, &w.items,
                   &w.sig, &w.outk, &w.outlam, &w.outvec, &w.stage_vec, &w.stage_lam, &w.probe};
    for (auto b : d) b->release();
    DevBuf* d2[] = {&w.pred_meta, &w.pred_work, &w.p_off, &w.p_items, &w.p_wlim, &w.p_rat, &w.p_k, &w.p_lamoff, &w.p_vecoff,
                    &w.p_lam, &w.p_vec, &w.p_err, &w.p_kk, &w.p_pred, &w.p_status, &w.p_cols, &w.p_gsum, &w.p_hsum, &w.p_sumoff,
                    &w.p_exact, &w.p_lim, &w.p_mask};
    for (auto b : d2) b->release();
    DevBuf* d3[] = {&w.k_off, &w.k_items, &w.k_rat, &w.k_cnt, &w.k_num, &w.k_S, &w.k_ecnt, &w.k_eoff, &w.k_ea, &w.k_eb, &w.k_ew,
                    &w.k_co, &w.k_err, &w.k_mcnt, &w.k_has, &w.k_coff, &w.k_cur, &w.k_cuser, &w.k_crat, &w.k_useg, &w.tc_meta, &w.tc_tasks, &w.tc_planes, &w.tc_expo, &w.btt_gt, &w.btt_x, &w.btt_planes, &w.btt_expo, &w.btt_meta, &w.c_off, &w.c_col, &w.c_w, &w.c_wn, &w.c_vec};
    for (auto b : d3) b->release();
    DevBuf* d4[] = {&w.hhA, &w.hhQa, &w.hhQb, &w.hhS, &w.hhvec, &w.hhivec, &w.trd_acol, &w.trd_ypart, &w.trd_part, &w.trd_panels,
                    &w.sbr_panels, &w.sbr_small, &w.sbr_band, &w.sbr_v2, &w.sbr_prog, &w.sbr_list};
    for (auto b : d4) b->release();
    PinBuf* p[] = {&w.h_meta, &w.h_stage_vec, &w.h_stage_lam, &w.h_small, &w.h_k, &w.h_lamoff, &w.h_vecoff, &w.h_sig};
    for (auto b : p) b->release();
    if (ctx->own_w && ctx->d_w) cudaFree(ctx->d_w);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return GSI_OK;
}

extern "C" const char* gsi_last_error(const gsi_ctx* ctx) { return ctx ? ctx->err.c_str() : g_tls_err.c_str(); }
extern "C" const char* gsi_version(void) { return "gsi 0.1 sm_100a fp64"; }
extern "C" int gsi_set_workspace_limit(gsi_ctx* ctx, int64_t bytes) {
    if (!ctx || bytes < ((int64_t)64 << 20)) return gsi_fail(ctx, GSI_ERR_INVALID, "workspace limit must be >= 64 MiB");
    ctx->ws_limit = bytes;
    return GSI_OK;
}
extern "C" int gsi_small_max(const gsi_ctx* ctx) { return ctx ? ctx->small_max : GSI_SMALL_DEFAULT; }
extern "C" int gsi_sync(gsi_ctx* ctx) {
    if (!ctx) return GSI_ERR_INVALID;
    GSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GSI_OK;
}

// ---- weights ----------------------------------------------------------------------------------
static int drop_weights(gsi_ctx* ctx) {
    if (ctx->own_w && ctx->d_w) cudaFree(ctx->d_w);
    ctx->d_w = nullptr; ctx->w_rows = 0; ctx->own_w = false;
    return GSI_OK;
}
extern "C" int gsi_set_weights_host(gsi_ctx* ctx, const double* w, int rows) {
    if (!ctx || !w || rows < 1) return gsi_fail(ctx, GSI_ERR_INVALID, "gsi_set_weights_host: bad arguments");
    GSI_CUDA(ctx, cudaSetDevice(ctx->device));
    drop_weights(ctx);
    GSI_CUDA(ctx, cudaMalloc((void**)&ctx->d_w, (size_t)rows * rows * sizeof(double)));
    ctx->own_w = true; ctx->w_rows = rows;
    GSI_CUDA(ctx, cudaMemcpyAsync(ctx->d_w, w, (size_t)rows * rows * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    GSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GSI_OK;
}
extern "C" int gsi_set_weights_device(gsi_ctx* ctx, const double* d_w, int rows) {
    if (!ctx || !d_w || rows < 1) return gsi_fail(ctx, GSI_ERR_INVALID, "gsi_set_weights_device: bad arguments");
    drop_weights(ctx);
    ctx->d_w = const_cast<double*>(d_w); ctx->w_rows = rows; ctx->own_w = false;
    return GSI_OK;
}
extern "C" int gsi_set_weights_edges(gsi_ctx* ctx, const int32_t* m1, const int32_t* m2, const double* w, int64_t ne, int* rows_out) {
    if (!ctx || (ne > 0 && (!m1 || !m2 || !w))) return gsi_fail(ctx, GSI_ERR_INVALID, "gsi_set_weights_edges: bad arguments");
    GSI_CUDA(ctx, cudaSetDevice(ctx->device));
    int mx = 0;
    for (int64_t e = 0; e < ne; ++e) {
        if (m1[e] < 0 || m2[e] < 0) return gsi_fail(ctx, GSI_ERR_INVALID, "negative movie id in edge %lld", (long long)e);
        mx = std::max(mx, std::max(m1[e], m2[e]));
    }
    const int rows = mx + 1;
    drop_weights(ctx);
    GSI_CUDA(ctx, cudaMalloc((void**)&ctx->d_w, (size_t)rows * rows * sizeof(double)));
    ctx->own_w = true; ctx->w_rows = rows;
    GSI_CUDA(ctx, cudaMemsetAsync(ctx->d_w, 0, (size_t)rows * rows * sizeof(double), ctx->stream));
    if (ne > 0) {
        struct Tmp {                                                // freed on every exit path
            int32_t *da = nullptr, *db = nullptr; double* dw = nullptr;
            ~Tmp() { cudaFree(da); cudaFree(db); cudaFree(dw); }
        } t;
        GSI_CUDA(ctx, cudaMalloc((void**)&t.da, ne * 4));
        GSI_CUDA(ctx, cudaMalloc((void**)&t.db, ne * 4));
        GSI_CUDA(ctx, cudaMalloc((void**)&t.dw, ne * 8));
        GSI_CUDA(ctx, cudaMemcpyAsync(t.da, m1, ne * 4, cudaMemcpyHostToDevice, ctx->stream));
        GSI_CUDA(ctx, cudaMemcpyAsync(t.db, m2, ne * 4, cudaMemcpyHostToDevice, ctx->stream));
        GSI_CUDA(ctx, cudaMemcpyAsync(t.dw, w, ne * 8, cudaMemcpyHostToDevice, ctx->stream));
        scatter_edges_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, ctx->stream>>>(t.da, t.db, t.dw, ne, ctx->d_w, rows);
        GSI_CUDA(ctx, cudaGetLastError());
        GSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (rows_out) *rows_out = rows;
    return GSI_OK;
}
extern "C" int gsi_get_weights(gsi_ctx* ctx, double** d_w, int* rows) {
    if (!ctx || !ctx->d_w) return gsi_fail(ctx, GSI_ERR_STATE, "no weight table set");
    if (d_w) *d_w = ctx->d_w;
    if (rows) *rows = ctx->w_rows;
    return GSI_OK;
}

// ---- planner ----------------------------------------------------------------------------------
struct Job { int64_t user; int n; int64_t item_off; };
struct Chunk { bool large; int begin, end; bool bj = false; };   // [begin, end) into the sorted job list; bj: block-Jacobi path

static inline int64_t pad_slots(int n) { return (int64_t)n * std::max(n, 2); }
static inline int ld_of(int n) { return (n + 7) & ~7; }
static inline int hh_np(int n) { return (n + 63) & ~63; }
static int sbr_min_n();
// workspace of one user on the Householder path: A, Qa, Qb, S (np^2 each) + vectors; two-stage users add the reflector
// blocks of the bulge chase (~np^2), the compact band and the panel buffers
static inline int64_t hh_ws_doubles(int n) {
    const int64_t np = hh_np(n);
    const bool two_stage = sbr_min_n() > 0 && n >= sbr_min_n();
    return two_stage ? 5 * np * np + 1024 * np : 4 * np * np + 32 * np;
}
static inline int nb_of(int n, int B) { int nb = (n + B - 1) / B; return nb + (nb & 1); }

static int plan(gsi_ctx* ctx, int64_t nu, const int64_t* off, std::vector<Job>& small, std::vector<Job>& large,
                std::vector<Chunk>& chunks) {
    small.clear(); large.clear(); chunks.clear();
    for (int64_t u = 0; u < nu; ++u) {
        const int64_t n = off[u + 1] - off[u];
        if (n < 1) return gsi_fail(ctx, GSI_ERR_INVALID, "user %lld has %lld rated movies (need >= 1)", (long long)u, (long long)n);
        if (n > 46000) return gsi_fail(ctx, GSI_ERR_INVALID, "user %lld: n = %lld too large", (long long)u, (long long)n);
        (n <= ctx->small_max ? small : large).push_back({u, (int)n, off[u]});
    }
    auto desc = [](const Job& a, const Job& b) { return a.n != b.n ? a.n > b.n : a.user < b.user; };
    std::sort(small.begin(), small.end(), desc);
    std::sort(large.begin(), large.end(), desc);
    // large first (longest jobs first), then small
    const int64_t budget = ctx->ws_limit / (int64_t)sizeof(double);
    int b = 0;
    // users too large for the shared-memory vectors of the tridiagonalisation kernel (n > GSI_HH_MAX_N, e.g. the
    // heaviest Netflix users) stay on the block-Jacobi path
    int n_bj_first = (int)large.size();                                         // large[0 .. n_bj_first) -> block Jacobi
    if (!ctx->large_bj) {
        n_bj_first = 0;
        // (the two-stage reduction has no such limit: with it -- the default -- nobody takes the block-Jacobi path)
        while (sbr_min_n() == 0 && n_bj_first < (int)large.size() && large[n_bj_first].n > GSI_HH_MAX_N) ++n_bj_first;   // sorted descending
    }
    b = n_bj_first;
    while (b < (int)large.size() && !ctx->large_bj) {
        // Householder path: any mix of sizes, bounded by the workspace (4 np^2 doubles per user)
        int64_t used = 0;
        int e = b;
        while (e < (int)large.size() && e - b < 60000) {
            const int64_t need = hh_ws_doubles(large[e].n);
            if (e > b && used + need > budget) break;
            used += need; ++e;
        }
        chunks.push_back({true, b, e, false});
        b = e;
    }
    b = 0;
    while (b < n_bj_first) {
        const int nmax = large[b].n;
        const int B = ctx->bj_m / 2, MM = ctx->bj_m * ctx->bj_m;
        const int nb = nb_of(nmax, B), ncols = nb * B;
        int64_t used = 0;
        int e = b;
        while (e < n_bj_first && e - b < 60000) {
            const int n = large[e].n;
            if (e > b && (double)n < 0.7 * nmax) break;
            const int64_t need = (int64_t)ld_of(n) * ncols + pad_slots(n) + (int64_t)(nb / 2) * MM * (1 + (ld_of(nmax) + GSI_BJ_ROWS - 1) / GSI_BJ_ROWS);
            if (e > b && used + need > budget) break;
            used += need; ++e;
        }
        chunks.push_back({true, b, e, true});
        b = e;
    }
    b = 0;
    while (b < (int)small.size()) {
        int64_t used = 0;
        int e = b;
        while (e < (int)small.size() && e - b < (1 << 20)) {
            const int64_t need = pad_slots(small[e].n) + small[e].n;
            if (e > b && used + need > budget) break;
            used += need; ++e;
        }
        chunks.push_back({false, b, e, false});
        b = e;
    }
    return GSI_OK;
}

// meta upload helper: lay arrays out in one pinned block, copy once
struct MetaBuilder {
    std::vector<char> host;
    template <class T> size_t add(const std::vector<T>& v) {
        size_t o = (host.size() + 15) & ~(size_t)15;
        host.resize(o + v.size() * sizeof(T));
        if (!v.empty()) memcpy(host.data() + o, v.data(), v.size() * sizeof(T));
        return o;
    }
    size_t reserve(size_t bytes) {
        size_t o = (host.size() + 15) & ~(size_t)15;
        host.resize(o + bytes, 0);
        return o;
    }
};

struct RunOut {               // where a chunk's records go
    double* d_lam; int64_t lam_cap;
    double* d_vec; int64_t vec_cap;
    int64_t* d_totals;        // [3] device: running lam, vec totals, overflow flag
    int32_t* d_k; int64_t* d_lam_off; int64_t* d_vec_off;   // caller-order arrays [n_users]
    double* d_sig_min;
    // host path only: pinned destination of d_vec (same offsets).  A chunk runner that copies its eigenvector blocks itself,
    // group by group behind the kernels of the next group, sets *vec_copied.
    double* h_vec = nullptr; int64_t* h_bounds = nullptr; bool* vec_copied = nullptr;
};

static int upload_meta(gsi_ctx* ctx, MetaBuilder& mb, char** d_base) {
    Workspace& ws = WS(ctx);
    int rc;
    if ((rc = ws.meta.ensure(ctx, mb.host.size())) != GSI_OK) return rc;
    if ((rc = ws.h_meta.ensure(ctx, mb.host.size())) != GSI_OK) return rc;
    // h_meta may still be in flight from the previous chunk's async copy
    GSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(ws.h_meta.p, mb.host.data(), mb.host.size());
    GSI_CUDA(ctx, cudaMemcpyAsync(ws.meta.p, ws.h_meta.p, mb.host.size(), cudaMemcpyHostToDevice, ctx->stream));
    *d_base = ws.meta.as<char>();
    return GSI_OK;
}

// offsets (scan over all jobs of the chunk) and compaction of the jobs [jb, je) (default: all)
static int finish_chunk(gsi_ctx* ctx, OutJobs J, const RunOut& out, int64_t max_nk, const int32_t* h_n = nullptr,
                        bool do_scan = true, int jb = 0, int je = -1) {
    Workspace& ws = WS(ctx);
    GsiSpan sp(ctx, GSI_T_COMPACT, 2);
    if (do_scan) {
        out_scan_kernel<<<1, 1024, 0, ctx->stream>>>(J, out.d_totals, out.lam_cap, out.vec_cap, out.d_k, out.d_lam_off, out.d_vec_off);
        GSI_CUDA(ctx, cudaGetLastError());
    }
    if (je < 0) je = J.nj;
    // jobs are sorted by n descending: one launch per group of similar size (h_n), so that the slice grid matches
    int b = jb;
    while (b < je) {
        int e = je;
        int64_t nk = max_nk;
        if (h_n) {
            e = b;
            while (e < je && 2 * h_n[e] > h_n[b]) ++e;
            nk = (int64_t)h_n[b] * std::max(h_n[b], 2);
        }
        OutJobs G = J;
        G.nj = e - b; G.n = J.n + b; G.k = J.k + b; G.user = J.user + b; G.vec_pad = J.vec_pad + b; G.lam_pad = J.lam_pad + b;
        G.vec_dst = J.vec_dst + b; G.lam_dst = J.lam_dst + b;
        const unsigned slices = (unsigned)std::max<int64_t>(1, (nk + GSI_COMPACT_SLICE - 1) / GSI_COMPACT_SLICE);
        out_compact_kernel<<<dim3(G.nj, slices), 256, 0, ctx->stream>>>(G, ws.vec_pad.as<double>(), ws.lam_pad.as<double>(), out.d_vec, out.d_lam, out.lam_cap, out.vec_cap);
        GSI_CUDA(ctx, cudaGetLastError());
        b = e;
    }
    sp.end();
    return GSI_OK;
}

template <int EPT>
static cudaError_t launch_eig_cta(gsi_ctx* ctx, const SParams& P, int njobs, int nmax) {
    const int threads = EPT <= 2 ? 256 : (EPT == 3 ? 512 : 1024);
    eig_cta_kernel<EPT><<<njobs, threads, eig_cta_smem_bytes(nmax), ctx->stream>>>(P);
    return cudaGetLastError();
}

// ---- small chunk: one fused kernel per (EPT, smem) class -----------------------------------------
static int run_small_chunk(gsi_ctx* ctx, const Job* jobs, int nj, const int32_t* d_items, const RunOut& out) {
    Workspace& ws = WS(ctx);
    std::vector<int64_t> item_off(nj), vec_off(nj), lam_off(nj), user(nj);
    std::vector<int32_t> n(nj);
    int64_t vtot = 0, ltot = 0, max_nk = 0;
    for (int j = 0; j < nj; ++j) {
        item_off[j] = jobs[j].item_off; n[j] = jobs[j].n; user[j] = jobs[j].user;
        vec_off[j] = vtot; lam_off[j] = ltot;
        vtot += pad_slots(jobs[j].n); ltot += std::max(jobs[j].n, 2);
        max_nk = std::max(max_nk, pad_slots(jobs[j].n));
    }
    MetaBuilder mb;
    const size_t o_item = mb.add(item_off), o_vec = mb.add(vec_off), o_lam = mb.add(lam_off), o_user = mb.add(user), o_n = mb.add(n);
    const size_t o_k = mb.reserve(nj * 4), o_sw = mb.reserve(nj * 4), o_vd = mb.reserve(nj * 8), o_ld = mb.reserve(nj * 8);
    int rc;
    if ((rc = ws.vec_pad.ensure(ctx, vtot * 8)) != GSI_OK) return rc;
    if ((rc = ws.lam_pad.ensure(ctx, ltot * 8)) != GSI_OK) return rc;
    char* base;
    if ((rc = upload_meta(ctx, mb, &base)) != GSI_OK) return rc;
    SParams P;
    P.W = ctx->d_w; P.w_rows = ctx->w_rows; P.items = d_items;
    P.job_item_off = (const int64_t*)(base + o_item); P.job_n = (const int32_t*)(base + o_n);
    P.job_vec_off = (const int64_t*)(base + o_vec); P.job_lam_off = (const int64_t*)(base + o_lam);
    P.sig_min = out.d_sig_min; P.job_k = (int32_t*)(base + o_k); P.job_sweeps = (int32_t*)(base + o_sw);
    P.lam_pad = ws.lam_pad.as<double>(); P.vec_pad = ws.vec_pad.as<double>();
    // jobs are sorted by n descending: launch runs that share ceil(n/16)
    int b = 0;
    while (b < nj) {
        const int key = (n[b] + 15) / 16;
        int e = b;
        while (e < nj && (n[e] + 15) / 16 == key) ++e;
        const int nmax = n[b], ept = (nmax + 31) / 32;
        P.job_base = b;
        GsiSpan sp(ctx, GSI_T_EIG_CTA, 1);
        cudaError_t ce;
        switch (ept) {
            case 1: ce = launch_eig_cta<1>(ctx, P, e - b, nmax); break;
            case 2: ce = launch_eig_cta<2>(ctx, P, e - b, nmax); break;
            case 3: ce = launch_eig_cta<3>(ctx, P, e - b, nmax); break;
            case 4: ce = launch_eig_cta<4>(ctx, P, e - b, nmax); break;
            default: ce = launch_eig_cta<5>(ctx, P, e - b, nmax); break;
        }
        sp.end();
        GSI_CUDA(ctx, ce);
        b = e;
    }
    OutJobs J;
    J.nj = nj; J.n = P.job_n; J.k = P.job_k; J.user = (const int64_t*)(base + o_user);
    J.vec_pad = P.job_vec_off; J.lam_pad = P.job_lam_off;
    J.vec_dst = (int64_t*)(base + o_vd); J.lam_dst = (int64_t*)(base + o_ld);
    return finish_chunk(ctx, J, out, max_nk);
}

// ---- large chunk: Laplacian kernels, block-Jacobi rounds, finalisation ------------------------------
static int run_large_chunk(gsi_ctx* ctx, const Job* jobs, int nj, const int32_t* d_items, const RunOut& out) {
    Workspace& ws = WS(ctx);
    const int nmax = jobs[0].n;                    // sorted descending
    const int M = ctx->bj_m, B = M / 2, MM = M * M;
    const int nb = nb_of(nmax, B), ncols = nb * B, half = nb / 2;
    const int splits = (ld_of(nmax) + GSI_BJ_ROWS - 1) / GSI_BJ_ROWS;
    std::vector<int64_t> item_off(nj), vec_off(nj), lam_off(nj), user(nj), g_off(nj), row_off(nj);
    std::vector<int32_t> n(nj), ld(nj);
    int64_t vtot = 0, ltot = 0, gtot = 0, rtot = 0, max_nk = 0;
    for (int j = 0; j < nj; ++j) {
        item_off[j] = jobs[j].item_off; n[j] = jobs[j].n; user[j] = jobs[j].user; ld[j] = ld_of(jobs[j].n);
        vec_off[j] = vtot; lam_off[j] = ltot; g_off[j] = gtot; row_off[j] = rtot;
        vtot += pad_slots(jobs[j].n); ltot += std::max(jobs[j].n, 2);
        gtot += (int64_t)ld[j] * ncols; rtot += ld[j];
        max_nk = std::max(max_nk, pad_slots(jobs[j].n));
    }
    MetaBuilder mb;
    const size_t o_item = mb.add(item_off), o_vec = mb.add(vec_off), o_lam = mb.add(lam_off), o_user = mb.add(user),
                 o_goff = mb.add(g_off), o_roff = mb.add(row_off), o_n = mb.add(n), o_ldim = mb.add(ld);
    const size_t o_k = mb.reserve(nj * 4), o_vd = mb.reserve(nj * 8), o_ld = mb.reserve(nj * 8), o_sigmax = mb.reserve(nj * 4),
                 o_done = mb.reserve(nj * 4), o_smax = mb.reserve(nj * 8), o_sw = mb.reserve(nj * 4), o_rem = mb.reserve(16);
    int rc;
    if ((rc = ws.vec_pad.ensure(ctx, vtot * 8)) != GSI_OK) return rc;
    if ((rc = ws.lam_pad.ensure(ctx, ltot * 8)) != GSI_OK) return rc;
    if ((rc = ws.G.ensure(ctx, gtot * 8)) != GSI_OK) return rc;
    if ((rc = ws.rows.ensure(ctx, rtot * 16)) != GSI_OK) return rc;
    if ((rc = ws.cols.ensure(ctx, (size_t)nj * ncols * 16)) != GSI_OK) return rc;
    if ((rc = ws.perm.ensure(ctx, (size_t)nj * ncols * 4)) != GSI_OK) return rc;
    if ((rc = ws.hpart.ensure(ctx, (size_t)nj * half * splits * MM * 8)) != GSI_OK) return rc;
    if ((rc = ws.q.ensure(ctx, (size_t)nj * half * MM * 8)) != GSI_OK) return rc;
    if ((rc = ws.h_small.ensure(ctx, 64)) != GSI_OK) return rc;
    char* base;
    if ((rc = upload_meta(ctx, mb, &base)) != GSI_OK) return rc;
    LChunk C;
    C.tiled = 0;
    C.nu = nj; C.nb = nb; C.ncols = ncols; C.splits = splits;
    C.n = (const int32_t*)(base + o_n); C.ld = (const int32_t*)(base + o_ldim);
    C.g_off = (const int64_t*)(base + o_goff); C.item_off = (const int64_t*)(base + o_item);
    C.row_off = (const int64_t*)(base + o_roff); C.vec_off = (const int64_t*)(base + o_vec); C.lam_off = (const int64_t*)(base + o_lam);
    C.G = ws.G.as<double>(); C.deg = ws.rows.as<double>(); C.scale = ws.rows.as<double>() + rtot;
    C.colnorm = ws.cols.as<double>(); C.colval = ws.cols.as<double>() + (size_t)nj * ncols;
    C.perm = ws.perm.as<int32_t>();
    C.sigmax = (unsigned int*)(base + o_sigmax); C.done = (int32_t*)(base + o_done);
    C.smax = (unsigned long long*)(base + o_smax); C.sweeps = (int32_t*)(base + o_sw);
    C.k = (int32_t*)(base + o_k); C.remaining = (int32_t*)(base + o_rem);
    C.Hpart = ws.hpart.as<double>(); C.Q = ws.q.as<double>();
    cudaStream_t st = ctx->stream;
    {   // Laplacian stage
        GsiSpan sp(ctx, GSI_T_LAP, 5);
        GSI_CUDA(ctx, cudaMemsetAsync(C.G, 0, gtot * 8, st));
        const int tiles = (nmax + 31) / 32;
        lap_gather_kernel<<<dim3(tiles * tiles, 1, nj), dim3(32, 8), 0, st>>>(C, ctx->d_w, ctx->w_rows, d_items, tiles);
        lap_degree_kernel<<<dim3((nmax + 127) / 128, 1, nj), 128, 0, st>>>(C);
        lap_transform_kernel<<<dim3((nmax + 127) / 128, (nmax + 7) / 8, nj), 128, 0, st>>>(C);
        lap_sigmin_kernel<<<dim3((nmax + 127) / 128, 1, nj), 128, 0, st>>>(C, out.d_sig_min);
        lap_symmetrize_kernel<<<dim3(tiles * tiles, 1, nj), dim3(32, 8), 0, st>>>(C, tiles, 1.0);
        GSI_CUDA(ctx, cudaGetLastError());
        sp.end();
    }
    int32_t* h_rem = ws.h_small.as<int32_t>();
    const size_t inner_smem = (size_t)2 * M * (M + 1) * sizeof(double);
    auto launch_round = [&](int r, bool sample) {
        const dim3 g3(half, splits, nj), g1(half, 1, nj);
        GsiSpan a(ctx, GSI_T_BJ_GRAM, 1, sample ? 1 : 0);
        if (M == 64) bj_gram_kernel<64><<<g3, 128, 0, st>>>(C, r); else bj_gram_kernel<32><<<g3, 128, 0, st>>>(C, r);
        a.end();
        GsiSpan b(ctx, GSI_T_BJ_INNER, 1, sample ? 1 : 0);
        if (M == 64) bj_inner_kernel<64><<<g1, 512, inner_smem, st>>>(C, r); else bj_inner_kernel<32><<<g1, 256, inner_smem, st>>>(C, r);
        b.end();
        GsiSpan c(ctx, GSI_T_BJ_UPDATE, 1, sample ? 1 : 0);
        if (M == 64) bj_update_kernel<64><<<g3, 128, 0, st>>>(C, r); else bj_update_kernel<32><<<g3, 128, 0, st>>>(C, r);
        c.end();
    };
    for (int sweep = 0; sweep < GSI_MAX_SWEEPS; ++sweep) {
        // diagonal round first (within-block pairs), then the nb-1 circle rounds (cross pairs)
        for (int r = -1; r < nb - 1; ++r) launch_round(r, ctx->timing && ((r + 1) % 8 == 0));
        GSI_CUDA(ctx, cudaMemsetAsync(C.remaining, 0, 4, st));
        bj_check_kernel<<<(nj + 255) / 256, 256, 0, st>>>(C);
        GSI_CUDA(ctx, cudaGetLastError());
        GSI_CUDA(ctx, cudaMemcpyAsync(h_rem, C.remaining, 4, cudaMemcpyDeviceToHost, st));
        GSI_CUDA(ctx, cudaStreamSynchronize(st));
        if (*h_rem == 0) break;
    }
    {
        GsiSpan sp(ctx, GSI_T_FINALIZE, 3);
        fin_colnorm_kernel<<<dim3((ncols + 7) / 8, 1, nj), 256, 0, st>>>(C);
        fin_rank_kernel<<<nj, 1024, 0, st>>>(C, ws.lam_pad.as<double>());
        const int tiles_i = (nmax + 31) / 32, tiles_r = (nmax + 31) / 32;
        fin_emit_kernel<<<dim3(tiles_i * tiles_r, 1, nj), dim3(32, 8), 0, st>>>(C, ws.vec_pad.as<double>(), tiles_r);
        GSI_CUDA(ctx, cudaGetLastError());
        sp.end();
    }
    OutJobs J;
    J.nj = nj; J.n = C.n; J.k = C.k; J.user = (const int64_t*)(base + o_user);
    J.vec_pad = C.vec_off; J.lam_pad = C.lam_off;
    J.vec_dst = (int64_t*)(base + o_vd); J.lam_dst = (int64_t*)(base + o_ld);
    return finish_chunk(ctx, J, out, max_nk);
}

#include "hh_host.cuh"
#include "lc_host.cuh"

static int check_csr(gsi_ctx* ctx, int64_t nu, const int64_t* off, const int32_t* items) {
    if (off[0] != 0) return gsi_fail(ctx, GSI_ERR_INVALID, "offsets[0] must be 0");
    for (int64_t u = 0; u < nu; ++u) {
        if (off[u + 1] <= off[u]) return gsi_fail(ctx, GSI_ERR_INVALID, "user %lld has no rated movies", (long long)u);
        if (items)
            for (int64_t t = off[u] + 1; t < off[u + 1]; ++t)
                if (items[t] <= items[t - 1])
                    return gsi_fail(ctx, GSI_ERR_INVALID, "user %lld: movie ids must be strictly ascending (unique)", (long long)u);
    }
    return GSI_OK;
}

// ---- device API -----------------------------------------------------------------------------
extern "C" int gsi_precompute_device(gsi_ctx* ctx, int64_t nu, const int64_t* h_offsets, const int32_t* d_items,
                                     double* d_sig_min, int32_t* d_k, int64_t* d_lam_off, int64_t* d_vec_off,
                                     double* d_lam, int64_t lam_cap, double* d_vec, int64_t vec_cap, int64_t* totals) {
    if (!ctx) return gsi_fail(nullptr, GSI_ERR_INVALID, "null context");
    if (!ctx->d_w) return gsi_fail(ctx, GSI_ERR_STATE, "gsi_precompute: no weight table set (call gsi_set_weights_*)");
    if (nu < 0 || !h_offsets || (nu > 0 && (!d_items || !d_sig_min || !d_k || !d_lam_off || !d_vec_off || !d_lam || !d_vec)))
        return gsi_fail(ctx, GSI_ERR_INVALID, "gsi_precompute_device: null argument");
    GSI_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = check_csr(ctx, nu, h_offsets, nullptr)) != GSI_OK) return rc;
    Workspace& ws = WS(ctx);
    if ((rc = ws.totals.ensure(ctx, 64)) != GSI_OK) return rc;
    if ((rc = ws.h_small.ensure(ctx, 64)) != GSI_OK) return rc;
    GSI_CUDA(ctx, cudaMemsetAsync(ws.totals.p, 0, 64, ctx->stream));
    std::vector<Job> small, large;
    std::vector<Chunk> chunks;
    if ((rc = plan(ctx, nu, h_offsets, small, large, chunks)) != GSI_OK) return rc;
    RunOut out{d_lam, lam_cap, d_vec, vec_cap, ws.totals.as<int64_t>(), d_k, d_lam_off, d_vec_off, d_sig_min};
    for (const Chunk& c : chunks) {
        rc = c.large ? (c.bj ? run_large_chunk : run_hh_chunk)(ctx, large.data() + c.begin, c.end - c.begin, d_items, out)
                     : run_small_chunk(ctx, small.data() + c.begin, c.end - c.begin, d_items, out);
        if (rc != GSI_OK) return rc;
    }
    int64_t* h_tot = (int64_t*)(ws.h_small.as<char>() + 16);
    GSI_CUDA(ctx, cudaMemcpyAsync(h_tot, ws.totals.p, 24, cudaMemcpyDeviceToHost, ctx->stream));
    GSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (totals) { totals[0] = h_tot[0]; totals[1] = h_tot[1]; }
    if (h_tot[2]) return gsi_fail(ctx, GSI_ERR_CAPACITY, "output capacity too small: need lam %lld / vec %lld doubles",
                                  (long long)h_tot[0], (long long)h_tot[1]);
    return GSI_OK;
}

// ---- streaming host API ------------------------------------------------------------------------
extern "C" int gsi_precompute_stream(gsi_ctx* ctx, int64_t nu, const int64_t* offsets, const int32_t* items,
                                     gsi_record_sink sink, void* opaque) {
    if (!ctx) return gsi_fail(nullptr, GSI_ERR_INVALID, "null context");
    if (!ctx->d_w) return gsi_fail(ctx, GSI_ERR_STATE, "gsi_precompute: no weight table set (call gsi_set_weights_*)");
    if (nu < 0 || !offsets || !sink || (nu > 0 && !items)) return gsi_fail(ctx, GSI_ERR_INVALID, "gsi_precompute_stream: null argument");
    GSI_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = check_csr(ctx, nu, offsets, items)) != GSI_OK) return rc;
    if (nu == 0) return GSI_OK;
    Workspace& ws = WS(ctx);
    const int64_t nnz = offsets[nu];
    if ((rc = ws.totals.ensure(ctx, 64)) != GSI_OK) return rc;
    if ((rc = ws.h_small.ensure(ctx, 256)) != GSI_OK) return rc;
    if ((rc = ws.items.ensure(ctx, nnz * 4)) != GSI_OK) return rc;
    if ((rc = ws.sig.ensure(ctx, nnz * 8)) != GSI_OK) return rc;
    if ((rc = ws.outk.ensure(ctx, nu * 4)) != GSI_OK) return rc;
    if ((rc = ws.outlam.ensure(ctx, nu * 8)) != GSI_OK) return rc;
    if ((rc = ws.outvec.ensure(ctx, nu * 8)) != GSI_OK) return rc;
    if ((rc = ws.h_k.ensure(ctx, nu * 4)) != GSI_OK) return rc;
    if ((rc = ws.h_lamoff.ensure(ctx, nu * 8)) != GSI_OK) return rc;
    if ((rc = ws.h_vecoff.ensure(ctx, nu * 8)) != GSI_OK) return rc;
    if ((rc = ws.h_sig.ensure(ctx, nnz * 8)) != GSI_OK) return rc;
    GSI_CUDA(ctx, cudaMemcpyAsync(ws.items.p, items, nnz * 4, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<Job> small, large;
    std::vector<Chunk> chunks;
    if ((rc = plan(ctx, nu, offsets, small, large, chunks)) != GSI_OK) return rc;
    std::vector<int64_t> rec_user, rec_lam, rec_vec;
    std::vector<int32_t> rec_n, rec_k;
    for (const Chunk& c : chunks) {
        const Job* jobs = (c.large ? large.data() : small.data()) + c.begin;
        const int nj = c.end - c.begin;
        int64_t vcap = 0, lcap = 0;
        for (int j = 0; j < nj; ++j) { vcap += pad_slots(jobs[j].n); lcap += std::max(jobs[j].n, 2); }
        if ((rc = ws.stage_vec.ensure(ctx, vcap * 8)) != GSI_OK) return rc;
        if ((rc = ws.stage_lam.ensure(ctx, lcap * 8)) != GSI_OK) return rc;
        GSI_CUDA(ctx, cudaMemsetAsync(ws.totals.p, 0, 64, ctx->stream));
        RunOut out{ws.stage_lam.as<double>(), lcap, ws.stage_vec.as<double>(), vcap, ws.totals.as<int64_t>(),
                   ws.outk.as<int32_t>(), ws.outlam.as<int64_t>(), ws.outvec.as<int64_t>(), ws.sig.as<double>()};
        bool vec_copied = false;
        if (c.large && !c.bj) {                          // the Householder path copies its records group by group (hh_host.cuh)
            if ((rc = ws.h_stage_vec.ensure(ctx, vcap * 8)) != GSI_OK) return rc;
            out.h_vec = ws.h_stage_vec.as<double>();
            out.h_bounds = (int64_t*)(ws.h_small.as<char>() + 64);
            out.vec_copied = &vec_copied;
        }
        rc = c.large ? (c.bj ? run_large_chunk : run_hh_chunk)(ctx, jobs, nj, ws.items.as<int32_t>(), out)
                     : run_small_chunk(ctx, jobs, nj, ws.items.as<int32_t>(), out);
        if (rc != GSI_OK) return rc;
        int64_t* h_tot = (int64_t*)(ws.h_small.as<char>() + 16);
        GSI_CUDA(ctx, cudaMemcpyAsync(h_tot, ws.totals.p, 24, cudaMemcpyDeviceToHost, ctx->stream));
        GSI_CUDA(ctx, cudaMemcpyAsync(ws.h_k.p, ws.outk.p, nu * 4, cudaMemcpyDeviceToHost, ctx->stream));
        GSI_CUDA(ctx, cudaMemcpyAsync(ws.h_lamoff.p, ws.outlam.p, nu * 8, cudaMemcpyDeviceToHost, ctx->stream));
        GSI_CUDA(ctx, cudaMemcpyAsync(ws.h_vecoff.p, ws.outvec.p, nu * 8, cudaMemcpyDeviceToHost, ctx->stream));
        GSI_CUDA(ctx, cudaMemcpyAsync(ws.h_sig.p, ws.sig.p, nnz * 8, cudaMemcpyDeviceToHost, ctx->stream));
        GSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (h_tot[2]) return gsi_fail(ctx, GSI_ERR_CAPACITY, "internal staging overflow");
        if ((rc = ws.h_stage_lam.ensure(ctx, h_tot[0] * 8)) != GSI_OK) return rc;
        if ((rc = ws.h_stage_vec.ensure(ctx, h_tot[1] * 8)) != GSI_OK) return rc;
        GSI_CUDA(ctx, cudaMemcpyAsync(ws.h_stage_lam.p, ws.stage_lam.p, h_tot[0] * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (!vec_copied) GSI_CUDA(ctx, cudaMemcpyAsync(ws.h_stage_vec.p, ws.stage_vec.p, h_tot[1] * 8, cudaMemcpyDeviceToHost, ctx->stream));
        GSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (vec_copied) GSI_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
        rec_user.resize(nj); rec_lam.resize(nj); rec_vec.resize(nj); rec_n.resize(nj); rec_k.resize(nj);
        for (int j = 0; j < nj; ++j) {
            const int64_t u = jobs[j].user;
            rec_user[j] = u; rec_n[j] = jobs[j].n; rec_k[j] = ws.h_k.as<int32_t>()[u];
            rec_lam[j] = ws.h_lamoff.as<int64_t>()[u]; rec_vec[j] = ws.h_vecoff.as<int64_t>()[u];
        }
        gsi_record_chunk ch;
        ch.n_records = nj; ch.user_index = rec_user.data(); ch.n = rec_n.data(); ch.k = rec_k.data();
        ch.lam_off = rec_lam.data(); ch.vec_off = rec_vec.data();
        ch.lam = ws.h_stage_lam.as<double>(); ch.vec = ws.h_stage_vec.as<double>(); ch.sig_min = ws.h_sig.as<double>();
        if (sink(opaque, &ch) != 0) return gsi_fail(ctx, GSI_ERR_SINK, "record sink failed");
    }
    return GSI_OK;
}

// ---- array host API (memcpy sink over the streaming path) ----------------------------------------
struct HostSink {
    const int64_t* offsets; double* sig_min; int32_t* k; int64_t* lam_off; int64_t* vec_off;
    double* lam; int64_t lam_cap; double* vec; int64_t vec_cap; int64_t lam_used = 0, vec_used = 0; bool overflow = false;
};
static int host_sink(void* opaque, const gsi_record_chunk* ch) {
    HostSink* s = (HostSink*)opaque;
    for (int64_t j = 0; j < ch->n_records; ++j) {
        const int64_t u = ch->user_index[j];
        const int n = ch->n[j], k = ch->k[j];
        s->k[u] = k; s->lam_off[u] = s->lam_used; s->vec_off[u] = s->vec_used;
        if (s->lam_used + k > s->lam_cap || s->vec_used + (int64_t)n * k > s->vec_cap) s->overflow = true;
        else {
            memcpy(s->lam + s->lam_used, ch->lam + ch->lam_off[j], (size_t)k * 8);
            memcpy(s->vec + s->vec_used, ch->vec + ch->vec_off[j], (size_t)n * k * 8);
        }
        s->lam_used += k; s->vec_used += (int64_t)n * k;
        memcpy(s->sig_min + s->offsets[u], ch->sig_min + s->offsets[u], (size_t)n * 8);
    }
    return 0;
}
extern "C" int gsi_precompute_host(gsi_ctx* ctx, int64_t nu, const int64_t* offsets, const int32_t* items, double* sig_min,
                                   int32_t* k, int64_t* lam_off, int64_t* vec_off, double* lam, int64_t lam_cap,
                                   double* vec, int64_t vec_cap, int64_t* totals) {
    if (!ctx) return gsi_fail(nullptr, GSI_ERR_INVALID, "null context");
    if (nu > 0 && (!sig_min || !k || !lam_off || !vec_off || !lam || !vec)) return gsi_fail(ctx, GSI_ERR_INVALID, "gsi_precompute_host: null argument");
    HostSink s{offsets, sig_min, k, lam_off, vec_off, lam, lam_cap, vec, vec_cap};
    int rc = gsi_precompute_stream(ctx, nu, offsets, items, host_sink, &s);
    if (totals) { totals[0] = s.lam_used; totals[1] = s.vec_used; }
    if (rc != GSI_OK) return rc;
    if (s.overflow) return gsi_fail(ctx, GSI_ERR_CAPACITY, "output capacity too small: need lam %lld / vec %lld doubles",
                                    (long long)s.lam_used, (long long)s.vec_used);
    return GSI_OK;
}

// ---- predict ------------------------------------------------------------------------------------
struct PredTask { int64_t pair; int32_t user; int32_t k; int32_t n; };

extern "C" int gsi_predict_device(gsi_ctx* ctx, int64_t nu, const int64_t* h_off, const int32_t* h_k,
                                  const int64_t* d_off, const int32_t* d_items, const double* d_w_lim,
                                  const double* d_ratings, const int32_t* d_k, const int64_t* d_lam_off,
                                  const int64_t* d_vec_off, const double* d_lam, const double* d_vec,
                                  const uint8_t* pair_mask, float* d_err, int32_t* d_kk, double* d_pred,
                                  int32_t* d_status, int32_t* d_cols, int64_t* n_pairs_done) {
    if (!ctx) return gsi_fail(nullptr, GSI_ERR_INVALID, "null context");
    if (!ctx->d_w) return gsi_fail(ctx, GSI_ERR_STATE, "gsi_predict: no weight table set (call gsi_set_weights_*)");
    if (nu < 0 || !h_off || !h_k || (nu > 0 && (!d_off || !d_items || !d_w_lim || !d_ratings || !d_k || !d_lam_off || !d_vec_off ||
                                                  !d_lam || !d_vec || !d_err || !d_kk || !d_pred || !d_status || !d_cols)))
        return gsi_fail(ctx, GSI_ERR_INVALID, "gsi_predict_device: null argument");
    GSI_CUDA(ctx, cudaSetDevice(ctx->device));
    Workspace& ws = WS(ctx);
    const int64_t nnz = nu ? h_off[nu] : 0;
    // status defaults to SKIPPED, err to 0
    if (nnz) {
        std::vector<int32_t> skipped((size_t)nnz, GSI_PRED_SKIPPED);
        GSI_CUDA(ctx, cudaMemcpyAsync(d_status, skipped.data(), nnz * 4, cudaMemcpyHostToDevice, ctx->stream));
        GSI_CUDA(ctx, cudaMemsetAsync(d_err, 0, nnz * 4, ctx->stream));
        GSI_CUDA(ctx, cudaMemsetAsync(d_kk, 0, nnz * 4, ctx->stream));
        GSI_CUDA(ctx, cudaMemsetAsync(d_pred, 0, nnz * 8, ctx->stream));
        GSI_CUDA(ctx, cudaMemsetAsync(d_cols, 0, nnz * 4, ctx->stream));
        GSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    // ---- per-user sums (complement-row form) and the per-pair plan (kk, lim) ----
    int rc;
    std::vector<int64_t> sum_off((size_t)nu + 1, 0);
    for (int64_t u = 0; u < nu; ++u) sum_off[u + 1] = sum_off[u] + h_k[u];
    if ((rc = ws.p_gsum.ensure(ctx, std::max<int64_t>(1, sum_off[nu]) * 8)) != GSI_OK) return rc;
    if ((rc = ws.p_hsum.ensure(ctx, std::max<int64_t>(1, sum_off[nu]) * 8)) != GSI_OK) return rc;
    if ((rc = ws.p_sumoff.ensure(ctx, (size_t)(nu + 1) * 8)) != GSI_OK) return rc;
    if ((rc = ws.p_exact.ensure(ctx, std::max<int64_t>(1, nu))) != GSI_OK) return rc;
    if ((rc = ws.p_lim.ensure(ctx, std::max<int64_t>(1, nnz) * 4)) != GSI_OK) return rc;
    if ((rc = ws.p_mask.ensure(ctx, std::max<int64_t>(1, nnz))) != GSI_OK) return rc;
    PredParams P;
    memset(&P, 0, sizeof P);
    P.W = ctx->d_w; P.w_rows = ctx->w_rows; P.items = d_items; P.offsets = d_off; P.w_lim = d_w_lim; P.ratings = d_ratings;
    P.k = d_k; P.lam_off = d_lam_off; P.vec_off = d_vec_off; P.lam = d_lam; P.vec = d_vec;
    P.err = d_err; P.kk = d_kk; P.pred = d_pred; P.status = d_status; P.cols_used = d_cols;
    P.gsum = ws.p_gsum.as<double>(); P.hsum = ws.p_hsum.as<double>(); P.sum_off = ws.p_sumoff.as<int64_t>();
    P.user_exact = ws.p_exact.as<uint8_t>(); P.plan_lim = ws.p_lim.as<int32_t>();
    P.pair_mask_dev = nullptr;
    std::vector<int32_t> h_kk((size_t)nnz), h_lim((size_t)nnz);
    std::vector<uint8_t> h_exact((size_t)nu);
    if (nnz) {
        GSI_CUDA(ctx, cudaMemcpyAsync(ws.p_sumoff.p, sum_off.data(), (size_t)(nu + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        if (pair_mask) {
            GSI_CUDA(ctx, cudaMemcpyAsync(ws.p_mask.p, pair_mask, (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream));
            P.pair_mask_dev = ws.p_mask.as<uint8_t>();
        }
        GsiSpan sp(ctx, GSI_T_PREDICT, 2);
        {
            HhTrace tr(ctx, "predict: per-user sums + per-pair plan");
            pred_user_sums_kernel<<<(unsigned)nu, 256, 0, ctx->stream>>>(P, ws.p_gsum.as<double>(), ws.p_hsum.as<double>(), ws.p_exact.as<uint8_t>(), (int)nu);
            pred_plan_kernel<<<(unsigned)nu, 256, 0, ctx->stream>>>(P, (int)nu);
        }
        sp.end();
        GSI_CUDA(ctx, cudaGetLastError());
        GSI_CUDA(ctx, cudaMemcpyAsync(h_kk.data(), d_kk, (size_t)nnz * 4, cudaMemcpyDeviceToHost, ctx->stream));
        GSI_CUDA(ctx, cudaMemcpyAsync(h_lim.data(), ws.p_lim.p, (size_t)nnz * 4, cudaMemcpyDeviceToHost, ctx->stream));
        GSI_CUDA(ctx, cudaMemcpyAsync(h_exact.data(), ws.p_exact.p, (size_t)nu, cudaMemcpyDeviceToHost, ctx->stream));
        GSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    const bool no_wood = getenv("GSI_PRED_DIRECT") && atoi(getenv("GSI_PRED_DIRECT")) != 0;   // probe / test: direct form only
    if (no_wood) P.user_exact = nullptr;
    // classes by the ORDER of the pair's solve: min(lim, n - kk) when the complement-row form applies, else lim
    // (pairs up to order 192 keep everything in shared memory, bigger ones use a per-CTA scratch in L2)
    static const int kClassC[] = {16, 32, 48, 64, 96, 128, 160, 192};
    const int n_smem_classes = 8;
    struct ClassAcc { std::vector<PredTask> tasks; int nmax = 0, kmax = 0; };
    std::vector<ClassAcc> cls(n_smem_classes + 1);
    for (int64_t u = 0; u < nu; ++u) {
        const int n = (int)(h_off[u + 1] - h_off[u]);
        for (int i = 0; i < n; ++i) {
            const int64_t pair = h_off[u] + i;
            if (pair_mask && !pair_mask[pair]) continue;
            const int lim = h_lim[pair], nr = n - h_kk[pair];
            const int dim = (!no_wood && h_exact[u] && nr < lim) ? nr : lim;
            int c = 0;
            while (c < n_smem_classes && (dim > kClassC[c] || predict2_smem_bytes(kClassC[c], n, true, P2_R, lim) > 226 * 1024)) ++c;
            cls[c].tasks.push_back({pair, (int32_t)u, dim, n});
            cls[c].nmax = std::max(cls[c].nmax, n); cls[c].kmax = std::max(cls[c].kmax, lim);
        }
    }
    int64_t done = 0;
    for (int c = 0; c <= n_smem_classes; ++c) {
        std::vector<PredTask>& tasks = cls[c].tasks;
        if (tasks.empty()) continue;
        const bool in_smem = c < n_smem_classes;
        // big tasks: descending order so that a wave has similar cost
        if (!in_smem) std::stable_sort(tasks.begin(), tasks.end(), [](const PredTask& a, const PredTask& b) { return a.k > b.k; });
        std::vector<int64_t> t_pair(tasks.size());
        std::vector<int32_t> t_user(tasks.size());
        const int nmax = cls[c].nmax, kmax = std::max(cls[c].kmax, 2);
        for (size_t t = 0; t < tasks.size(); ++t) { t_pair[t] = tasks[t].pair; t_user[t] = tasks[t].user; }
        MetaBuilder mb;
        const size_t o_pair = mb.add(t_pair), o_user = mb.add(t_user);
        if ((rc = ws.pred_meta.ensure(ctx, mb.host.size())) != GSI_OK) return rc;
        GSI_CUDA(ctx, cudaMemcpyAsync(ws.pred_meta.p, mb.host.data(), mb.host.size(), cudaMemcpyHostToDevice, ctx->stream));
        GSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        P.task_pair = (const int64_t*)(ws.pred_meta.as<char>() + o_pair);
        P.task_user = (const int32_t*)(ws.pred_meta.as<char>() + o_user);
        P.work = nullptr; P.work_stride = 0; P.nmax = nmax; P.kmax = kmax; P.m_in_smem = in_smem ? 1 : 0;
        if (in_smem) {
            P.cmax = kClassC[c];
            P.chunk_rows = P2_R;
            const size_t smem = predict2_smem_bytes(P.cmax, nmax, true, P2_R, kmax);
            GSI_CUDA(ctx, cudaFuncSetAttribute(predict2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
            const int64_t wave = 1 << 20;
            for (int64_t b = 0; b < (int64_t)tasks.size(); b += wave) {
                const int cnt = (int)std::min<int64_t>(wave, tasks.size() - b);
                P.task_base = (int)b;
                char label[112];
                snprintf(label, sizeof label, "predict2 order<=%d: %d pairs, nmax %d, kmax %d, %zu B smem", P.cmax, cnt, nmax, kmax, smem);
                HhTrace tr(ctx, label);
                GsiSpan sp(ctx, GSI_T_PREDICT, 1);
                predict2_kernel<<<cnt, 256, smem, ctx->stream>>>(P);
                sp.end();
                GSI_CUDA(ctx, cudaGetLastError());
            }
        } else {
            // waves of CTAs, each with a private scratch for its matrix in global memory
            size_t b = 0;
            while (b < tasks.size()) {
                const int kwave = tasks[b].k;                       // largest order of this wave (sorted)
                const int64_t stride = (int64_t)predict2_tiles_dbl(kwave);
                int64_t wave = std::max<int64_t>(1, std::min<int64_t>(4 * ctx->sm_count, (ctx->ws_limit / 2) / (stride * 8)));
                wave = std::min<int64_t>(wave, tasks.size() - b);
                if ((rc = ws.pred_work.ensure(ctx, (size_t)wave * stride * 8)) != GSI_OK) return rc;
                P.cmax = kwave; P.work = ws.pred_work.as<double>(); P.work_stride = stride; P.task_base = (int)b;
                P.chunk_rows = 64;                                  // as many staged rows as fit beside the index arrays
                static const size_t stage_cap = getenv("GSI_PRED_STAGE_KB") ? (size_t)atoi(getenv("GSI_PRED_STAGE_KB")) * 1024 : 110 * 1024;   // two CTAs per SM
                while (P.chunk_rows > 8 && predict2_smem_bytes(P.cmax, nmax, false, P.chunk_rows, kmax) > stage_cap) P.chunk_rows /= 2;
                const size_t smem = predict2_smem_bytes(P.cmax, nmax, false, P.chunk_rows, kmax);
                if (smem > 226 * 1024) return gsi_fail(ctx, GSI_ERR_INVALID, "predict: user too large for the staging buffers (n=%d, order=%d)", nmax, kwave);
                GSI_CUDA(ctx, cudaFuncSetAttribute(predict2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
                char label[112];
                snprintf(label, sizeof label, "predict2 (M in L2) order<=%d: %d pairs, nmax %d, %zu B smem", P.cmax, (int)wave, nmax, smem);
                HhTrace tr(ctx, label);
                GsiSpan sp(ctx, GSI_T_PREDICT, 1);
                predict2_kernel<<<(int)wave, 256, smem, ctx->stream>>>(P);
                sp.end();
                GSI_CUDA(ctx, cudaGetLastError());
                b += wave;
            }
        }
        done += (int64_t)tasks.size();
        GSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // pred_meta is reused by the next class
    }
    if (n_pairs_done) *n_pairs_done = done;
    return GSI_OK;
}

extern "C" int gsi_predict_host(gsi_ctx* ctx, int64_t nu, const int64_t* offsets, const int32_t* items, const double* w_lim,
                                const double* ratings, const int32_t* k, const int64_t* lam_off, const int64_t* vec_off,
                                const double* lam, int64_t lam_len, const double* vec, int64_t vec_len, const uint8_t* pair_mask,
                                float* err, int32_t* kk, double* pred, int32_t* status, int32_t* cols) {
    if (!ctx) return gsi_fail(nullptr, GSI_ERR_INVALID, "null context");
    if (nu < 0 || !offsets || (nu > 0 && (!items || !w_lim || !ratings || !k || !lam_off || !vec_off || !lam || !vec || !err || !kk || !pred || !status || !cols)))
        return gsi_fail(ctx, GSI_ERR_INVALID, "gsi_predict_host: null argument");
    if (nu == 0) return GSI_OK;
    GSI_CUDA(ctx, cudaSetDevice(ctx->device));
    Workspace& ws = WS(ctx);
    const int64_t nnz = offsets[nu];
    for (int64_t u = 0; u < nu; ++u) {
        const int64_t n = offsets[u + 1] - offsets[u];
        if (n < 1 || k[u] < 1) return gsi_fail(ctx, GSI_ERR_INVALID, "user %lld: empty record", (long long)u);
        if (lam_off[u] < 0 || lam_off[u] + k[u] > lam_len || vec_off[u] < 0 || vec_off[u] + n * k[u] > vec_len)
            return gsi_fail(ctx, GSI_ERR_INVALID, "user %lld: record offsets outside lam/vec", (long long)u);
    }
    int rc;
    struct Up { DevBuf* b; const void* src; size_t bytes; };
    Up ups[] = {{&ws.p_off, offsets, (size_t)(nu + 1) * 8}, {&ws.p_items, items, (size_t)nnz * 4}, {&ws.p_wlim, w_lim, (size_t)nnz * 8},
                {&ws.p_rat, ratings, (size_t)nnz * 8}, {&ws.p_k, k, (size_t)nu * 4}, {&ws.p_lamoff, lam_off, (size_t)nu * 8},
                {&ws.p_vecoff, vec_off, (size_t)nu * 8}, {&ws.p_lam, lam, (size_t)lam_len * 8}, {&ws.p_vec, vec, (size_t)vec_len * 8}};
    for (auto& x : ups) {
        if ((rc = x.b->ensure(ctx, x.bytes)) != GSI_OK) return rc;
        GSI_CUDA(ctx, cudaMemcpyAsync(x.b->p, x.src, x.bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    if ((rc = ws.p_err.ensure(ctx, nnz * 4)) != GSI_OK) return rc;
    if ((rc = ws.p_kk.ensure(ctx, nnz * 4)) != GSI_OK) return rc;
    if ((rc = ws.p_pred.ensure(ctx, nnz * 8)) != GSI_OK) return rc;
    if ((rc = ws.p_status.ensure(ctx, nnz * 4)) != GSI_OK) return rc;
    if ((rc = ws.p_cols.ensure(ctx, nnz * 4)) != GSI_OK) return rc;
    GSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    rc = gsi_predict_device(ctx, nu, offsets, k, ws.p_off.as<int64_t>(), ws.p_items.as<int32_t>(), ws.p_wlim.as<double>(),
                            ws.p_rat.as<double>(), ws.p_k.as<int32_t>(), ws.p_lamoff.as<int64_t>(), ws.p_vecoff.as<int64_t>(),
                            ws.p_lam.as<double>(), ws.p_vec.as<double>(), pair_mask, ws.p_err.as<float>(), ws.p_kk.as<int32_t>(),
                            ws.p_pred.as<double>(), ws.p_status.as<int32_t>(), ws.p_cols.as<int32_t>(), nullptr);
    if (rc != GSI_OK) return rc;
    GSI_CUDA(ctx, cudaMemcpyAsync(err, ws.p_err.p, nnz * 4, cudaMemcpyDeviceToHost, ctx->stream));
    GSI_CUDA(ctx, cudaMemcpyAsync(kk, ws.p_kk.p, nnz * 4, cudaMemcpyDeviceToHost, ctx->stream));
    GSI_CUDA(ctx, cudaMemcpyAsync(pred, ws.p_pred.p, nnz * 8, cudaMemcpyDeviceToHost, ctx->stream));
    GSI_CUDA(ctx, cudaMemcpyAsync(status, ws.p_status.p, nnz * 4, cudaMemcpyDeviceToHost, ctx->stream));
    GSI_CUDA(ctx, cudaMemcpyAsync(cols, ws.p_cols.p, nnz * 4, cudaMemcpyDeviceToHost, ctx->stream));
    GSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GSI_OK;
}

// ---- knn chain ------------------------------------------------------------------------------------
static int knn_upload_csr(gsi_ctx* ctx, int64_t nu, const int64_t* off, const int32_t* items, const float* ratings, int* nmax) {
    Workspace& ws = WS(ctx);
    int rc;
    if (off[0] != 0) return gsi_fail(ctx, GSI_ERR_INVALID, "offsets[0] must be 0");
    int mx = 0;
    for (int64_t u = 0; u < nu; ++u) {
        if (off[u + 1] < off[u]) return gsi_fail(ctx, GSI_ERR_INVALID, "offsets must be non-decreasing");
        mx = std::max<int64_t>(mx, off[u + 1] - off[u]);
    }
    *nmax = mx;
    const int64_t nnz = off[nu];
    if ((rc = ws.k_off.ensure(ctx, (nu + 1) * 8)) != GSI_OK) return rc;
    if ((rc = ws.k_items.ensure(ctx, std::max<int64_t>(nnz, 1) * 4)) != GSI_OK) return rc;
    GSI_CUDA(ctx, cudaMemcpyAsync(ws.k_off.p, off, (nu + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    GSI_CUDA(ctx, cudaMemcpyAsync(ws.k_items.p, items, nnz * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (ratings) {
        if ((rc = ws.k_rat.ensure(ctx, std::max<int64_t>(nnz, 1) * 4)) != GSI_OK) return rc;
        GSI_CUDA(ctx, cudaMemcpyAsync(ws.k_rat.p, ratings, nnz * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    return GSI_OK;
}


// ---- test / measurement hook of the tcgen05 FP64-equivalent GEMM (tc_gemm.cu) ----------------------------------------
extern "C" int gsi_debug_tc_gemm(gsi_ctx* ctx, int m, int n, int k, const double* a, int64_t lda, const double* b, int64_t ldb,
                                 double* c, int64_t ldc, int slices, int reps, double* ms_slice, double* ms_gemm) {
    if (!ctx) return gsi_fail(nullptr, GSI_ERR_INVALID, "null context");
    if (m < 1 || n < 1 || k < 1 || !a || !b || !c || lda < m || ldb < k || ldc < m || slices < 6 || slices > 8 || reps < 1)
        return gsi_fail(ctx, GSI_ERR_INVALID, "gsi_debug_tc_gemm: bad argument");
    GSI_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    double *dA = nullptr, *dB = nullptr, *dC = nullptr;
    int8_t *Ap = nullptr, *Bp = nullptr;
    int32_t *ea = nullptr, *eb = nullptr;
    TcTask* dT = nullptr;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    int rc = GSI_OK;
    auto fin = [&]() {
        cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(Ap); cudaFree(Bp); cudaFree(ea); cudaFree(eb); cudaFree(dT);
        for (auto& e : ev) if (e) cudaEventDestroy(e);
    };
#define TCG(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fin(); return gsi_fail(ctx, GSI_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } } while (0)
    TCG(cudaMalloc((void**)&dA, (size_t)lda * k * 8));
    TCG(cudaMalloc((void**)&dB, (size_t)ldb * n * 8));
    TCG(cudaMalloc((void**)&dC, (size_t)ldc * n * 8));
    TCG(cudaMalloc((void**)&Ap, tc_gemm_plane_bytes_a(m, k, slices)));
    TCG(cudaMalloc((void**)&Bp, tc_gemm_plane_bytes_b(k, n, slices)));
    TCG(cudaMalloc((void**)&ea, (size_t)m * 4));
    TCG(cudaMalloc((void**)&eb, (size_t)n * 4));
    for (auto& e : ev) TCG(cudaEventCreate(&e));
    TCG(cudaMemcpyAsync(dA, a, (size_t)lda * k * 8, cudaMemcpyHostToDevice, st));
    TCG(cudaMemcpyAsync(dB, b, (size_t)ldb * n * 8, cudaMemcpyHostToDevice, st));
    TCG(cudaMemsetAsync(dC, 0, (size_t)ldc * n * 8, st));
    TcTask h[2];
    memset(h, 0, sizeof h);
    h[0].A = dA; h[0].lda = lda; h[0].B = dB; h[0].ldb = ldb; h[0].C = dC; h[0].ldc = ldc; h[0].M = m; h[0].N = n; h[0].K = k;
    h[0].Ap = Ap; h[0].Bp = Bp; h[0].ea = ea; h[0].eb = eb;
    TCG(cudaMalloc((void**)&dT, sizeof h));
    TCG(cudaMemcpyAsync(dT, h, sizeof h, cudaMemcpyHostToDevice, st));
    TcBatch A;
    A.tasks = dT; A.ntasks = 1; A.Mmax = m; A.Nmax = n; A.Kmax = k; A.S = slices;
    float t_slice = 0.f, t_gemm = 0.f;
    for (int r = 0; r < reps; ++r) {
        TCG(cudaEventRecord(ev[0], st));
        TCG(tc_gemm_slice(A, st));
        TCG(cudaEventRecord(ev[1], st));
        TCG(tc_gemm_mma(A, st, ctx->sm_count));
        TCG(cudaEventRecord(ev[2], st));
    }
    TCG(cudaStreamSynchronize(st));
    TCG(cudaEventElapsedTime(&t_slice, ev[0], ev[1]));
    TCG(cudaEventElapsedTime(&t_gemm, ev[1], ev[2]));
    TCG(cudaMemcpyAsync(c, dC, (size_t)ldc * n * 8, cudaMemcpyDeviceToHost, st));
    TCG(cudaStreamSynchronize(st));
#undef TCG
    if (ms_slice) *ms_slice = t_slice;
    if (ms_gemm) *ms_gemm = t_gemm;
    fin();
    return rc;
}

extern "C" int gsi_knn_build_host(gsi_ctx* ctx, int64_t nu, const int64_t* off, const int32_t* items, const float* ratings,
                                  int rows, int install_weights, int64_t* n_edges) {
    if (!ctx) return gsi_fail(nullptr, GSI_ERR_INVALID, "null context");
    if (nu < 0 || rows < 1 || !off || (nu > 0 && (!items || !ratings))) return gsi_fail(ctx, GSI_ERR_INVALID, "gsi_knn_build_host: bad argument");
    GSI_CUDA(ctx, cudaSetDevice(ctx->device));
    Workspace& ws = WS(ctx);
    int rc, nmax = 0;
    if ((rc = knn_upload_csr(ctx, nu, off, items, ratings, &nmax)) != GSI_OK) return rc;
    const size_t nn = (size_t)rows * rows;
    const char* lg = getenv("GSI_KNN_LEGACY");
    bool legacy = lg && atoi(lg) != 0;                         // 1: the r01 scatter with global atomics (kept for before / after timings)
    for (int64_t t = 0, e = off[nu]; t < e && !legacy; ++t) {  // the item-stationary kernel counts in quarter units: half-star grid only
        const float r2 = ratings[t] * 2.f;
        if (!(r2 >= 0.f && r2 <= 255.f) || r2 != (float)(int)r2) legacy = true;
    }
    if ((rc = ws.k_num.ensure(ctx, nn * 4)) != GSI_OK) return rc;
    if (legacy) {
        if ((rc = ws.k_cnt.ensure(ctx, nn * 4)) != GSI_OK) return rc;
        if ((rc = ws.k_S.ensure(ctx, nn * 4)) != GSI_OK) return rc;
    }
    if ((rc = ws.k_ecnt.ensure(ctx, (size_t)(rows + 1) * 8)) != GSI_OK) return rc;
    if ((rc = ws.k_eoff.ensure(ctx, (size_t)(rows + 1) * 8)) != GSI_OK) return rc;
    if ((rc = ws.h_small.ensure(ctx, 64)) != GSI_OK) return rc;
    cudaStream_t st = ctx->stream;
    double* Wd = nullptr;
    if (install_weights) {
        drop_weights(ctx);
        GSI_CUDA(ctx, cudaMalloc((void**)&ctx->d_w, nn * sizeof(double)));
        ctx->own_w = true; ctx->w_rows = rows;
        if (legacy) GSI_CUDA(ctx, cudaMemsetAsync(ctx->d_w, 0, nn * sizeof(double), st));
        Wd = ctx->d_w;
    }
    const int64_t nnz = off[nu];
    GsiSpan sp(ctx, GSI_T_KNN, legacy ? 4 : 7);
    if (legacy) {
        GSI_CUDA(ctx, cudaMemsetAsync(ws.k_cnt.p, 0, nn * 4, st));
        GSI_CUDA(ctx, cudaMemsetAsync(ws.k_num.p, 0, nn * 4, st));
        GSI_CUDA(ctx, cudaMemsetAsync(ws.k_S.p, 0, nn * 4, st));
        if (nu > 0 && nmax > 1)
            knn_accumulate_kernel<<<dim3((unsigned)nu, (nmax + 127) / 128), 128, 0, st>>>(ws.k_off.as<int64_t>(), ws.k_items.as<int32_t>(), ws.k_rat.as<float>(),
                                                                                   rows, ws.k_cnt.as<int>(), ws.k_num.as<float>(), ws.k_S.as<float>());
        knn_finalize_kernel<<<rows, 256, 0, st>>>(rows, ws.k_cnt.as<int>(), ws.k_num.as<float>(), ws.k_S.as<float>(), 0, ws.k_ecnt.as<int64_t>(),
                                                   nullptr, nullptr, nullptr, nullptr, nullptr);
    } else {
        // item-major transpose + per-(user, tile) slice starts, then one CTA per (item, column tile): kern_knn.cuh
        const int ntile = (rows + KNN_CT - 1) / KNN_CT;
        if (ntile + 1 > 128) return gsi_fail(ctx, GSI_ERR_INVALID, "gsi_knn_build_host: %d items exceed the tile table", rows);
        if ((rc = ws.k_coff.ensure(ctx, (size_t)(rows + 1) * 8)) != GSI_OK) return rc;
        if ((rc = ws.k_cur.ensure(ctx, (size_t)(rows + 1) * 8)) != GSI_OK) return rc;
        if ((rc = ws.k_cuser.ensure(ctx, (size_t)std::max<int64_t>(nnz, 1) * 4)) != GSI_OK) return rc;
        if ((rc = ws.k_crat.ensure(ctx, (size_t)std::max<int64_t>(nnz, 1) * 4)) != GSI_OK) return rc;
        if ((rc = ws.k_useg.ensure(ctx, (size_t)std::max<int64_t>(nu, 1) * (ntile + 1) * 4)) != GSI_OK) return rc;
        GSI_CUDA(ctx, cudaMemsetAsync(ws.k_cur.p, 0, (size_t)(rows + 1) * 8, st));
        if (nnz > 0) knn_csc_count_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(nnz, ws.k_items.as<int32_t>(), rows, ws.k_cur.as<unsigned long long>());
        knn_scan_kernel<<<1, 1024, 0, st>>>(rows, ws.k_cur.as<int64_t>(), ws.k_coff.as<int64_t>());
        GSI_CUDA(ctx, cudaMemsetAsync(ws.k_cur.p, 0, (size_t)(rows + 1) * 8, st));
        if (nu > 0)
            knn_csc_fill_kernel<<<dim3((unsigned)nu, std::max(1, (nmax + 127) / 128)), 128, 0, st>>>(
                ws.k_off.as<int64_t>(), ws.k_items.as<int32_t>(), ws.k_rat.as<float>(), rows, ntile, ws.k_coff.as<int64_t>(),
                ws.k_cur.as<unsigned long long>(), ws.k_cuser.as<int32_t>(), ws.k_crat.as<float>(), ws.k_useg.as<int32_t>());
        knn_row_kernel<<<dim3(rows, ntile), 256, 0, st>>>(ws.k_off.as<int64_t>(), ws.k_items.as<int32_t>(), ws.k_rat.as<float>(), ws.k_coff.as<int64_t>(),
                                                          ws.k_cuser.as<int32_t>(), ws.k_crat.as<float>(), ws.k_useg.as<int32_t>(), rows, ntile,
                                                          ws.k_num.as<float>(), Wd);
        knn_compact_kernel<<<rows, 256, 0, st>>>(rows, ws.k_num.as<float>(), 0, ws.k_ecnt.as<int64_t>(), nullptr, nullptr, nullptr, nullptr);
    }
    knn_scan_kernel<<<1, 1024, 0, st>>>(rows, ws.k_ecnt.as<int64_t>(), ws.k_eoff.as<int64_t>());
    GSI_CUDA(ctx, cudaGetLastError());
    int64_t* h_ne = (int64_t*)(ws.h_small.as<char>() + 48);
    GSI_CUDA(ctx, cudaMemcpyAsync(h_ne, ws.k_eoff.as<int64_t>() + rows, 8, cudaMemcpyDeviceToHost, st));
    GSI_CUDA(ctx, cudaStreamSynchronize(st));
    const int64_t ne = *h_ne;
    ws.knn_edges = ne;
    if ((rc = ws.k_ea.ensure(ctx, std::max<int64_t>(ne, 1) * 4)) != GSI_OK) return rc;
    if ((rc = ws.k_eb.ensure(ctx, std::max<int64_t>(ne, 1) * 4)) != GSI_OK) return rc;
    if ((rc = ws.k_ew.ensure(ctx, std::max<int64_t>(ne, 1) * 4)) != GSI_OK) return rc;
    if (legacy)
        knn_finalize_kernel<<<rows, 256, 0, st>>>(rows, ws.k_cnt.as<int>(), ws.k_num.as<float>(), ws.k_S.as<float>(), 1, ws.k_ecnt.as<int64_t>(),
                                                   ws.k_eoff.as<int64_t>(), ws.k_ea.as<int32_t>(), ws.k_eb.as<int32_t>(), ws.k_ew.as<float>(), Wd);
    else
        knn_compact_kernel<<<rows, 256, 0, st>>>(rows, ws.k_num.as<float>(), 1, ws.k_ecnt.as<int64_t>(), ws.k_eoff.as<int64_t>(),
                                                  ws.k_ea.as<int32_t>(), ws.k_eb.as<int32_t>(), ws.k_ew.as<float>());
    sp.end();
    GSI_CUDA(ctx, cudaGetLastError());
    GSI_CUDA(ctx, cudaStreamSynchronize(st));
    if (n_edges) *n_edges = ne;
    return GSI_OK;
}

extern "C" int gsi_knn_edges_host(gsi_ctx* ctx, int32_t* m1, int32_t* m2, float* w, int64_t cap) {
    if (!ctx) return gsi_fail(nullptr, GSI_ERR_INVALID, "null context");
    Workspace& ws = WS(ctx);
    const int64_t ne = ws.knn_edges;
    if (cap < ne) return gsi_fail(ctx, GSI_ERR_CAPACITY, "edge buffers too small: need %lld", (long long)ne);
    if (ne == 0) return GSI_OK;
    if (!m1 || !m2 || !w) return gsi_fail(ctx, GSI_ERR_INVALID, "gsi_knn_edges_host: null argument");
    GSI_CUDA(ctx, cudaSetDevice(ctx->device));
    GSI_CUDA(ctx, cudaMemcpyAsync(m1, ws.k_ea.p, ne * 4, cudaMemcpyDeviceToHost, ctx->stream));
    GSI_CUDA(ctx, cudaMemcpyAsync(m2, ws.k_eb.p, ne * 4, cudaMemcpyDeviceToHost, ctx->stream));
    GSI_CUDA(ctx, cudaMemcpyAsync(w, ws.k_ew.p, ne * 4, cudaMemcpyDeviceToHost, ctx->stream));
    GSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GSI_OK;
}

extern "C" int gsi_knn_corated_host(gsi_ctx* ctx, int64_t nu, const int64_t* off, const int32_t* items, int rows, uint8_t* co) {
    if (!ctx) return gsi_fail(nullptr, GSI_ERR_INVALID, "null context");
    if (nu < 0 || rows < 1 || !off || !co || (nu > 0 && !items)) return gsi_fail(ctx, GSI_ERR_INVALID, "gsi_knn_corated_host: bad argument");
    GSI_CUDA(ctx, cudaSetDevice(ctx->device));
    Workspace& ws = WS(ctx);
    int rc, nmax = 0;
    if ((rc = knn_upload_csr(ctx, nu, off, items, nullptr, &nmax)) != GSI_OK) return rc;
    const size_t nn = (size_t)rows * rows;
    if ((rc = ws.k_co.ensure(ctx, nn)) != GSI_OK) return rc;
    GSI_CUDA(ctx, cudaMemsetAsync(ws.k_co.p, 0, nn, ctx->stream));
    GsiSpan sp(ctx, GSI_T_KNN, 1);
    if (nu > 0 && nmax > 1)
        knn_corated_kernel<<<dim3((unsigned)nu, (nmax + 127) / 128), 128, 0, ctx->stream>>>(ws.k_off.as<int64_t>(), ws.k_items.as<int32_t>(), rows,
                                                                                         ws.k_co.as<unsigned char>());
    sp.end();
    GSI_CUDA(ctx, cudaGetLastError());
    GSI_CUDA(ctx, cudaMemcpyAsync(co, ws.k_co.p, nn, cudaMemcpyDeviceToHost, ctx->stream));
    GSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GSI_OK;
}

extern "C" int gsi_knn3_host(gsi_ctx* ctx, int64_t nu, const int64_t* off, const int32_t* items, const float* ratings,
                             float* movie_err_sum, int32_t* movie_cnt, uint8_t* has_edge) {
    if (!ctx) return gsi_fail(nullptr, GSI_ERR_INVALID, "null context");
    if (!ctx->d_w) return gsi_fail(ctx, GSI_ERR_STATE, "gsi_knn3: no weight table set");
    if (nu < 0 || !off || !movie_err_sum || !movie_cnt || !has_edge || (nu > 0 && (!items || !ratings)))
        return gsi_fail(ctx, GSI_ERR_INVALID, "gsi_knn3_host: bad argument");
    GSI_CUDA(ctx, cudaSetDevice(ctx->device));
    Workspace& ws = WS(ctx);
    const int rows = ctx->w_rows;
    int rc, nmax = 0;
    if ((rc = knn_upload_csr(ctx, nu, off, items, ratings, &nmax)) != GSI_OK) return rc;
    if ((rc = ws.k_err.ensure(ctx, (size_t)rows * 4)) != GSI_OK) return rc;
    if ((rc = ws.k_mcnt.ensure(ctx, (size_t)rows * 4)) != GSI_OK) return rc;
    if ((rc = ws.k_has.ensure(ctx, (size_t)rows)) != GSI_OK) return rc;
    cudaStream_t st = ctx->stream;
    GSI_CUDA(ctx, cudaMemsetAsync(ws.k_err.p, 0, (size_t)rows * 4, st));
    GSI_CUDA(ctx, cudaMemsetAsync(ws.k_mcnt.p, 0, (size_t)rows * 4, st));
    GSI_CUDA(ctx, cudaMemsetAsync(ws.k_has.p, 0, (size_t)rows, st));
    GsiSpan sp(ctx, GSI_T_KNN, 2);
    if (nu > 0 && nmax > 0)
        knn3_kernel<<<dim3((unsigned)nu, (nmax + 127) / 128), 128, 0, st>>>(ws.k_off.as<int64_t>(), ws.k_items.as<int32_t>(), ws.k_rat.as<float>(),
                                                                     ctx->d_w, rows, ws.k_err.as<float>(), ws.k_mcnt.as<int>());
    knn3_vertices_kernel<<<rows, 256, 0, st>>>(ctx->d_w, rows, ws.k_has.as<unsigned char>());
    sp.end();
    GSI_CUDA(ctx, cudaGetLastError());
    GSI_CUDA(ctx, cudaMemcpyAsync(movie_err_sum, ws.k_err.p, (size_t)rows * 4, cudaMemcpyDeviceToHost, st));
    GSI_CUDA(ctx, cudaMemcpyAsync(movie_cnt, ws.k_mcnt.p, (size_t)rows * 4, cudaMemcpyDeviceToHost, st));
    GSI_CUDA(ctx, cudaMemcpyAsync(has_edge, ws.k_has.p, (size_t)rows, cudaMemcpyDeviceToHost, st));
    GSI_CUDA(ctx, cudaStreamSynchronize(st));
    return GSI_OK;
}

// ---- Chebyshev graph filter (cheby.cpp) -----------------------------------------------------------
extern "C" int gsi_cheby_filter_host(gsi_ctx* ctx, int64_t nv, const int64_t* row_off, const int32_t* col, const double* w,
                                     const double* x, int ncoef, const double* coef, double* y) {
    if (!ctx) return gsi_fail(nullptr, GSI_ERR_INVALID, "null context");
    if (nv < 0 || nv > INT32_MAX || !row_off || ncoef < 2 || !coef || (nv > 0 && (!x || !y)))
        return gsi_fail(ctx, GSI_ERR_INVALID, "gsi_cheby_filter_host: need nv >= 0, ncoef >= 2 and non-null arrays");
    if (nv == 0) return GSI_OK;
    const int64_t nnz = row_off[nv];
    if (row_off[0] != 0 || nnz < 0 || (nnz > 0 && (!col || !w))) return gsi_fail(ctx, GSI_ERR_INVALID, "gsi_cheby_filter_host: bad CSR");
    for (int64_t i = 0; i < nv; ++i)
        if (row_off[i + 1] < row_off[i]) return gsi_fail(ctx, GSI_ERR_INVALID, "gsi_cheby_filter_host: row offsets must not decrease");
    for (int64_t e = 0; e < nnz; ++e)
        if (col[e] < 0 || col[e] >= nv) return gsi_fail(ctx, GSI_ERR_INVALID, "gsi_cheby_filter_host: column %d outside 0..%lld", col[e], (long long)nv - 1);
    GSI_CUDA(ctx, cudaSetDevice(ctx->device));
    Workspace& ws = WS(ctx);
    cudaStream_t st = ctx->stream;
    int rc;
    if ((rc = ws.c_off.ensure(ctx, (nv + 1) * 8)) != GSI_OK) return rc;
    if ((rc = ws.c_col.ensure(ctx, std::max<int64_t>(nnz, 1) * 4)) != GSI_OK) return rc;
    if ((rc = ws.c_w.ensure(ctx, std::max<int64_t>(nnz, 1) * 8)) != GSI_OK) return rc;
    if ((rc = ws.c_wn.ensure(ctx, std::max<int64_t>(nnz, 1) * 8)) != GSI_OK) return rc;
    if ((rc = ws.c_vec.ensure(ctx, 5 * nv * 8)) != GSI_OK) return rc;
    GSI_CUDA(ctx, cudaMemcpyAsync(ws.c_off.p, row_off, (nv + 1) * 8, cudaMemcpyHostToDevice, st));
    if (nnz) {
        GSI_CUDA(ctx, cudaMemcpyAsync(ws.c_col.p, col, nnz * 4, cudaMemcpyHostToDevice, st));
        GSI_CUDA(ctx, cudaMemcpyAsync(ws.c_w.p, w, nnz * 8, cudaMemcpyHostToDevice, st));
    }
    double* deg = ws.c_vec.as<double>();
    double* t[3] = {deg + nv, deg + 2 * nv, deg + 3 * nv};
    double* yd = deg + 4 * nv;
    GSI_CUDA(ctx, cudaMemcpyAsync(t[0], x, nv * 8, cudaMemcpyHostToDevice, st));
    const int64_t* d_off = ws.c_off.as<int64_t>();
    const int32_t* d_col = ws.c_col.as<int32_t>();
    const unsigned grid = (unsigned)((nv * 32 + 255) / 256);
    {
        GsiSpan sp(ctx, GSI_T_CHEBY, ncoef + 1);
        cheby_degree_kernel<<<grid, 256, 0, st>>>((int)nv, d_off, ws.c_w.as<double>(), deg);
        cheby_normalise_kernel<<<grid, 256, 0, st>>>((int)nv, d_off, d_col, ws.c_w.as<double>(), deg, ws.c_wn.as<double>());
        // superstep 1: t[0] = T0 = x, t[1] = T1;  superstep k: T_{k+1} from T_k (cur) and T_{k-1} (old), rotating buffers
        cheby_step_kernel<<<grid, 256, 0, st>>>((int)nv, d_off, d_col, ws.c_wn.as<double>(), t[0], t[0], t[1], yd, coef[0], coef[1], 1);
        int old = 0, cur = 1, nxt = 2;
        for (int k = 2; k < ncoef; ++k) {
            cheby_step_kernel<<<grid, 256, 0, st>>>((int)nv, d_off, d_col, ws.c_wn.as<double>(), t[cur], t[old], t[nxt], yd, coef[k], 0.0, 0);
            const int tmp = old; old = cur; cur = nxt; nxt = tmp;
        }
        sp.end();
        GSI_CUDA(ctx, cudaGetLastError());
    }
    GSI_CUDA(ctx, cudaMemcpyAsync(y, yd, nv * 8, cudaMemcpyDeviceToHost, st));
    GSI_CUDA(ctx, cudaStreamSynchronize(st));
    return GSI_OK;
}

// ---- measurement ---------------------------------------------------------------------------------
extern "C" int gsi_timing_enable(gsi_ctx* ctx, int on) { if (!ctx) return GSI_ERR_INVALID; ctx->timing = on != 0; return GSI_OK; }
extern "C" int gsi_timing_reset(gsi_ctx* ctx) {
    if (!ctx) return GSI_ERR_INVALID;
    cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    drain_spans(ctx);
    for (int i = 0; i < GSI_T_COUNT; ++i) { ctx->t_ms[i] = 0; ctx->t_launch[i] = 0; ctx->t_samples[i] = 0; }
    return GSI_OK;
}
extern "C" int gsi_timing_get(gsi_ctx* ctx, double* ms, int64_t* launches, int64_t* samples) {
    if (!ctx) return GSI_ERR_INVALID;
    GSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    drain_spans(ctx);
    for (int i = 0; i < GSI_T_COUNT; ++i) {
        if (ms) ms[i] = ctx->t_ms[i];
        if (launches) launches[i] = ctx->t_launch[i];
        if (samples) samples[i] = ctx->t_samples[i];
    }
    return GSI_OK;
}
extern "C" int gsi_measure_fp64_tflops(gsi_ctx* ctx, int use_dmma, double* tflops) {
    if (!ctx || !tflops) return GSI_ERR_INVALID;
    GSI_CUDA(ctx, cudaSetDevice(ctx->device));
    Workspace& ws = WS(ctx);
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 1 << 14;
    int rc;
    if ((rc = ws.probe.ensure(ctx, (size_t)blocks * threads * 8)) != GSI_OK) return rc;
    cudaEvent_t a, b;
    GSI_CUDA(ctx, cudaEventCreate(&a));
    GSI_CUDA(ctx, cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        GSI_CUDA(ctx, cudaEventRecord(a, ctx->stream));
        if (use_dmma) fp64_dmma_probe<<<blocks, threads, 0, ctx->stream>>>(ws.probe.as<double>(), iters);
        else fp64_fma_probe<<<blocks, threads, 0, ctx->stream>>>(ws.probe.as<double>(), iters);
        GSI_CUDA(ctx, cudaEventRecord(b, ctx->stream));
        GSI_CUDA(ctx, cudaEventSynchronize(b));
        float ms;
        GSI_CUDA(ctx, cudaEventElapsedTime(&ms, a, b));
        if (rep > 0) best = std::min(best, ms);
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    // fma probe: 8 FMA per thread per iter; dmma probe: 8 mma (8x8x4 = 256 FMA per warp) per iter
    const double flop = use_dmma ? (double)blocks * (threads / 32) * iters * 8.0 * 512.0
                                 : (double)blocks * threads * iters * 8.0 * 2.0;
    *tflops = flop / (best * 1e-3) / 1e12;
    return GSI_OK;
}

#include "group_host.cuh"
