// Device group: every GPU of the box behind one call (include/gsi.h, "device group").  The reference's tool spreads the
// users over all workers of the box by itself (precompute_local_threads.cpp:300-314: one task per user on a thread pool);
// here a group owns one context per device and runs one host thread per device while a call is in flight.  Users are
// dealt by longest-processing-time-first on n^3 + 64 n^2 (the same rule as shard.py: deterministic, no data-path
// collective); the weight table is replicated once, device to device.  Included by gsi.cu after the single-context API.
#pragma once
#include <dlfcn.h>

#include <mutex>
#include <queue>
#include <thread>

struct gsi_group {
    std::vector<gsi_ctx*> ctx;
    std::vector<int> dev;
    std::string err;
    const char* bcast = "single";
    // NCCL, resolved at run time (no link-time dependency: the tools run where only the driver and this library exist)
    void* nccl_lib = nullptr;
    std::vector<void*> comms;
};

static int group_fail(gsi_group* g, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (g) g->err = buf;
    g_tls_err = buf;
    return code;
}

extern "C" int gsi_group_create(gsi_group** out, int n_devices, const int* devices) {
    if (!out) return gsi_fail(nullptr, GSI_ERR_INVALID, "gsi_group_create: out is null");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return gsi_fail(nullptr, GSI_ERR_CUDA, "gsi_group_create: no CUDA device (%s); libgsi has no CPU fallback", cudaGetErrorString(e));
    gsi_group* g = new gsi_group();
    if (n_devices <= 0 || !devices) for (int d = 0; d < ndev; ++d) g->dev.push_back(d);
    else g->dev.assign(devices, devices + n_devices);
    for (size_t i = 0; i < g->dev.size(); ++i) {
        // (the same ordinal twice is refused: two members would queue behind each other on one device.  GSI_GROUP_ALLOW_DUP=1
        // lifts that for tests of the dealing / merging logic on a single-GPU box; NCCL cannot span it, peer copy is used)
        for (size_t j = 0; j < i && !getenv("GSI_GROUP_ALLOW_DUP"); ++j)
            if (g->dev[j] == g->dev[i]) { delete g; return gsi_fail(nullptr, GSI_ERR_INVALID, "gsi_group_create: device %d listed twice", g->dev[i]); }
        gsi_ctx* c = nullptr;
        const int rc = gsi_create(&c, g->dev[i], nullptr);
        if (rc != GSI_OK) {
            for (gsi_ctx* o : g->ctx) gsi_destroy(o);
            delete g;
            return rc;                                     // g_tls_err holds gsi_create's message
        }
        g->ctx.push_back(c);
    }
    *out = g;
    return GSI_OK;
}

// ---- NCCL through dlopen -------------------------------------------------------------------------------------------
typedef int (*nccl_comm_init_all_t)(void** comms, int ndev, const int* devlist);
typedef int (*nccl_comm_destroy_t)(void* comm);
typedef int (*nccl_group_t)(void);
typedef int (*nccl_broadcast_t)(const void* send, void* recv, size_t count, int dtype, int root, void* comm, cudaStream_t st);
#define GSI_NCCL_DOUBLE 8       // ncclFloat64 (nccl.h: ncclDataType_t)

static void group_drop_nccl(gsi_group* g) {
    if (g->nccl_lib) {
        nccl_comm_destroy_t destroy = (nccl_comm_destroy_t)dlsym(g->nccl_lib, "ncclCommDestroy");
        if (destroy) for (void* c : g->comms) if (c) destroy(c);
        g->comms.clear();
        dlclose(g->nccl_lib);
        g->nccl_lib = nullptr;
    }
}

extern "C" int gsi_group_destroy(gsi_group* g) {
    if (!g) return GSI_OK;
    group_drop_nccl(g);
    for (gsi_ctx* c : g->ctx) gsi_destroy(c);
    delete g;
    return GSI_OK;
}
extern "C" int gsi_group_size(const gsi_group* g) { return g ? (int)g->ctx.size() : 0; }
extern "C" gsi_ctx* gsi_group_ctx(gsi_group* g, int i) { return (g && i >= 0 && i < (int)g->ctx.size()) ? g->ctx[i] : nullptr; }
extern "C" const char* gsi_group_last_error(const gsi_group* g) { return g ? g->err.c_str() : g_tls_err.c_str(); }
extern "C" const char* gsi_group_broadcast_path(const gsi_group* g) { return g ? g->bcast : ""; }
extern "C" int gsi_group_set_workspace_limit(gsi_group* g, int64_t bytes) {
    if (!g) return gsi_fail(nullptr, GSI_ERR_INVALID, "null group");
    for (gsi_ctx* c : g->ctx) {
        const int rc = gsi_set_workspace_limit(c, bytes);
        if (rc != GSI_OK) return group_fail(g, rc, "%s", gsi_last_error(c));
    }
    return GSI_OK;
}

// W lives on member 0; replicate it.  NCCL broadcast (one communicator per device, group call from this thread) when the
// library loads and initialises; otherwise peer copies along a binomial tree (0 -> 1, {0,1} -> {2,3}, ...), every copy
// device to device over NVLink / NVSwitch.  GSI_GROUP_BCAST=peer|nccl forces one of them.
static int group_replicate_weights(gsi_group* g, int rows) {
    const int nd = (int)g->ctx.size();
    const size_t count = (size_t)rows * rows;
    gsi_ctx* c0 = g->ctx[0];
    for (int i = 1; i < nd; ++i) {
        gsi_ctx* c = g->ctx[i];
        GSI_CUDA(c, cudaSetDevice(c->device));
        drop_weights(c);
        GSI_CUDA(c, cudaMalloc((void**)&c->d_w, count * sizeof(double)));
        c->own_w = true; c->w_rows = rows;
    }
    const char* force = getenv("GSI_GROUP_BCAST");
    bool use_nccl = !(force && strcmp(force, "peer") == 0) && !getenv("GSI_GROUP_ALLOW_DUP");
    if (use_nccl && !g->nccl_lib) {
        g->nccl_lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (g->nccl_lib) {
            nccl_comm_init_all_t init = (nccl_comm_init_all_t)dlsym(g->nccl_lib, "ncclCommInitAll");
            g->comms.assign(nd, nullptr);
            if (!init || init(g->comms.data(), nd, g->dev.data()) != 0) group_drop_nccl(g);
        }
    }
    if (use_nccl && g->nccl_lib) {
        nccl_group_t gs = (nccl_group_t)dlsym(g->nccl_lib, "ncclGroupStart"), ge = (nccl_group_t)dlsym(g->nccl_lib, "ncclGroupEnd");
        nccl_broadcast_t bc = (nccl_broadcast_t)dlsym(g->nccl_lib, "ncclBroadcast");
        if (gs && ge && bc) {
            int bad = gs();
            for (int i = 0; i < nd && !bad; ++i) {
                cudaSetDevice(g->dev[i]);
                bad = bc(c0->d_w, g->ctx[i]->d_w, count, GSI_NCCL_DOUBLE, 0, g->comms[i], g->ctx[i]->stream);
            }
            bad = ge() || bad;
            for (int i = 0; i < nd; ++i) { cudaSetDevice(g->dev[i]); if (cudaStreamSynchronize(g->ctx[i]->stream) != cudaSuccess) bad = 1; }
            if (!bad) { g->bcast = "nccl"; return GSI_OK; }
            if (force && strcmp(force, "nccl") == 0) return group_fail(g, GSI_ERR_CUDA, "ncclBroadcast of the weight table failed");
            cudaGetLastError();
        }
    } else if (force && strcmp(force, "nccl") == 0) {
        return group_fail(g, GSI_ERR_CUDA, "GSI_GROUP_BCAST=nccl but libnccl.so.2 could not be loaded / initialised");
    }
    // binomial tree of peer copies
    for (int i = 0; i < nd; ++i)
        for (int j = 0; j < nd; ++j)
            if (i != j) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, g->dev[i], g->dev[j]);
                if (can) { cudaSetDevice(g->dev[i]); if (cudaDeviceEnablePeerAccess(g->dev[j], 0) != cudaSuccess) cudaGetLastError(); }
            }
    for (int have = 1; have < nd; have *= 2) {
        for (int s = 0; s < have && s + have < nd; ++s) {
            gsi_ctx* src = g->ctx[s];
            gsi_ctx* dst = g->ctx[s + have];
            GSI_CUDA(dst, cudaSetDevice(dst->device));
            GSI_CUDA(dst, cudaMemcpyPeerAsync(dst->d_w, dst->device, src->d_w, src->device, count * sizeof(double), dst->stream));
        }
        for (int s = 0; s < have && s + have < nd; ++s) {
            gsi_ctx* dst = g->ctx[s + have];
            GSI_CUDA(dst, cudaSetDevice(dst->device));
            GSI_CUDA(dst, cudaStreamSynchronize(dst->stream));
        }
    }
    g->bcast = "peer";
    return GSI_OK;
}

extern "C" int gsi_group_set_weights_host(gsi_group* g, const double* w, int rows) {
    if (!g || g->ctx.empty()) return gsi_fail(nullptr, GSI_ERR_INVALID, "null group");
    int rc = gsi_set_weights_host(g->ctx[0], w, rows);
    if (rc != GSI_OK) return group_fail(g, rc, "%s", gsi_last_error(g->ctx[0]));
    g->bcast = "single";
    if (g->ctx.size() == 1) return GSI_OK;
    rc = group_replicate_weights(g, rows);
    if (rc != GSI_OK && g->err.empty()) g->err = g_tls_err;
    return rc;
}

// longest-processing-time-first (shard.py lpt_assign): owner[u] in [0, nd), ties by index, deterministic
static void group_lpt(const std::vector<double>& cost, int nd, std::vector<int>& owner) {
    const int64_t nu = (int64_t)cost.size();
    std::vector<int64_t> order(nu);
    std::iota(order.begin(), order.end(), (int64_t)0);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return cost[a] > cost[b]; });
    typedef std::pair<double, int> Load;                   // (load, shard): smallest load first, then smallest shard
    std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
    for (int s = 0; s < nd; ++s) heap.push({0.0, s});
    owner.assign(nu, 0);
    for (int64_t u : order) {
        Load l = heap.top(); heap.pop();
        owner[u] = l.second;
        heap.push({l.first + cost[u], l.second});
    }
}

struct GroupSinkState {
    std::mutex* mu; gsi_record_sink sink; void* opaque;
    const std::vector<int64_t>* global_of;     // shard-local user index -> caller's user index
    const int64_t* goff; const int64_t* loff;  // caller's offsets, the shard's offsets
    double* sig_all;                           // [nnz of the caller's CSR]
    std::vector<int64_t> remap;
};
static int group_sink(void* opaque, const gsi_record_chunk* ch) {
    GroupSinkState* S = (GroupSinkState*)opaque;
    S->remap.resize(ch->n_records);
    for (int64_t j = 0; j < ch->n_records; ++j) {
        const int64_t lu = ch->user_index[j], gu = (*S->global_of)[lu];
        S->remap[j] = gu;
        memcpy(S->sig_all + S->goff[gu], ch->sig_min + S->loff[lu], (size_t)ch->n[j] * sizeof(double));   // disjoint ranges per user
    }
    gsi_record_chunk out = *ch;
    out.user_index = S->remap.data();
    out.sig_min = S->sig_all;
    std::lock_guard<std::mutex> lock(*S->mu);
    return S->sink(S->opaque, &out);
}

extern "C" int gsi_group_precompute_stream(gsi_group* g, int64_t nu, const int64_t* offsets, const int32_t* items,
                                           gsi_record_sink sink, void* opaque) {
    if (!g || g->ctx.empty()) return gsi_fail(nullptr, GSI_ERR_INVALID, "null group");
    const int nd = (int)g->ctx.size();
    if (nd == 1) {
        const int rc = gsi_precompute_stream(g->ctx[0], nu, offsets, items, sink, opaque);
        if (rc != GSI_OK) g->err = gsi_last_error(g->ctx[0]);
        return rc;
    }
    if (nu < 0 || !offsets || !sink || (nu > 0 && !items)) return group_fail(g, GSI_ERR_INVALID, "gsi_group_precompute_stream: null argument");
    if (nu == 0) return GSI_OK;
    if (offsets[0] != 0) return group_fail(g, GSI_ERR_INVALID, "offsets[0] must be 0");
    std::vector<double> cost(nu);
    for (int64_t u = 0; u < nu; ++u) {
        const double n = (double)(offsets[u + 1] - offsets[u]);
        if (n < 1) return group_fail(g, GSI_ERR_INVALID, "user %lld has no rated movies", (long long)u);
        cost[u] = n * n * n + 64.0 * n * n;
    }
    std::vector<int> owner;
    group_lpt(cost, nd, owner);
    std::vector<std::vector<int64_t>> global_of(nd), loff(nd);
    std::vector<std::vector<int32_t>> litems(nd);
    for (int d = 0; d < nd; ++d) loff[d].push_back(0);
    for (int64_t u = 0; u < nu; ++u) {
        const int d = owner[u];
        global_of[d].push_back(u);
        litems[d].insert(litems[d].end(), items + offsets[u], items + offsets[u + 1]);
        loff[d].push_back((int64_t)litems[d].size());
    }
    std::vector<double> sig_all((size_t)offsets[nu]);
    std::mutex mu;
    std::vector<int> rcs(nd, GSI_OK);
    std::vector<GroupSinkState> st(nd);
    std::vector<std::thread> pool;
    for (int d = 0; d < nd; ++d) {
        st[d] = GroupSinkState{&mu, sink, opaque, &global_of[d], offsets, loff[d].data(), sig_all.data(), {}};
        pool.emplace_back([&, d]() {
            if (global_of[d].empty()) return;
            rcs[d] = gsi_precompute_stream(g->ctx[d], (int64_t)global_of[d].size(), loff[d].data(), litems[d].data(), group_sink, &st[d]);
        });
        if (getenv("GSI_GROUP_ALLOW_DUP")) pool.back().join();     // test mode (members share a device): one member at a time
    }
    for (auto& t : pool) if (t.joinable()) t.join();
    for (int d = 0; d < nd; ++d)
        if (rcs[d] != GSI_OK) return group_fail(g, rcs[d], "device %d: %s", g->dev[d], gsi_last_error(g->ctx[d]));
    return GSI_OK;
}

extern "C" int gsi_group_predict_host(gsi_group* g, int64_t nu, const int64_t* offsets, const int32_t* items, const double* w_lim,
                                      const double* ratings, const int32_t* k, const int64_t* lam_off, const int64_t* vec_off,
                                      const double* lam, int64_t lam_len, const double* vec, int64_t vec_len, const uint8_t* pair_mask,
                                      float* err, int32_t* kk, double* pred, int32_t* status, int32_t* cols) {
    if (!g || g->ctx.empty()) return gsi_fail(nullptr, GSI_ERR_INVALID, "null group");
    const int nd = (int)g->ctx.size();
    if (nd == 1) {
        const int rc = gsi_predict_host(g->ctx[0], nu, offsets, items, w_lim, ratings, k, lam_off, vec_off, lam, lam_len, vec, vec_len,
                                        pair_mask, err, kk, pred, status, cols);
        if (rc != GSI_OK) g->err = gsi_last_error(g->ctx[0]);
        return rc;
    }
    if (nu < 0 || !offsets || (nu > 0 && (!items || !w_lim || !ratings || !k || !lam_off || !vec_off || !lam || !vec || !err || !kk || !pred || !status || !cols)))
        return group_fail(g, GSI_ERR_INVALID, "gsi_group_predict_host: null argument");
    if (nu == 0) return GSI_OK;
    std::vector<double> cost(nu);
    for (int64_t u = 0; u < nu; ++u) {
        const int64_t n = offsets[u + 1] - offsets[u];
        if (n < 1 || k[u] < 1) return group_fail(g, GSI_ERR_INVALID, "user %lld: empty record", (long long)u);
        if (lam_off[u] < 0 || lam_off[u] + k[u] > lam_len || vec_off[u] < 0 || vec_off[u] + n * k[u] > vec_len)
            return group_fail(g, GSI_ERR_INVALID, "user %lld: record offsets outside lam/vec", (long long)u);
        int64_t pairs = n;
        if (pair_mask) { pairs = 0; for (int64_t t = offsets[u]; t < offsets[u + 1]; ++t) pairs += pair_mask[t] != 0; }
        cost[u] = (double)pairs * (double)n * (double)k[u] * (double)k[u] + 1.0;
    }
    std::vector<int> owner;
    group_lpt(cost, nd, owner);
    struct Part {
        std::vector<int64_t> global_of, off, lamoff, vecoff;
        std::vector<int32_t> items, k, kk, status, cols;
        std::vector<double> wlim, rat, lam, vec, pred;
        std::vector<uint8_t> mask;
        std::vector<float> err;
    };
    std::vector<Part> parts(nd);
    for (int d = 0; d < nd; ++d) parts[d].off.push_back(0);
    for (int64_t u = 0; u < nu; ++u) {
        Part& P = parts[owner[u]];
        const int64_t b = offsets[u], e = offsets[u + 1], n = e - b;
        P.global_of.push_back(u);
        P.items.insert(P.items.end(), items + b, items + e);
        P.wlim.insert(P.wlim.end(), w_lim + b, w_lim + e);
        P.rat.insert(P.rat.end(), ratings + b, ratings + e);
        if (pair_mask) P.mask.insert(P.mask.end(), pair_mask + b, pair_mask + e);
        P.off.push_back((int64_t)P.items.size());
        P.k.push_back(k[u]);
        P.lamoff.push_back((int64_t)P.lam.size());
        P.lam.insert(P.lam.end(), lam + lam_off[u], lam + lam_off[u] + k[u]);
        P.vecoff.push_back((int64_t)P.vec.size());
        P.vec.insert(P.vec.end(), vec + vec_off[u], vec + vec_off[u] + n * k[u]);
    }
    std::vector<int> rcs(nd, GSI_OK);
    std::vector<std::thread> pool;
    for (int d = 0; d < nd; ++d) {
        pool.emplace_back([&, d]() {
            Part& P = parts[d];
            if (P.global_of.empty()) return;
            const size_t nnz = P.items.size();
            P.err.resize(nnz); P.kk.resize(nnz); P.pred.resize(nnz); P.status.resize(nnz); P.cols.resize(nnz);
            rcs[d] = gsi_predict_host(g->ctx[d], (int64_t)P.global_of.size(), P.off.data(), P.items.data(), P.wlim.data(), P.rat.data(),
                                      P.k.data(), P.lamoff.data(), P.vecoff.data(), P.lam.data(), (int64_t)P.lam.size(), P.vec.data(),
                                      (int64_t)P.vec.size(), pair_mask ? P.mask.data() : nullptr, P.err.data(), P.kk.data(),
                                      P.pred.data(), P.status.data(), P.cols.data());
        });
        if (getenv("GSI_GROUP_ALLOW_DUP")) pool.back().join();
    }
    for (auto& t : pool) if (t.joinable()) t.join();
    for (int d = 0; d < nd; ++d)
        if (rcs[d] != GSI_OK) return group_fail(g, rcs[d], "device %d: %s", g->dev[d], gsi_last_error(g->ctx[d]));
    for (int d = 0; d < nd; ++d) {
        const Part& P = parts[d];
        for (size_t j = 0; j < P.global_of.size(); ++j) {
            const int64_t u = P.global_of[j], b = offsets[u], n = offsets[u + 1] - b, lb = P.off[j];
            memcpy(err + b, P.err.data() + lb, (size_t)n * sizeof(float));
            memcpy(kk + b, P.kk.data() + lb, (size_t)n * sizeof(int32_t));
            memcpy(pred + b, P.pred.data() + lb, (size_t)n * sizeof(double));
            memcpy(status + b, P.status.data() + lb, (size_t)n * sizeof(int32_t));
            memcpy(cols + b, P.cols.data() + lb, (size_t)n * sizeof(int32_t));
        }
    }
    return GSI_OK;
}
