// Shared body of precompute_local / precompute_local_threads: same cwd-relative inputs and the same
// out_eigen_ record layout as the reference (precompute_local.cpp:84-282,
// precompute_local_threads.cpp:215-317); the per-user math is one gsi_precompute_stream() call.
#pragma once
#include <atomic>
#include <thread>

#include "host_io.hpp"

namespace gsihost {

struct EigenSink {
    const Csr* csr; FILE* out; int n_threads; size_t done = 0; bool binary = false;
};

// Binary side format of the records (README.md:29 "TODO: binary output"), written to out_eigen_.bin
// instead of the text when GSI_EIGEN_BINARY=1 and read back by local_calc_precomp under the same switch.  Little endian:
//   file   = "GSIEIG01" record*
//   record = u32 user', i32 n, i32 k, i32 0, i32 movie[n] (+ i32 0 if n is odd), f64 sig_min[n], f64 lambda[k], f64 U[n*k] (row-major)
// Same fields and order as the text record (precompute_local.cpp:265-278), but the doubles keep all 53 bits (the text keeps 6 digits).
static const char kEigenBinMagic[8] = {'G', 'S', 'I', 'E', 'I', 'G', '0', '1'};
inline void format_record_bin(const Csr& csr, const gsi_record_chunk* ch, int64_t j, std::string& s) {
    const int64_t u = ch->user_index[j];
    const int32_t n = ch->n[j], k = ch->k[j];
    const uint32_t uid = csr.users[u];
    const int32_t head[4] = {(int32_t)uid, n, k, 0};
    s.append((const char*)head, sizeof head);
    s.append((const char*)(csr.items.data() + csr.offsets[u]), (size_t)n * 4);
    if (n & 1) { const int32_t z = 0; s.append((const char*)&z, 4); }
    s.append((const char*)(ch->sig_min + csr.offsets[u]), (size_t)n * 8);
    s.append((const char*)(ch->lam + ch->lam_off[j]), (size_t)k * 8);
    s.append((const char*)(ch->vec + ch->vec_off[j]), (size_t)n * k * 8);
}

// one out_eigen_ record (precompute_local.cpp:265-278): every token followed by a space
inline void format_record(const Csr& csr, const gsi_record_chunk* ch, int64_t j, std::string& s) {
    const int64_t u = ch->user_index[j];
    const int n = ch->n[j], k = ch->k[j];
    s.reserve(s.size() + 16 * (size_t)n + 10 * (size_t)k + 9 * (size_t)n * k + 64);
    append_int(s, csr.users[u]); append_int(s, n); append_int(s, k);
    const double* sig = ch->sig_min + csr.offsets[u];
    for (int i = 0; i < n; ++i) { append_int(s, csr.items[csr.offsets[u] + i]); append_g(s, sig[i]); }
    s.push_back('\n');
    const double* lam = ch->lam + ch->lam_off[j];
    for (int i = 0; i < k; ++i) append_g(s, lam[i]);
    s.push_back('\n');
    const double* vec = ch->vec + ch->vec_off[j];
    for (int64_t t = 0; t < (int64_t)n * k; ++t) append_g(s, vec[t]);
    s.push_back('\n');
}

// the analogue of save_output() (precompute_local_threads.cpp:89-98): records are appended as the
// chunks complete; formatting is spread over n_threads host threads, file order = processing order
inline int eigen_sink(void* opaque, const gsi_record_chunk* ch) {
    EigenSink* S = (EigenSink*)opaque;
    const int64_t nr = ch->n_records;
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(S->n_threads, nr));
    std::vector<std::string> parts(nt);
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t)
        pool.emplace_back([&, t]() {
            const int64_t b = nr * t / nt, e = nr * (t + 1) / nt;
            for (int64_t j = b; j < e; ++j) S->binary ? format_record_bin(*S->csr, ch, j, parts[t]) : format_record(*S->csr, ch, j, parts[t]);
        });
    for (auto& th : pool) th.join();
    for (auto& p : parts)
        if (!p.empty() && fwrite(p.data(), 1, p.size(), S->out) != p.size()) return 1;
    S->done += (size_t)nr;
    printf("%g%%\n", 100.0 * (double)S->done / (double)(S->csr->users.size()));
    return 0;
}

inline int precompute_main(int n_threads) {
    const std::string path = "movielens/", suffix = ".validate";
    std::vector<Triple> rows;
    for (const std::string& f : list_files(path, [&](const std::string& n) { return ends_with(n, suffix); })) {
        printf("Reading file: %s\n", f.c_str());
        read_triples(f, rows);
    }
    Csr csr = build_csr(rows, /*map_user=*/true);        // users[INT_MAX - user][movie] = rating  :107-108
    std::vector<double> table;
    int wrows = 1;
    load_weights_table(table, wrows);
    FILE* out = fopen("out_eigen_", "w");               // truncate :150-151
    if (!out) { perror("out_eigen_"); return 1; }
    const char* bin_env = getenv("GSI_EIGEN_BINARY");
    const bool binary = bin_env && atoi(bin_env) != 0;
    if (binary) {                                       // the text file stays truncated (no stale records), the records go to out_eigen_.bin
        fclose(out);
        out = fopen("out_eigen_.bin", "wb");
        if (!out) { perror("out_eigen_.bin"); return 1; }
        fwrite(kEigenBinMagic, 1, sizeof kEigenBinMagic, out);
    } else {
        remove("out_eigen_.bin");                       // a text run must not leave an older binary file behind
    }
    printf("Number of movies: %d\n", wrows - 1);
    printf("Number of users: %zu\n", csr.users.size());
    if (csr.users.empty()) { fclose(out); return 0; }
    gsi_ctx* ctx = nullptr;
    const char* dev = getenv("GSI_DEVICE");
    if (gsi_create(&ctx, dev ? atoi(dev) : 0, nullptr) != GSI_OK) { fclose(out); return fail(nullptr, "gsi_create"); }
    int rc = gsi_set_weights_host(ctx, table.data(), wrows);
    std::vector<double>().swap(table);
    if (rc != GSI_OK) { fclose(out); return fail(ctx, "gsi_set_weights_host"); }
    EigenSink sink{&csr, out, n_threads, 0, binary};
    rc = gsi_precompute_stream(ctx, (int64_t)csr.users.size(), csr.offsets.data(), csr.items.data(), eigen_sink, &sink);
    fclose(out);
    if (rc != GSI_OK) return fail(ctx, "gsi_precompute_stream");
    gsi_destroy(ctx);
    return 0;
}

}  // namespace gsihost
