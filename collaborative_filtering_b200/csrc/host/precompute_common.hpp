// Shared body of precompute_local / precompute_local_threads: same cwd-relative inputs and the same
// out_eigen_ record layout as the reference (precompute_local.cpp:84-282,
// precompute_local_threads.cpp:215-317); the per-user math is one gsi_precompute_stream() call.
#pragma once
#include <sys/types.h>

#include <atomic>
#include <thread>

#include "host_io.hpp"

namespace gsihost {

struct RecordPos { uint32_t user; int64_t off, len; };      // where a record sits in the file as written

struct EigenSink {
    const Csr* csr; FILE* out; int n_threads; size_t done = 0; bool binary = false;
    int64_t written = 0;                  // bytes behind the header
    std::vector<RecordPos> index;         // one entry per record, arrival order
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
    std::vector<std::vector<int64_t>> ends(nt);          // end of every record inside its part
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t)
        pool.emplace_back([&, t]() {
            const int64_t b = nr * t / nt, e = nr * (t + 1) / nt;
            for (int64_t j = b; j < e; ++j) {
                S->binary ? format_record_bin(*S->csr, ch, j, parts[t]) : format_record(*S->csr, ch, j, parts[t]);
                ends[t].push_back((int64_t)parts[t].size());
            }
        });
    for (auto& th : pool) th.join();
    for (int t = 0; t < nt; ++t) {
        const std::string& p = parts[t];
        if (!p.empty() && fwrite(p.data(), 1, p.size(), S->out) != p.size()) return 1;
        const int64_t b = nr * t / nt;
        int64_t prev = 0;
        for (size_t q = 0; q < ends[t].size(); ++q) {
            S->index.push_back({S->csr->users[ch->user_index[b + (int64_t)q]], S->written + prev, ends[t][q] - prev});
            prev = ends[t][q];
        }
        S->written += (int64_t)p.size();
    }
    S->done += (size_t)nr;
    printf("%g%%\n", 100.0 * (double)S->done / (double)(S->csr->users.size()));
    return 0;
}

// Record order of the file.  The reference writes in boost::unordered_map iteration order (precompute_local.cpp:165) or in
// thread completion order (precompute_local_threads.cpp:89-98) -- neither is reproducible; the DEFINED order here is
// ascending user' (SURVEY.md appendix B6).  Records arrive in processing order (largest users first, and from several
// devices in timing order), so the file is permuted once at the end.  The order matters downstream only through the
// reference's reader bug B1 (local_calc_precomp.cpp:414,437,440: every cutoff is read from the concatenation of all
// records read so far), which makes the default local_calc_precomp output depend on it.
inline bool reorder_records(const char* path, size_t header, std::vector<RecordPos>& index) {
    bool sorted = true;
    for (size_t i = 1; i < index.size() && sorted; ++i) sorted = index[i - 1].user < index[i].user;
    if (sorted) return true;
    std::stable_sort(index.begin(), index.end(), [](const RecordPos& a, const RecordPos& b) { return a.user < b.user; });
    const std::string tmp = std::string(path) + ".tmp";
    FILE* in = fopen(path, "rb");
    FILE* out = fopen(tmp.c_str(), "wb");
    if (!in || !out) { if (in) fclose(in); if (out) fclose(out); return false; }
    std::vector<char> buf;
    bool ok = true;
    if (header) { buf.resize(header); ok = fread(buf.data(), 1, header, in) == header && fwrite(buf.data(), 1, header, out) == header; }
    for (const RecordPos& r : index) {
        if (!ok) break;
        buf.resize((size_t)r.len);
        ok = fseeko(in, (off_t)(header + r.off), SEEK_SET) == 0 && fread(buf.data(), 1, buf.size(), in) == buf.size() &&
             fwrite(buf.data(), 1, buf.size(), out) == buf.size();
    }
    fclose(in);
    ok = (fclose(out) == 0) && ok;
    if (!ok) { remove(tmp.c_str()); return false; }
    return rename(tmp.c_str(), path) == 0;
}

// devices of the run: GSI_DEVICE=k pins one device, GSI_DEVICES=a,b,... lists them, default = every visible device
// (the reference's tool uses every worker of the box, precompute_local_threads.cpp:300-314)
inline std::vector<int> devices_from_env() {
    std::vector<int> devs;
    if (const char* one = getenv("GSI_DEVICE")) { devs.push_back(atoi(one)); return devs; }
    if (const char* list = getenv("GSI_DEVICES")) {
        for (const char* p = list; *p;) {
            char* q;
            const long v = strtol(p, &q, 10);
            if (q == p) break;
            devs.push_back((int)v);
            p = (*q == ',') ? q + 1 : q;
        }
    }
    return devs;                                         // empty = all
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
    gsi_group* grp = nullptr;
    const std::vector<int> devs = devices_from_env();
    if (gsi_group_create(&grp, (int)devs.size(), devs.empty() ? nullptr : devs.data()) != GSI_OK) { fclose(out); return fail(nullptr, "gsi_group_create"); }
    int rc = gsi_group_set_weights_host(grp, table.data(), wrows);
    std::vector<double>().swap(table);
    if (rc != GSI_OK) { fclose(out); fprintf(stderr, "gsi_group_set_weights_host: %s\n", gsi_group_last_error(grp)); gsi_group_destroy(grp); return 1; }
    printf("Devices: %d (weights replicated: %s)\n", gsi_group_size(grp), gsi_group_broadcast_path(grp));
    EigenSink sink;
    sink.csr = &csr; sink.out = out; sink.n_threads = n_threads; sink.binary = binary;
    rc = gsi_group_precompute_stream(grp, (int64_t)csr.users.size(), csr.offsets.data(), csr.items.data(), eigen_sink, &sink);
    const bool closed = fclose(out) == 0;
    if (rc != GSI_OK) { fprintf(stderr, "gsi_group_precompute_stream: %s\n", gsi_group_last_error(grp)); gsi_group_destroy(grp); return 1; }
    gsi_group_destroy(grp);
    if (!closed || !reorder_records(binary ? "out_eigen_.bin" : "out_eigen_", binary ? sizeof kEigenBinMagic : 0, sink.index)) {
        perror("out_eigen_");
        return 1;
    }
    return 0;
}

}  // namespace gsihost
