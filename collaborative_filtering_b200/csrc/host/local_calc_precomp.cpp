// local_calc_precomp [--pct P] [--verbosity V] -- drop-in for local_calc_precomp.cpp:484-583.
// Reads out_eigen_ (load_precomputed_data :406-482), ./out_fin_* (graph_loader :122-136) and
// ./out_test_rat_* (graph_test_loader :138-160); writes out_res_* lines "movie user' mse kk"
// (graph_writer :393-404).  apply() over all (movie, test user) pairs (:217-380) is one
// gsi_predict_host() call.
//
// Reference quirks kept or made explicit:
//   B1  the parser's sigs_min vector is never cleared, so the cutoff of pair (user, row) is read
//       from the concatenation of all records read so far (:414,437,440,271).  Kept by default for
//       drop-in behaviour; --fix-b1 (or GSI_FIX_B1=1) uses the record's own sig_min.
//   B10 --pct samples movie vertices with rand() seeded by time (:489,221); here the seed is
//       GSI_SEED when set, time otherwise.
#include <math.h>
#include <time.h>

#include <cstring>
#include <random>

#include "host_io.hpp"
using namespace gsihost;

int main(int argc, char** argv) {
    unsigned comp_pct = 100;
    int verbosity = 0;
    bool fix_b1 = getenv("GSI_FIX_B1") && atoi(getenv("GSI_FIX_B1")) != 0;
    int positional = 0;
    for (int i = 1; i < argc; ++i) {                      // clopts: --pct, --verbosity, both positional :493-504
        std::string a = argv[i];
        auto value = [&](const std::string& name, std::string& out) -> bool {
            if (a == "--" + name && i + 1 < argc) { out = argv[++i]; return true; }
            if (starts_with(a, "--" + name + "=")) { out = a.substr(name.size() + 3); return true; }
            return false;
        };
        std::string v;
        if (value("pct", v)) comp_pct = (unsigned)atoi(v.c_str());
        else if (value("verbosity", v)) verbosity = atoi(v.c_str());
        else if (a == "--fix-b1") fix_b1 = true;
        else if (!starts_with(a, "--")) { if (positional++ == 0) comp_pct = (unsigned)atoi(a.c_str()); else verbosity = atoi(a.c_str()); }
        else { printf("Error in parsing command line arguments.\n"); return EXIT_FAILURE; }
    }
    // ---- out_eigen_ ----
    printf("Loading precomputed data.\nReading file: out_eigen_\n");
    const char* bin_env = getenv("GSI_EIGEN_BINARY");     // binary side format written by precompute_local under the same switch
    const bool binary = bin_env && atoi(bin_env) != 0;
    std::string text;
    if (!read_file(binary ? "out_eigen_.bin" : "out_eigen_", text)) { fprintf(stderr, "cannot read %s\n", binary ? "out_eigen_.bin" : "out_eigen_"); return EXIT_FAILURE; }
    std::vector<unsigned> users;                          // user' per record, arrival order
    std::vector<int64_t> offsets(1, 0), lam_off, vec_off;
    std::vector<int32_t> items, kvec;
    std::vector<double> sig_own, sig_run, lam, vec, w_lim;
    std::unordered_map<unsigned, int64_t> rec_of;         // last record of a user wins (:470)
    int state = 0, n = 0, k = 0;
    bool bad = false;
    if (binary) {                                         // precompute_common.hpp: "GSIEIG01" then records
        const char* p = text.data();
        const char* end = p + text.size();
        bad = text.size() < 8 || memcmp(p, "GSIEIG01", 8) != 0;
        p += 8;
        while (!bad && p < end) {
            int32_t head[4];
            if (end - p < 16) { bad = true; break; }
            memcpy(head, p, 16); p += 16;
            n = head[1]; k = head[2];
            const size_t ib = ((size_t)n + (n & 1)) * 4, need = ib + ((size_t)n + k + (size_t)n * k) * 8;
            if (n < 0 || k < 0 || (size_t)(end - p) < need) { bad = true; break; }
            users.push_back((unsigned)head[0]);
            const size_t i0 = items.size(), s0 = sig_own.size();
            items.resize(i0 + n); memcpy(items.data() + i0, p, (size_t)n * 4); p += ib;
            sig_own.resize(s0 + n); memcpy(sig_own.data() + s0, p, (size_t)n * 8); p += (size_t)n * 8;
            sig_run.insert(sig_run.end(), sig_own.begin() + s0, sig_own.end());             // never cleared in the reference (B1)
            for (int i = 0; i < n; ++i) w_lim.push_back(fix_b1 ? sig_own[s0 + i] : sig_run[i]);
            offsets.push_back((int64_t)items.size());
            kvec.push_back(k);
            lam_off.push_back((int64_t)lam.size());
            lam.resize(lam.size() + k); memcpy(lam.data() + lam.size() - k, p, (size_t)k * 8); p += (size_t)k * 8;
            vec_off.push_back((int64_t)vec.size());
            vec.resize(vec.size() + (size_t)n * k); memcpy(vec.data() + vec.size() - (size_t)n * k, p, (size_t)n * k * 8); p += (size_t)n * k * 8;
            rec_of[users.back()] = (int64_t)users.size() - 1;
        }
    } else
    for_each_line(text, [&](const char* b, const char* e) {
        if (bad) return;
        LineTok t(b, e);
        double v;
        if (state == 0) {
            unsigned long long u, nn, kk;
            if (!(t.next_u64(u) && t.next_u64(nn) && t.next_u64(kk))) { bad = true; return; }
            n = (int)nn; k = (int)kk;
            users.push_back((unsigned)u);
            for (int i = 0; i < n; ++i) {
                unsigned long long m;
                if (!(t.next_u64(m) && t.next_double(v))) { bad = true; return; }     // assert :434
                items.push_back((int32_t)m);
                sig_own.push_back(v);
                sig_run.push_back(v);                     // never cleared in the reference (B1)
            }
            // cutoff of pair (record, row i): sigs_min[i] of the running vector (:271)
            for (int i = 0; i < n; ++i) w_lim.push_back(fix_b1 ? sig_own[sig_own.size() - n + i] : sig_run[i]);
            offsets.push_back((int64_t)items.size());
            kvec.push_back(k);
            state = 1;
        } else if (state == 1) {
            lam_off.push_back((int64_t)lam.size());
            for (int i = 0; i < k; ++i) { if (!t.next_double(v)) { bad = true; return; } lam.push_back(v); }
            state = 2;
        } else {
            vec_off.push_back((int64_t)vec.size());
            for (int64_t i = 0; i < (int64_t)n * k; ++i) { if (!t.next_double(v)) { bad = true; return; } vec.push_back(v); }
            rec_of[users.back()] = (int64_t)users.size() - 1;
            state = 0;
        }
    });
    if (bad || state != 0) { fprintf(stderr, "out_eigen_: malformed record\n"); return EXIT_FAILURE; }
    printf("Loaded %zu test users\n", rec_of.size());
    // ---- graph + test ratings ----
    printf("Loading graph.\n");
    std::vector<double> table;
    int wrows = 1;
    load_weights_table(table, wrows);
    std::map<unsigned, std::map<unsigned, double>> test_rat;
    load_movie_ratings("out_test_rat_", test_rat, /*as_float=*/true);
    const int64_t nu = (int64_t)users.size();
    const int64_t nnz = offsets[nu];
    // the user's own ratings aligned with the record rows; a missing entry reads as 0 like the
    // reference's map operator[] (:257)
    std::vector<double> ratings(nnz, 0.0);
    std::vector<uint8_t> mask(nnz, 0);
    for (int64_t u = 0; u < nu; ++u)
        for (int64_t j = offsets[u]; j < offsets[u + 1]; ++j) {
            auto mv = test_rat.find((unsigned)items[j]);
            if (mv == test_rat.end()) continue;
            auto ur = mv->second.find(users[u]);
            if (ur != mv->second.end()) ratings[j] = ur->second;
        }
    // pairs: every (movie vertex with >= 1 rating, user in its ratings) whose user has a record;
    // --pct samples movie vertices (:221)
    const char* seed_env = getenv("GSI_SEED");
    std::mt19937 rng(seed_env ? (unsigned)atoll(seed_env) : (unsigned)time(NULL));
    size_t skipped_users = 0;
    for (auto& mv : test_rat) {
        if (mv.second.empty() || !((rng() % 100) < comp_pct)) continue;
        for (auto& ur : mv.second) {
            auto r = rec_of.find(ur.first);
            if (r == rec_of.end()) { ++skipped_users; continue; }
            const int64_t u = r->second;
            const int32_t* b = items.data() + offsets[u];
            const int32_t* e = items.data() + offsets[u + 1];
            const int32_t* it = std::find(b, e, (int32_t)mv.first);
            if (it != e) mask[it - items.data()] = 1;
        }
    }
    if (skipped_users) fprintf(stderr, "warning: %zu test ratings of users without an out_eigen_ record skipped\n", skipped_users);
    // ---- GPU ----
    // every visible device (GSI_DEVICE / GSI_DEVICES restrict them): users with all their pairs are dealt over the GPUs
    gsi_group* grp = nullptr;
    std::vector<int> devs;
    if (const char* one = getenv("GSI_DEVICE")) devs.push_back(atoi(one));
    else if (const char* list = getenv("GSI_DEVICES"))
        for (const char* p = list; *p;) { char* q; const long v = strtol(p, &q, 10); if (q == p) break; devs.push_back((int)v); p = (*q == ',') ? q + 1 : q; }
    if (gsi_group_create(&grp, (int)devs.size(), devs.empty() ? nullptr : devs.data()) != GSI_OK) return fail(nullptr, "gsi_group_create");
    auto group_fail = [&](const char* what) { fprintf(stderr, "%s: %s\n", what, gsi_group_last_error(grp)); gsi_group_destroy(grp); return EXIT_FAILURE; };
    if (gsi_group_set_weights_host(grp, table.data(), wrows) != GSI_OK) return group_fail("gsi_group_set_weights_host");
    std::vector<double>().swap(table);
    std::vector<float> err(nnz);
    std::vector<int32_t> kk(nnz), status(nnz), cols(nnz);
    std::vector<double> pred(nnz);
    printf("Running on %d device(s) ...\n", gsi_group_size(grp));
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (gsi_group_predict_host(grp, nu, offsets.data(), items.data(), w_lim.data(), ratings.data(), kvec.data(), lam_off.data(),
                               vec_off.data(), lam.data(), (int64_t)lam.size(), vec.data(), (int64_t)vec.size(), mask.data(), err.data(),
                               kk.data(), pred.data(), status.data(), cols.data()) != GSI_OK)
        return group_fail("gsi_group_predict_host");
    clock_gettime(CLOCK_MONOTONIC, &t1);
    gsi_group_destroy(grp);
    // ---- out_res: "movie user' mse kk\n", ascending movie then user' ----
    struct Row { unsigned movie, user; float mse; int kk, status; };
    std::vector<Row> rowsv;
    for (int64_t u = 0; u < nu; ++u)
        for (int64_t j = offsets[u]; j < offsets[u + 1]; ++j)
            if (mask[j]) rowsv.push_back({(unsigned)items[j], users[u], err[j], kk[j], status[j]});
    std::sort(rowsv.begin(), rowsv.end(), [](const Row& a, const Row& b) { return a.movie != b.movie ? a.movie < b.movie : a.user < b.user; });
    FILE* f = fopen("out_res_1_of_1", "w");
    if (!f) { perror("out_res_1_of_1"); return EXIT_FAILURE; }
    std::string res_text;
    double se = 0, se_ok = 0;
    size_t cnt = 0, ok = 0, illposed = 0, empty = 0;
    for (const Row& r : rowsv) {
        append_int(res_text, r.movie); append_int(res_text, r.user); append_g(res_text, (double)r.mse); append_int(res_text, r.kk);
        res_text.back() = '\n';                          // "movie user' mse kk\n"  (:397-399)
        if (res_text.size() > (1u << 22)) { fwrite(res_text.data(), 1, res_text.size(), f); res_text.clear(); }
        if (verbosity == 1 && r.mse != r.mse)
            printf("==== NaN: movieID: %u userID: %u connected: %d ====\n", r.movie, r.user, r.kk);
        if (r.status == GSI_PRED_EMPTY) { ++empty; continue; }
        se += r.mse; ++cnt;
        if (r.status == GSI_PRED_OK) { se_ok += r.mse; ++ok; } else ++illposed;
    }
    fwrite(res_text.data(), 1, res_text.size(), f);
    fclose(f);
    const double secs = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    printf("----------------------------------------------------------\n");
    printf("Final Runtime (seconds):   %g\nUpdates executed: %zu\nUpdate Rate (updates/second): %g\n", secs, rowsv.size(), rowsv.size() / secs);
    printf("RMSE (all non-empty pairs): %g over %zu; RMSE (well-posed pairs): %g over %zu; ill-posed: %zu; empty: %zu\n",
           cnt ? sqrt(se / cnt) : 0.0, cnt, ok ? sqrt(se_ok / ok) : 0.0, ok, illposed, empty);
    return EXIT_SUCCESS;
}
