// local_calc [--pct P] [--verbosity V] -- drop-in for local_calc.cpp:563-673 (the per-MOVIE variant, no out_eigen_).
// Reads ./out_fin_* (graph_loader :102-117) and ./out_test_rat_* (graph_test_loader :119-141); writes out_res_* lines
// "movie user' mse kk" (graph_writer :549-561).  Both engines (neigh_program :166-200, vertex_program :246-526) are one
// gsi_local_calc_host() call: local graph and eigensolve per movie, exact cutoff and least squares per (movie, user).
//
// --pct samples movie vertices with rand() seeded by time (:266, :567); here the seed is GSI_SEED when set, time
// otherwise.  Movies whose local graph has fewer than 3 nodes produce no lines (:271-272).
#include <math.h>
#include <time.h>

#include <random>

#include "host_io.hpp"
using namespace gsihost;

int main(int argc, char** argv) {
    unsigned comp_pct = 100;
    int verbosity = 0;
    int positional = 0;
    for (int i = 1; i < argc; ++i) {                      // clopts: --pct, --verbosity, both positional :570-581
        std::string a = argv[i];
        auto value = [&](const std::string& name, std::string& out) -> bool {
            if (a == "--" + name && i + 1 < argc) { out = argv[++i]; return true; }
            if (starts_with(a, "--" + name + "=")) { out = a.substr(name.size() + 3); return true; }
            return false;
        };
        std::string v;
        if (value("pct", v)) comp_pct = (unsigned)atoi(v.c_str());
        else if (value("verbosity", v)) verbosity = atoi(v.c_str());
        else if (!starts_with(a, "--")) { if (positional++ == 0) comp_pct = (unsigned)atoi(a.c_str()); else verbosity = atoi(a.c_str()); }
        else { printf("Error in parsing command line arguments.\n"); return EXIT_FAILURE; }
    }
    printf("Loading graph.\n");
    std::vector<double> table;
    int wrows = 1;
    load_weights_table(table, wrows);
    std::map<unsigned, std::map<unsigned, double>> test_rat;      // movie -> user' -> rating
    load_movie_ratings("out_test_rat_", test_rat, /*as_float=*/true);
    // user CSR of the test ratings (ascending user', ascending movie)
    std::map<unsigned, std::vector<std::pair<unsigned, double>>> by_user;
    for (auto& mv : test_rat)
        for (auto& ur : mv.second) by_user[ur.first].emplace_back(mv.first, ur.second);
    std::vector<unsigned> users;
    std::vector<int64_t> offsets(1, 0);
    std::vector<int32_t> items;
    std::vector<double> ratings;
    for (auto& u : by_user) {
        users.push_back(u.first);
        for (auto& mr : u.second) { items.push_back((int32_t)mr.first); ratings.push_back(mr.second); }
        offsets.push_back((int64_t)items.size());
    }
    const int64_t nu = (int64_t)users.size(), nnz = offsets[nu];
    // --pct samples movie vertices (:266)
    const char* seed_env = getenv("GSI_SEED");
    std::mt19937 rng(seed_env ? (unsigned)atoll(seed_env) : (unsigned)time(NULL));
    std::map<unsigned, bool> chosen;
    for (auto& mv : test_rat) chosen[mv.first] = (rng() % 100) < comp_pct;
    std::vector<uint8_t> mask(nnz, 0);
    for (int64_t t = 0; t < nnz; ++t) mask[t] = chosen[(unsigned)items[t]] ? 1 : 0;
    // ---- GPU ----
    gsi_ctx* ctx = nullptr;
    const char* dev = getenv("GSI_DEVICE");
    if (gsi_create(&ctx, dev ? atoi(dev) : 0, nullptr) != GSI_OK) return fail(nullptr, "gsi_create");
    if (gsi_set_weights_host(ctx, table.data(), wrows) != GSI_OK) return fail(ctx, "gsi_set_weights_host");
    std::vector<double>().swap(table);
    std::vector<float> err(nnz);
    std::vector<int32_t> kk(nnz), status(nnz), cols(nnz);
    std::vector<double> pred(nnz), w_lim(nnz);
    printf("Running ...\n");
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (gsi_local_calc_host(ctx, nu, offsets.data(), items.data(), ratings.data(), mask.data(), err.data(), kk.data(), pred.data(),
                            status.data(), cols.data(), w_lim.data()) != GSI_OK)
        return fail(ctx, "gsi_local_calc_host");
    clock_gettime(CLOCK_MONOTONIC, &t1);
    gsi_destroy(ctx);
    // ---- out_res: "movie user' mse kk\n", ascending movie then user' ----
    struct Row { unsigned movie, user; float mse; int kk, status; double w_lim; int cols; };
    std::vector<Row> rowsv;
    for (int64_t u = 0; u < nu; ++u)
        for (int64_t j = offsets[u]; j < offsets[u + 1]; ++j)
            if (status[j] != GSI_PRED_SKIPPED) rowsv.push_back({(unsigned)items[j], users[u], err[j], kk[j], status[j], w_lim[j], cols[j]});
    std::sort(rowsv.begin(), rowsv.end(), [](const Row& a, const Row& b) { return a.movie != b.movie ? a.movie < b.movie : a.user < b.user; });
    FILE* f = fopen("out_res_1_of_1", "w");
    if (!f) { perror("out_res_1_of_1"); return EXIT_FAILURE; }
    std::string res_text;
    double se = 0, se_ok = 0;
    size_t cnt = 0, ok = 0, illposed = 0, empty = 0;
    for (const Row& r : rowsv) {
        append_int(res_text, r.movie); append_int(res_text, r.user); append_g(res_text, (double)r.mse); append_int(res_text, r.kk);
        res_text.back() = '\n';                          // "movie user' mse kk\n"  (:552-556)
        if (res_text.size() > (1u << 22)) { fwrite(res_text.data(), 1, res_text.size(), f); res_text.clear(); }
        if (verbosity == 1)
            printf("==== Showing movieID: %u userID: %u ==== w_lim: %g lim: %d known: %d mse: %g\n", r.movie, r.user, r.w_lim, r.cols, r.kk, (double)r.mse);
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
