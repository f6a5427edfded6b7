// knn3 -- drop-in for knn3.cpp:266-331: reads ./out_fin_* (edges kept iff (float)w > 0.1) and
// ./out_test_rat_*, prints "Knn Average MSE: x".
#include "host_io.hpp"
using namespace gsihost;

int main(int, char**) {
    std::vector<double> table;
    int wrows = 1;
    load_weights_table(table, wrows);
    std::map<unsigned, std::map<unsigned, double>> test_rat;
    load_movie_ratings("out_test_rat_", test_rat, /*as_float=*/true);
    std::vector<Triple> rowsv;
    unsigned mx = (unsigned)wrows - 1;
    for (auto& mv : test_rat) {
        mx = std::max(mx, mv.first);
        for (auto& ur : mv.second) rowsv.push_back({ur.first, mv.first, ur.second});
    }
    if ((int)mx + 1 > wrows) {                           // movies that only appear in the test files
        const int nr = (int)mx + 1;
        std::vector<double> t2((size_t)nr * nr, 0.0);
        for (int i = 0; i < wrows; ++i) memcpy(&t2[(size_t)i * nr], &table[(size_t)i * wrows], (size_t)wrows * 8);
        table.swap(t2);
        wrows = nr;
    }
    Csr csr = build_csr(rowsv, false);
    std::vector<float> r32(csr.ratings.begin(), csr.ratings.end());
    gsi_ctx* ctx = nullptr;
    const char* dev = getenv("GSI_DEVICE");
    if (gsi_create(&ctx, dev ? atoi(dev) : 0, nullptr) != GSI_OK) return fail(nullptr, "gsi_create");
    if (gsi_set_weights_host(ctx, table.data(), wrows) != GSI_OK) return fail(ctx, "gsi_set_weights_host");
    std::vector<float> err(wrows);
    std::vector<int32_t> cnt(wrows);
    std::vector<uint8_t> has(wrows);
    if (gsi_knn3_host(ctx, (int64_t)csr.users.size(), csr.offsets.data(), csr.items.data(), r32.data(), err.data(), cnt.data(), has.data()) != GSI_OK)
        return fail(ctx, "gsi_knn3_host");
    gsi_destroy(ctx);
    float total = 0.f;                                   // float sums as in error_vertex_data :234-256
    int nv = 0;
    for (int m = 0; m < wrows; ++m) {
        if (cnt[m] > 0 || has[m]) ++nv;
        if (cnt[m] > 0) { const float e = err[m] / (float)cnt[m]; if (e == e) total += e; }
    }
    printf("Knn Average MSE: %g\n", nv ? total / nv : 0.0);     // print_finalize :261-264
    return 0;
}
