// knn2 -- drop-in for knn2.cpp:167-219: reads ./out_rat_* (per-movie train ratings) and ./out_edg_*
// (co-rated lists), writes out_fin_* ("src dst w" for w > 0.01).  weights_calc (:127-146) over all
// edges is one GPU pass over the ratings transposed back to CSR-by-user.
#include <set>

#include "host_io.hpp"
using namespace gsihost;

int main(int, char**) {
    std::map<unsigned, std::map<unsigned, double>> rat;
    load_movie_ratings("out_rat_", rat, /*as_float=*/false);
    // edges of the graph: only pairs listed in out_edg_* get a weight (:104-121)
    std::vector<std::string> files = list_files("./", [](const std::string& n) { return starts_with(n, "out_edg_"); });
    std::vector<std::pair<unsigned, unsigned>> edges;
    unsigned mx = 0;
    for (const std::string& f : files) {
        std::string text;
        if (!read_file(f, text)) continue;
        for_each_line(text, [&](const char* b, const char* e) {
            LineTok t(b, e);
            unsigned long long m, j;
            if (!t.next_u64(m)) return;
            mx = std::max(mx, (unsigned)m);
            while (t.next_u64(j)) { edges.push_back({(unsigned)m, (unsigned)j}); mx = std::max(mx, (unsigned)j); }
        });
    }
    for (auto& mv : rat) mx = std::max(mx, mv.first);
    const int rows = (int)mx + 1;
    // transpose: CSR by user' over the train ratings
    std::vector<Triple> rowsv;
    for (auto& mv : rat) for (auto& ur : mv.second) rowsv.push_back({ur.first, mv.first, ur.second});
    Csr csr = build_csr(rowsv, /*map_user=*/false);
    std::vector<float> r32(csr.ratings.begin(), csr.ratings.end());
    gsi_ctx* ctx = nullptr;
    const char* dev = getenv("GSI_DEVICE");
    if (gsi_create(&ctx, dev ? atoi(dev) : 0, nullptr) != GSI_OK) return fail(nullptr, "gsi_create");
    int64_t ne = 0;
    if (gsi_knn_build_host(ctx, (int64_t)csr.users.size(), csr.offsets.data(), csr.items.data(), r32.data(), rows, 0, &ne) != GSI_OK)
        return fail(ctx, "gsi_knn_build_host");
    std::vector<int32_t> a(ne), b(ne);
    std::vector<float> w(ne);
    if (gsi_knn_edges_host(ctx, a.data(), b.data(), w.data(), ne) != GSI_OK) return fail(ctx, "gsi_knn_edges_host");
    gsi_destroy(ctx);
    std::sort(edges.begin(), edges.end());
    FILE* f = fopen("out_fin_1_of_1", "w");
    if (!f) { perror("out_fin_1_of_1"); return 1; }
    std::string text;
    for (int64_t e = 0; e < ne; ++e) {                   // "src dst w\n" iff w > 0.01  :155-163
        if (!std::binary_search(edges.begin(), edges.end(), std::make_pair((unsigned)a[e], (unsigned)b[e]))) continue;
        append_int(text, a[e]); append_int(text, b[e]); append_g(text, (double)w[e]);
        text.back() = '\n';                              // the weight ends the line
        if (text.size() > (1u << 22)) { fwrite(text.data(), 1, text.size(), f); text.clear(); }
    }
    fwrite(text.data(), 1, text.size(), f);
    fclose(f);
    return 0;
}
