// knn -- drop-in for knn.cpp:359-473: reads every file of movielens/ (role by suffix, :88-92),
// writes out_rat_*, out_test_rat_* (per-movie rating lists) and out_edg_* (co-rated movie lists).
// The co-rating relation (three GAS passes in the reference, :160-298) is one GPU kernel.
#include "host_io.hpp"
using namespace gsihost;

int main(int, char**) {
    std::vector<Triple> train, val;
    for (const std::string& f : list_files("movielens/", [](const std::string&) { return true; }))
        read_triples(f, ends_with(f, ".validate") ? val : train);
    // per-movie maps, user' = INT_MAX - user (:103)
    std::map<unsigned, std::map<unsigned, double>> rat, test_rat;
    for (const Triple& t : train) { rat[t.item][kUiMax - t.user] = t.rating; test_rat[t.item]; }
    for (const Triple& t : val) { test_rat[t.item][kUiMax - t.user] = t.rating; rat[t.item]; }
    auto write_rat = [](const char* name, const std::map<unsigned, std::map<unsigned, double>>& m) {
        FILE* f = fopen(name, "w");
        if (!f) { perror(name); exit(1); }
        std::string s;
        for (auto& mv : m) {                             // "movie user' rating ... \n"  :303-332
            s.clear();
            s += std::to_string(mv.first) + " ";
            for (auto& ur : mv.second) { s += std::to_string(ur.first) + " "; append_g(s, ur.second); }
            s += "\n";
            fwrite(s.data(), 1, s.size(), f);
        }
        fclose(f);
    };
    write_rat("out_rat_1_of_1", rat);
    write_rat("out_test_rat_1_of_1", test_rat);
    // co-rated lists over train AND validate edges (:218-227)
    std::vector<Triple> all = train;
    all.insert(all.end(), val.begin(), val.end());
    Csr csr = build_csr(all, true);
    unsigned mx = 0;
    for (const Triple& t : all) mx = std::max(mx, t.item);
    const int rows = (int)mx + 1;
    std::vector<uint8_t> co((size_t)rows * rows, 0);
    gsi_ctx* ctx = nullptr;
    const char* dev = getenv("GSI_DEVICE");
    if (gsi_create(&ctx, dev ? atoi(dev) : 0, nullptr) != GSI_OK) return fail(nullptr, "gsi_create");
    if (gsi_knn_corated_host(ctx, (int64_t)csr.users.size(), csr.offsets.data(), csr.items.data(), rows, co.data()) != GSI_OK)
        return fail(ctx, "gsi_knn_corated_host");
    gsi_destroy(ctx);
    FILE* f = fopen("out_edg_1_of_1", "w");
    if (!f) { perror("out_edg_1_of_1"); return 1; }
    std::string s;
    for (auto& mv : rat) {                               // "movie nbr nbr ... \n", ascending unique :337-357
        s.clear();
        s += std::to_string(mv.first) + " ";
        const uint8_t* row = co.data() + (size_t)mv.first * rows;
        for (int b = 0; b < rows; ++b) if (row[b]) s += std::to_string(b) + " ";
        s += "\n";
        fwrite(s.data(), 1, s.size(), f);
    }
    fclose(f);
    return 0;
}
