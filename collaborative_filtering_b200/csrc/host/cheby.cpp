// cheby -- drop-in for cheby.cpp (Chebyshev polynomial graph filter).  No arguments; cwd-relative files
// as in the reference (cheby.cpp:282-288, 377-380):
//   coeff*            filter coefficients, every number of every line (filter_loader :120-135)
//   graph_topology*   lines `a b w`; kept iff w > 0.1, both directions added (graph_loader :86-103)
//   graph_signal*     lines `vertex value` (graph_signal_loader :105-118)
//   -> graph_filtered_signal_1_of_1   lines `vertex value` (graph_signal_writer :140-148), ascending vertex id
// The three GraphLab engines (degree, init values, one synchronous superstep per further coefficient, :312-375) are one
// gsi_cheby_filter_host() call.  Vertices that only appear in the topology read a signal of 0 (uninitialised in the reference).
#include <time.h>

#include <map>

#include "host_io.hpp"

using namespace gsihost;

static double now_s() {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec + 1e-9 * t.tv_nsec;
}

int main(int, char**) {
    const double t_load = now_s();
    std::vector<double> coeff;
    for (const std::string& f : list_files("./", [](const std::string& n) { return n.rfind("coeff", 0) == 0; })) {
        std::string text;
        if (!read_file(f, text)) continue;
        for_each_line(text, [&](const char* b, const char* e) {
            LineTok t(b, e);
            double v;
            while (t.next_double(v)) coeff.push_back(v);
        });
    }
    printf("Loading graph.\nFilter lenght: %zu\n", coeff.size());
    if (coeff.size() < 2) { fprintf(stderr, "cheby: need at least two coefficients in coeff*\n"); return EXIT_FAILURE; }
    struct Edge { unsigned a, b; double w; };
    std::vector<Edge> edges;
    for (const std::string& f : list_files("./", [](const std::string& n) { return n.rfind("graph_topology", 0) == 0; })) {
        std::string text;
        if (!read_file(f, text)) continue;
        for_each_line(text, [&](const char* b, const char* e) {
            LineTok t(b, e);
            unsigned long long va, vb;
            double w;
            if (!(t.next_u64(va) && t.next_u64(vb) && t.next_double(w))) return;
            if (w > 0.1) { edges.push_back({(unsigned)va, (unsigned)vb, w}); edges.push_back({(unsigned)vb, (unsigned)va, w}); }
        });
    }
    std::map<unsigned, double> signal;
    for (const std::string& f : list_files("./", [](const std::string& n) { return n.rfind("graph_signal", 0) == 0; })) {
        std::string text;
        if (!read_file(f, text)) continue;
        for_each_line(text, [&](const char* b, const char* e) {
            LineTok t(b, e);
            unsigned long long v;
            double x;
            if (t.next_u64(v) && t.next_double(x)) signal[(unsigned)v] = x;
        });
    }
    for (const Edge& e : edges) { signal.emplace(e.a, 0.0); signal.emplace(e.b, 0.0); }
    // vertex ids -> 0 .. nv-1 (ascending), CSR over the out-edges in file order
    std::vector<unsigned> ids;
    std::map<unsigned, int> index;
    for (auto& kv : signal) { index[kv.first] = (int)ids.size(); ids.push_back(kv.first); }
    const int64_t nv = (int64_t)ids.size();
    std::vector<int64_t> row_off(nv + 1, 0);
    for (const Edge& e : edges) ++row_off[index[e.a] + 1];
    for (int64_t i = 0; i < nv; ++i) row_off[i + 1] += row_off[i];
    std::vector<int32_t> col(edges.size());
    std::vector<double> w(edges.size());
    std::vector<int64_t> fill(row_off.begin(), row_off.end() - 1);
    for (const Edge& e : edges) { const int64_t p = fill[index[e.a]]++; col[p] = index[e.b]; w[p] = e.w; }
    std::vector<double> x(nv), y(nv);
    for (int64_t i = 0; i < nv; ++i) x[i] = signal[ids[i]];
    // the two timing lines scale2.sh greps from the reference's output (cheby.cpp:287-288, :328-334)
    printf("Loading graph. Finished in %g\n", now_s() - t_load);
    printf("Num vertices: %lld\nNum edges: %zu\n", (long long)nv, edges.size());
    gsi_ctx* ctx = nullptr;
    const char* dev = getenv("GSI_DEVICE");
    if (gsi_create(&ctx, dev ? atoi(dev) : 0, nullptr) != GSI_OK) return fail(nullptr, "gsi_create");
    printf("Running ...\n");
    const double t_run = now_s();
    if (gsi_cheby_filter_host(ctx, nv, row_off.data(), col.data(), w.data(), x.data(), (int)coeff.size(), coeff.data(), y.data()) != GSI_OK)
        return fail(ctx, "gsi_cheby_filter_host");
    const double runtime = now_s() - t_run;
    const double updates = (double)nv * (double)coeff.size();        // one vertex update per vertex and coefficient (:262-264)
    printf("----------------------------------------------------------\nFinal Runtime (seconds):   %g\nUpdates executed: %.0f\n"
           "Update Rate (updates/second): %g\n", runtime, updates, updates / runtime);
    gsi_destroy(ctx);
    FILE* f = fopen("graph_filtered_signal_1_of_1", "w");
    if (!f) { perror("graph_filtered_signal_1_of_1"); return EXIT_FAILURE; }
    std::string text;
    for (int64_t i = 0; i < nv; ++i) { append_int(text, ids[i]); append_g(text, y[i]); text.back() = '\n'; }
    fwrite(text.data(), 1, text.size(), f);
    fclose(f);
    return EXIT_SUCCESS;
}
