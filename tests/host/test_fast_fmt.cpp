// CPU check of collaborative_filtering_b200/csrc/host/fast_fmt.hpp against libc: format_g6 == printf("%g"),
// parse_double == strtod (bitwise), on the value classes the tools print / read.  Exit code 0 = identical.
#include <time.h>
#include <random>
#include <string>
#include <vector>
#include "../../collaborative_filtering_b200/csrc/host/fast_fmt.hpp"

using namespace gsihost;

static long n_fmt = 0, n_parse = 0, n_bad = 0;

static void check(double v) {
    char a[64], b[64];
    const int la = format_g6(a, v);
    a[la] = 0;
    snprintf(b, sizeof b, "%g", v);
    ++n_fmt;
    if (strcmp(a, b) != 0) { if (n_bad++ < 20) fprintf(stderr, "format mismatch: %.17g -> fast '%s' libc '%s'\n", v, a, b); }
    // parse what was printed, and a 17-digit rendering (slow path), and compare with strtod
    for (int rep = 0; rep < 2; ++rep) {
        char t[64];
        if (rep == 0) strcpy(t, b); else snprintf(t, sizeof t, "%.17g", v);
        double x; const char* e;
        const bool ok = parse_double(t, t + strlen(t), x, e);
        char* q; const double y = strtod(t, &q);
        ++n_parse;
        if (!ok || e != q || memcmp(&x, &y, 8) != 0) { if (n_bad++ < 20) fprintf(stderr, "parse mismatch: '%s' fast %.17g libc %.17g\n", t, x, y); }
    }
}

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 300000;
    std::mt19937_64 rng(31413);
    std::uniform_real_distribution<double> u01(0.0, 1.0);
    std::normal_distribution<double> nrm(0.0, 1.0);
    const double specials[] = {0.0, -0.0, 1.0, -1.0, 10.0, 100000.0, 999999.5, 999999.4999999, 1e6, 1e-4, 1e-5, 0.1, 0.09999999999999999,
                               0.5, 1.5, 2.5e-5, 123456.5, 1234565.0, 1.0000005, 1.00000049999, 9.9999995, 99999.95, 1e22, 1e23, 1e-22,
                               5e-324, 1.7976931348623157e308, 2.2250738585072014e-308, 1.01, 0.707107, -1.11022e-16, 4.9406564584124654e-324,
                               1e290, 1e-290, 9.999995e289, 0.000099999949999, 0.00009999995, INFINITY, -INFINITY, NAN};
    for (double v : specials) check(v);
    for (int i = 0; i < iters; ++i) {
        check(u01(rng));                                    // eigenvalues / weights in (0, 1)
        check(1.0 + 0.5 * u01(rng));                        // sig_min
        check(nrm(rng) * pow(10.0, -3.0 * u01(rng)));       // eigenvector entries
        check(nrm(rng) * 1e-16);                            // round-off sized entries
        const double m = 100000.0 + (double)(rng() % 900000) + 0.5;          // exact ties at 6 digits
        check(m * pow(10.0, (double)((int)(rng() % 21) - 10)));
        check(ldexp(u01(rng) + 0.5, (int)(rng() % 1800) - 900));             // whole exponent range
        uint64_t bits = rng();
        double r; memcpy(&r, &bits, 8);
        check(r);                                           // arbitrary bit patterns (incl. NaN payloads, denormals)
    }
    printf("format checks %ld, parse checks %ld, mismatches %ld\n", n_fmt, n_parse, n_bad);
    // speed of the record writer's inner loop: eigenvector-like values
    std::vector<double> vals(2000000);
    for (double& x : vals) x = nrm(rng) * pow(10.0, -3.0 * u01(rng));
    char buf[64];
    long sink = 0;
    auto now = []() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; };
    double t0 = now();
    for (double x : vals) sink += snprintf(buf, sizeof buf, "%g ", x);
    double t1 = now();
    for (double x : vals) sink += format_g6(buf, x);
    double t2 = now();
    std::string text;
    for (double x : vals) { int l = format_g6(buf, x); text.append(buf, l); text.push_back(' '); }
    double t3 = now();
    double acc = 0;
    for (const char* q = text.data(), *e = q + text.size(); q < e;) { char* r; acc += strtod(q, &r); q = r + 1; }
    double t4 = now();
    for (const char* q = text.data(), *e = q + text.size(); q < e;) { double x; const char* r; parse_double(q, e, x, r); acc -= x; q = r + 1; }
    double t5 = now();
    printf("ns per value: snprintf %.1f, format_g6 %.1f, strtod %.1f, parse_double %.1f (%ld %g)\n", 1e9 * (t1 - t0) / vals.size(),
           1e9 * (t2 - t1) / vals.size(), 1e9 * (t4 - t3) / vals.size(), 1e9 * (t5 - t4) / vals.size(), sink, acc);
    return n_bad ? 1 : 0;
}
