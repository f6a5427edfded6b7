// Host-side text I/O shared by the CLI tools: the reference's on-disk formats
// and its file-discovery rules.  Parse + print only; all arithmetic is behind the C ABI (gsi.h).
#pragma once
#include <dirent.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <algorithm>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../../include/gsi.h"
#include "fast_fmt.hpp"

namespace gsihost {

static const unsigned kUiMax = 2147483647u;   // std::numeric_limits<int>::max(), precompute_local.cpp:19

inline bool ends_with(const std::string& s, const std::string& suf) {
    return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0;
}
inline bool starts_with(const std::string& s, const std::string& pre) { return s.compare(0, pre.size(), pre) == 0; }

// regular files of `dir` whose name passes `pred`, sorted (precompute_local.cpp:35-80)
template <class Pred>
std::vector<std::string> list_files(const std::string& dir, Pred pred) {
    std::vector<std::string> out;
    DIR* d = opendir(dir.c_str());
    if (!d) return out;
    while (dirent* e = readdir(d)) {
        std::string name = e->d_name;
        if (name == "." || name == "..") continue;
        std::string path = dir + name;
        struct stat st;
        if (stat(path.c_str(), &st) != 0 || !S_ISREG(st.st_mode)) continue;
        if (pred(name)) out.push_back(path);
    }
    closedir(d);
    std::sort(out.begin(), out.end());
    return out;
}

inline bool read_file(const std::string& path, std::string& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? n : 0);
    size_t got = n > 0 ? fread(&out[0], 1, n, f) : 0;
    fclose(f);
    out.resize(got);
    return true;
}

// cursor over whitespace separated tokens of one line
struct LineTok {
    const char* p; const char* end;
    LineTok(const char* b, const char* e) : p(b), end(e) {}
    void skip() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) ++p; }
    bool next_u64(unsigned long long& v) {
        skip();
        if (p >= end || *p < '0' || *p > '9') return false;
        char* q; v = strtoull(p, &q, 10); p = q; return true;
    }
    bool next_double(double& v) {
        skip();
        if (p >= end) return false;
        const char* q;
        if (!parse_double(p, end, v, q)) return false;      // == strtod, fast path for short decimal tokens
        p = q; return true;
    }
};

template <class Fn> void for_each_line(const std::string& text, Fn fn) {
    const char* b = text.data();
    const char* e = b + text.size();
    while (b < e) {
        const char* nl = (const char*)memchr(b, '\n', e - b);
        const char* le = nl ? nl : e;
        if (le > b) fn(b, le);          // empty lines skipped (precompute_local.cpp:103-104)
        b = nl ? nl + 1 : e;
    }
}

struct Triple { unsigned user, item; double rating; };

// "user item rating" lines of one ratings file (collaborative_filtering.dox:62-80)
inline void read_triples(const std::string& path, std::vector<Triple>& out) {
    std::string text;
    if (!read_file(path, text)) return;
    for_each_line(text, [&](const char* b, const char* e) {
        LineTok t(b, e);
        unsigned long long u, m; double r;
        if (t.next_u64(u) && t.next_u64(m) && t.next_double(r)) out.push_back({(unsigned)u, (unsigned)m, r});
    });
}

// CSR by user over (user -> {item -> rating}); ids ascending; duplicates: last wins
struct Csr {
    std::vector<unsigned> users;        // key of each row (ascending)
    std::vector<int64_t> offsets;
    std::vector<int32_t> items;
    std::vector<double> ratings;
};
inline Csr build_csr(const std::vector<Triple>& rows, bool map_user /* user' = INT_MAX - user */) {
    std::map<unsigned, std::map<unsigned, double>> by;
    for (const Triple& t : rows) by[map_user ? kUiMax - t.user : t.user][t.item] = t.rating;
    Csr c;
    c.offsets.push_back(0);
    for (auto& u : by) {
        c.users.push_back(u.first);
        for (auto& m : u.second) { c.items.push_back((int32_t)m.first); c.ratings.push_back(m.second); }
        c.offsets.push_back((int64_t)c.items.size());
    }
    return c;
}

// "m1 m2 w" lines of every ./out_fin_* file -> dense directed table, last wins
// (precompute_local.cpp:113-158; the 2000-id cap of :116 is lifted).  rows = max id + 1.
inline bool load_weights_table(std::vector<double>& table, int& rows) {
    std::vector<std::string> files = list_files("./", [](const std::string& n) { return starts_with(n, "out_fin_"); });
    struct E { unsigned a, b; double w; };
    std::vector<E> edges;
    unsigned mx = 0;
    for (const std::string& f : files) {
        printf("Reading file: %s\n", f.c_str());
        std::string text;
        if (!read_file(f, text)) continue;
        for_each_line(text, [&](const char* b, const char* e) {
            LineTok t(b, e);
            unsigned long long a, c; double w;
            if (t.next_u64(a) && t.next_u64(c) && t.next_double(w)) {
                edges.push_back({(unsigned)a, (unsigned)c, w});
                mx = std::max(mx, std::max((unsigned)a, (unsigned)c));
            }
        });
    }
    rows = (int)mx + 1;
    table.assign((size_t)rows * rows, 0.0);
    for (const E& e : edges) table[(size_t)e.a * rows + e.b] = e.w;
    return !files.empty();
}

// "movie user' rating user' rating ..." lines (out_rat_*, out_test_rat_*; knn.cpp:303-332)
inline void load_movie_ratings(const std::string& prefix, std::map<unsigned, std::map<unsigned, double>>& out, bool as_float) {
    std::vector<std::string> files = list_files("./", [&](const std::string& n) { return starts_with(n, prefix); });
    for (const std::string& f : files) {
        std::string text;
        if (!read_file(f, text)) continue;
        for_each_line(text, [&](const char* b, const char* e) {
            LineTok t(b, e);
            unsigned long long m, u; double r;
            if (!t.next_u64(m)) return;
            std::map<unsigned, double>& dst = out[(unsigned)m];
            while (t.next_u64(u) && t.next_double(r)) dst[(unsigned)u] = as_float ? (double)(float)r : r;
        });
    }
}

// "<value> " exactly as the reference's `strm << value << " "` prints it (default ostream format == %g)
inline void append_g(std::string& s, double v) {
    char buf[40];
    int n = format_g6(buf, v);
    buf[n++] = ' ';
    s.append(buf, n);
}
inline void append_int(std::string& s, long long v) {
    char buf[24];
    char* e = buf + sizeof buf;
    char* p = e;
    const bool neg = v < 0;
    unsigned long long u = neg ? 0ULL - (unsigned long long)v : (unsigned long long)v;
    do { *--p = (char)('0' + u % 10); u /= 10; } while (u);
    if (neg) *--p = '-';
    s.append(p, e - p);
    s.push_back(' ');
}

inline int fail(gsi_ctx* ctx, const char* what) {
    fprintf(stderr, "%s: %s\n", what, gsi_last_error(ctx));
    return 1;
}

}  // namespace gsihost
