// CPU restatement of the reference's per-user task -- TEST INFRASTRUCTURE ONLY (timed CPU baseline
// and second checker).  Nothing under collaborative_filtering_b200/ links or loads this file.
//
// PARITY UNPINNED: the reference cannot be built here (Eigen 3.x, Boost incl. the unofficial
// boost/threadpool.hpp, GraphLab v2.x are neither vendored nor installed), and it ships no golden
// vectors for this path.  This file restates compute_eigens()
// (/root/reference/precompute_local_threads.cpp:100-213) step by step, *including* the work the
// reference wastes -- the dense LU inverse of the diagonal degree matrix (:149) and the two dense
// n^3 products (:155) -- so that the timed baseline is honest about what the reference executes.
// Eigen's SelfAdjointEigenSolver (:164) is restated by the same algorithm class it documents:
// Householder tridiagonalisation with accumulated transformations followed by implicit-shift QL
// iterations on the tridiagonal matrix with the rotations applied to the eigenvector matrix, then
// an ascending sort (EISPACK tred2/tql2 formulation).  It is validated against
// oracle/gsi_oracle.py (numpy.linalg.eigh) in tests/test_oracle.py.
//
// The thread pool restates precompute_local_threads.cpp:300-314 (boost::threadpool, one task per
// user, shared read-only weights) with std::thread workers and an atomic task counter; the text
// record of :196-211 is formatted per user (printf("%g") == default ostream formatting).
//
// local_calc_movie() / cpuref_local_calc() restate the per-MOVIE variant's vertex program
// (/root/reference/local_calc.cpp:262-526) the same way: the same solver pair for both of its eigensolves, the dense
// products and the LU inverse it executes per user.  Validated against oracle/gsi_oracle.py in tests/test_local_calc.py.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

// ---- dense helpers (column-major, leading dimension n) ------------------------------------
inline double& at(std::vector<double>& m, int n, int i, int j) { return m[(size_t)j * n + i]; }

// C = A * B, all n x n column-major.  Straightforward jki loop (contiguous inner loop).
void gemm(int n, const double* a, const double* b, double* c) {
    std::fill(c, c + (size_t)n * n, 0.0);
    for (int j = 0; j < n; ++j) {
        double* cj = c + (size_t)j * n;
        for (int k = 0; k < n; ++k) {
            const double bkj = b[(size_t)j * n + k];
            const double* ak = a + (size_t)k * n;
            for (int i = 0; i < n; ++i) cj[i] += ak[i] * bkj;
        }
    }
}

// inv = A^{-1} through partial-pivoting LU against the identity (what MatrixXd::inverse() does for
// dynamic sizes; no singularity check).  a is destroyed.
void lu_inverse(int n, std::vector<double>& a, std::vector<double>& inv) {
    std::vector<int> piv(n);
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = std::fabs(at(a, n, k, k));
        for (int i = k + 1; i < n; ++i)
            if (std::fabs(at(a, n, i, k)) > best) { best = std::fabs(at(a, n, i, k)); p = i; }
        piv[k] = p;
        if (p != k)
            for (int j = 0; j < n; ++j) std::swap(at(a, n, k, j), at(a, n, p, j));
        const double d = at(a, n, k, k);
        for (int i = k + 1; i < n; ++i) at(a, n, i, k) /= d;
        for (int j = k + 1; j < n; ++j) {
            const double akj = at(a, n, k, j);
            double* cj = &a[(size_t)j * n];
            const double* ck = &a[(size_t)k * n];
            for (int i = k + 1; i < n; ++i) cj[i] -= ck[i] * akj;
        }
    }
    inv.assign((size_t)n * n, 0.0);
    for (int j = 0; j < n; ++j) at(inv, n, j, j) = 1.0;
    for (int k = 0; k < n; ++k)
        if (piv[k] != k)
            for (int j = 0; j < n; ++j) std::swap(at(inv, n, k, j), at(inv, n, piv[k], j));
    for (int j = 0; j < n; ++j) {
        double* x = &inv[(size_t)j * n];
        for (int k = 0; k < n; ++k) {           // L y = P e_j
            const double xk = x[k];
            if (xk != 0.0) {
                const double* ck = &a[(size_t)k * n];
                for (int i = k + 1; i < n; ++i) x[i] -= ck[i] * xk;
            }
        }
        for (int k = n - 1; k >= 0; --k) {      // U x = y
            x[k] /= at(a, n, k, k);
            const double xk = x[k];
            const double* ck = &a[(size_t)k * n];
            for (int i = 0; i < k; ++i) x[i] -= ck[i] * xk;
        }
    }
}

// ---- symmetric eigensolver: Householder tridiagonalisation + implicit QL ------------------
// v: n x n column-major, on entry the symmetric matrix (lower triangle is what matters, the
// caller mirrors it), on exit the eigenvectors (columns).  d: eigenvalues ascending.
void tridiagonalize(int n, std::vector<double>& v, std::vector<double>& d, std::vector<double>& e) {
    auto V = [&](int i, int j) -> double& { return v[(size_t)j * n + i]; };
    for (int j = 0; j < n; ++j) d[j] = V(n - 1, j);
    for (int i = n - 1; i > 0; --i) {
        double scale = 0.0, h = 0.0;
        for (int k = 0; k < i; ++k) scale += std::fabs(d[k]);
        if (scale == 0.0) {
            e[i] = d[i - 1];
            for (int j = 0; j < i; ++j) { d[j] = V(i - 1, j); V(i, j) = 0.0; V(j, i) = 0.0; }
        } else {
            for (int k = 0; k < i; ++k) { d[k] /= scale; h += d[k] * d[k]; }
            double f = d[i - 1];
            double g = std::sqrt(h);
            if (f > 0) g = -g;
            e[i] = scale * g;
            h -= f * g;
            d[i - 1] = f - g;
            for (int j = 0; j < i; ++j) e[j] = 0.0;
            for (int j = 0; j < i; ++j) {
                f = d[j];
                V(j, i) = f;
                g = e[j] + V(j, j) * f;
                double* cj = &v[(size_t)j * n];
                for (int k = j + 1; k <= i - 1; ++k) { g += cj[k] * d[k]; e[k] += cj[k] * f; }
                e[j] = g;
            }
            f = 0.0;
            for (int j = 0; j < i; ++j) { e[j] /= h; f += e[j] * d[j]; }
            const double hh = f / (h + h);
            for (int j = 0; j < i; ++j) e[j] -= hh * d[j];
            for (int j = 0; j < i; ++j) {
                f = d[j];
                g = e[j];
                double* cj = &v[(size_t)j * n];
                for (int k = j; k <= i - 1; ++k) cj[k] -= (f * e[k] + g * d[k]);
                d[j] = V(i - 1, j);
                V(i, j) = 0.0;
            }
        }
        d[i] = h;
    }
    for (int i = 0; i < n - 1; ++i) {           // accumulate transformations
        V(n - 1, i) = V(i, i);
        V(i, i) = 1.0;
        const double h = d[i + 1];
        if (h != 0.0) {
            const double* ci1 = &v[(size_t)(i + 1) * n];
            for (int k = 0; k <= i; ++k) d[k] = ci1[k] / h;
            for (int j = 0; j <= i; ++j) {
                double g = 0.0;
                double* cj = &v[(size_t)j * n];
                for (int k = 0; k <= i; ++k) g += ci1[k] * cj[k];
                for (int k = 0; k <= i; ++k) cj[k] -= g * d[k];
            }
        }
        for (int k = 0; k <= i; ++k) V(k, i + 1) = 0.0;
    }
    for (int j = 0; j < n; ++j) { d[j] = V(n - 1, j); V(n - 1, j) = 0.0; }
    V(n - 1, n - 1) = 1.0;
    e[0] = 0.0;
}

void implicit_ql(int n, std::vector<double>& v, std::vector<double>& d, std::vector<double>& e) {
    for (int i = 1; i < n; ++i) e[i - 1] = e[i];
    e[n - 1] = 0.0;
    double f = 0.0, tst1 = 0.0;
    const double eps = std::ldexp(1.0, -52);
    for (int l = 0; l < n; ++l) {
        tst1 = std::max(tst1, std::fabs(d[l]) + std::fabs(e[l]));
        int m = l;
        while (m < n) { if (std::fabs(e[m]) <= eps * tst1) break; ++m; }
        if (m > l) {
            int iter = 0;
            do {
                ++iter;
                double g = d[l];
                double p = (d[l + 1] - g) / (2.0 * e[l]);
                double r = std::hypot(p, 1.0);
                if (p < 0) r = -r;
                d[l] = e[l] / (p + r);
                d[l + 1] = e[l] * (p + r);
                const double dl1 = d[l + 1];
                double h = g - d[l];
                for (int i = l + 2; i < n; ++i) d[i] -= h;
                f += h;
                p = d[m];
                double c = 1.0, c2 = c, c3 = c;
                const double el1 = e[l + 1];
                double s = 0.0, s2 = 0.0;
                for (int i = m - 1; i >= l; --i) {
                    c3 = c2; c2 = c; s2 = s;
                    g = c * e[i];
                    h = c * p;
                    r = std::hypot(p, e[i]);
                    e[i + 1] = s * r;
                    s = e[i] / r;
                    c = p / r;
                    p = c * d[i] - s * g;
                    d[i + 1] = h + s * (c * g + s * d[i]);
                    double* ci = &v[(size_t)i * n];
                    double* ci1 = &v[(size_t)(i + 1) * n];
                    for (int k = 0; k < n; ++k) {
                        const double hk = ci1[k];
                        ci1[k] = s * ci[k] + c * hk;
                        ci[k] = c * ci[k] - s * hk;
                    }
                }
                p = -s * s2 * c3 * el1 * e[l] / dl1;
                e[l] = s * p;
                d[l] = c * p;
            } while (std::fabs(e[l]) > eps * tst1 && iter < 60);
        }
        d[l] = d[l] + f;
        e[l] = 0.0;
    }
    for (int i = 0; i < n - 1; ++i) {           // selection sort ascending, swapping columns
        int k = i;
        double p = d[i];
        for (int j = i + 1; j < n; ++j) if (d[j] < p) { k = j; p = d[j]; }
        if (k != i) {
            d[k] = d[i];
            d[i] = p;
            std::swap_ranges(&v[(size_t)i * n], &v[(size_t)i * n] + n, &v[(size_t)k * n]);
        }
    }
}

struct Job {
    const double* W; int ldw;
    const int64_t* offsets; const int32_t* items; const int32_t* user_ids; int n_users;
    double* sig_min; int32_t* k_out; double* lam_out; const int64_t* lam_off; double* vec_out; const int64_t* vec_off;
    int honest; int format_text;
    std::atomic<int> next{0};
    std::atomic<long long> text_bytes{0};
    std::mutex out_monitor;
};

// compute_eigens(user_id, ratings), precompute_local_threads.cpp:100-213
void compute_eigens(Job& job, int u, std::string& text) {
    const int32_t* movie_list = job.items + job.offsets[u];
    const int n = (int)(job.offsets[u + 1] - job.offsets[u]);
    const int rows = job.ldw;
    std::vector<double> ww((size_t)n * n), dd((size_t)n * n, 0.0), ll((size_t)n * n), ll2((size_t)n * n);
    for (int i = 0; i < n; ++i)                       // :118-125 gather W
        for (int j = 0; j < n; ++j)
            at(ww, n, i, j) = (std::max(movie_list[i], movie_list[j]) >= rows)
                                  ? 0.0 : job.W[(size_t)movie_list[i] * rows + movie_list[j]];
    for (int i = 0; i < n; ++i) {                     // :129-141 degree
        double count = 0;
        for (int j = 0; j < n; ++j) count += at(ww, n, i, j);
        at(dd, n, i, i) = (count == 0) ? 1.0 : count;
    }
    for (size_t t = 0; t < (size_t)n * n; ++t) ll[t] = dd[t] - ww[t];   // :144
    if (job.honest) {                                 // :149-155 dense inverse, sqrt, two products
        std::vector<double> tmp = dd, dd2, t1((size_t)n * n);
        lu_inverse(n, tmp, dd2);
        for (auto& x : dd2) x = std::sqrt(x);
        gemm(n, dd2.data(), ll.data(), t1.data());
        gemm(n, t1.data(), dd2.data(), ll2.data());
    } else {                                          // same values, O(n^2)
        std::vector<double> s(n);
        for (int i = 0; i < n; ++i) s[i] = std::sqrt(1.0 / at(dd, n, i, i));
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) at(ll2, n, i, j) = (s[i] * at(ll, n, i, j)) * s[j];
    }
    // :164-166 eigensolve (lower triangle only)
    std::vector<double> v((size_t)n * n), d(n), e(n);
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) at(v, n, i, j) = (i >= j) ? at(ll2, n, i, j) : at(ll2, n, j, i);
    if (n > 0) { tridiagonalize(n, v, d, e); implicit_ql(n, v, d, e); }
    // :169-182 sig_min with float accumulator
    double* sigs = job.sig_min + job.offsets[u];
    float sig_min_max = 0;
    for (int i = 0; i < n; ++i) {
        float sig = 0;
        for (int j = 0; j < n; ++j) sig += std::pow(at(ll2, n, i, j), 2);
        sig = std::sqrt(sig);
        sigs[i] = sig + 0.01;
        if (sig_min_max < sig) sig_min_max = sig;
    }
    sig_min_max += 0.01;
    // :185-194 cutoff
    int lim;
    for (lim = 0; lim < n; ++lim) if (d[lim] > sig_min_max) break;
    if (lim < 2) lim = 2;
    job.k_out[u] = lim;
    double* lam = job.lam_out + job.lam_off[u];       // max(n,2) slots per user
    double* vec = job.vec_out + job.vec_off[u];
    // canonical sign (H1 convention of the oracle): largest |component| positive
    for (int j = 0; j < std::min(lim, n); ++j) {
        int arg = 0; double best = -1;
        for (int i = 0; i < n; ++i) if (std::fabs(at(v, n, i, j)) > best) { best = std::fabs(at(v, n, i, j)); arg = i; }
        if (at(v, n, arg, j) < 0) for (int i = 0; i < n; ++i) at(v, n, i, j) = -at(v, n, i, j);
    }
    for (int j = 0; j < lim; ++j) lam[j] = (j < n) ? d[j] : 0.0;          // B5: zeros when n < 2
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < lim; ++j) vec[(size_t)i * lim + j] = (j < n) ? at(v, n, i, j) : 0.0;
    if (job.format_text) {                            // :196-211 text record
        char buf[64];
        text.clear();
        text += std::to_string(job.user_ids ? job.user_ids[u] : u) + " " + std::to_string(n) + " " + std::to_string(lim) + " ";
        for (int i = 0; i < n; ++i) { snprintf(buf, sizeof buf, "%d %g ", movie_list[i], sigs[i]); text += buf; }
        text += "\n";
        for (int j = 0; j < lim; ++j) { snprintf(buf, sizeof buf, "%g ", lam[j]); text += buf; }
        text += "\n";
        for (size_t t = 0; t < (size_t)n * lim; ++t) { snprintf(buf, sizeof buf, "%g ", vec[t]); text += buf; }
        text += "\n";
        std::lock_guard<std::mutex> lock(job.out_monitor);   // save_output :89-98
        job.text_bytes += (long long)text.size();
    }
}

// ---- local_calc.cpp: vertex_program::apply for one movie vertex (:262-526), restated with the same solvers --------
// Inputs are the arrays gsi_local_calc_host takes (dense directed weight table, test ratings as a CSR by user); the
// by-movie view of the ratings is built by the caller below.  `honest` executes the reference's dense degree-matrix
// inverse and the two n^3 products (:368-374) like compute_eigens above; the per-user work is always the reference's:
// the GEMM ll2_h * ll2_h^T, a full eigensolve of it (:435), the normal equations through a dense LU inverse (:485-490).
struct LcJob {
    const double* W; int rows;
    const int64_t* offsets; const int32_t* items; const double* ratings; int n_users;
    const uint8_t* pair_mask;
    const std::vector<std::vector<std::pair<int, int64_t>>>* by_movie;     // movie -> (user, position)
    float* err; int32_t* kk; double* pred; int32_t* status; int32_t* lim; double* w_lim;
    int honest;
    std::atomic<int> next{0};
};

inline double lc_edge(const LcJob& J, int a, int b) {       // graph_loader :102-117
    if (a == b || a >= J.rows || b >= J.rows) return 0.0;
    const float f = (float)J.W[(size_t)a * J.rows + b];
    return ((double)f > 0.1) ? (double)f : 0.0;
}

void local_calc_movie(LcJob& J, int m) {
    const auto& pairs = (*J.by_movie)[m];
    if (pairs.empty()) return;
    std::vector<int> nodes(1, m);                            // :283-290 (ascending neighbour order, B6)
    for (int j = 0; j < J.rows; ++j) if (lc_edge(J, m, j) != 0.0) nodes.push_back(j);
    const int n = (int)nodes.size();
    if (n < 3) return;                                       // :271-272
    std::vector<int> pos(J.rows, -1);
    for (int i = 0; i < n; ++i) pos[nodes[i]] = i;
    std::vector<double> ww((size_t)n * n, 0.0), dd((size_t)n * n, 0.0), ll((size_t)n * n), ll2((size_t)n * n);
    for (int i = 1; i < n; ++i) {                            // :324-335
        for (int j = 1; j < n; ++j) if (i != j) at(ww, n, i, j) = lc_edge(J, nodes[i], nodes[j]);
        at(ww, n, 0, i) = at(ww, n, i, 0) = lc_edge(J, m, nodes[i]);
    }
    for (int i = 0; i < n; ++i) {                            // :353-361
        double count = 0;
        for (int j = 0; j < n; ++j) count += at(ww, n, i, j);
        at(dd, n, i, i) = count;
    }
    for (size_t t = 0; t < (size_t)n * n; ++t) ll[t] = dd[t] - ww[t];   // :364
    if (J.honest) {                                          // :368-374
        std::vector<double> tmp = dd, dd2, t1((size_t)n * n);
        lu_inverse(n, tmp, dd2);
        for (auto& x : dd2) x = std::sqrt(x);
        gemm(n, dd2.data(), ll.data(), t1.data());
        gemm(n, t1.data(), dd2.data(), ll2.data());
    } else {
        std::vector<double> s(n);
        for (int i = 0; i < n; ++i) s[i] = std::sqrt(1.0 / at(dd, n, i, i));
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) at(ll2, n, i, j) = (s[i] * at(ll, n, i, j)) * s[j];
    }
    std::vector<double> v((size_t)n * n), d(n), e(n);        // :378
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) at(v, n, i, j) = (i >= j) ? at(ll2, n, i, j) : at(ll2, n, j, i);
    tridiagonalize(n, v, d, e); implicit_ql(n, v, d, e);
    std::vector<double> rat(n);
    std::vector<int> un, rt;
    for (const auto& pr : pairs) {                           // :393 for each user
        const int u = pr.first; const int64_t t0 = pr.second;
        std::fill(rat.begin(), rat.end(), 0.0);
        for (int64_t t = J.offsets[u]; t < J.offsets[u + 1]; ++t)
            if (J.items[t] < J.rows && pos[J.items[t]] >= 0) rat[pos[J.items[t]]] = J.ratings[t];
        const double rat_real = rat[0];
        rat[0] = 0.0;                                        // :405
        un.clear(); rt.clear();
        for (int i = 0; i < n; ++i) (rat[i] == 0.0 ? un : rt).push_back(i);     // :406-413
        const int nu_ = (int)un.size(), kk = (int)rt.size();
        J.kk[t0] = kk;
        // :417-436 smallest eigenvalue of ll2_h * ll2_h^T
        std::vector<double> g((size_t)nu_ * nu_), gd(nu_), ge(nu_);
        for (int a = 0; a < nu_; ++a)
            for (int b = 0; b <= a; ++b) {
                double acc = 0.0;
                for (int j = 0; j < n; ++j) acc += at(ll2, n, un[a], j) * at(ll2, n, un[b], j);
                g[(size_t)b * nu_ + a] = g[(size_t)a * nu_ + b] = acc;
            }
        tridiagonalize(nu_, g, gd, ge); implicit_ql(nu_, g, gd, ge);
        const double wl = std::sqrt(gd[0]);
        J.w_lim[t0] = wl;
        int lim;
        for (lim = 0; lim < n; ++lim) if (d[lim] > wl) break;                   // :443-451
        if (lim < 2) lim = 2;
        J.lim[t0] = lim;
        int st = 0;
        double p;
        if (kk == 0) { st = 1; p = std::nan(""); }                             // 0/0 :487
        else {
            std::vector<double> mm((size_t)lim * lim, 0.0), inv, rhs(lim, 0.0), x(lim, 0.0);
            double mean = 0.0;
            for (int t = 0; t < kk; ++t) mean += rat[rt[t]];
            mean /= kk;
            for (int a = 0; a < lim; ++a) {
                for (int b = 0; b < lim; ++b) {
                    double acc = 0.0;
                    for (int t = 0; t < kk; ++t) acc += at(v, n, rt[t], a) * at(v, n, rt[t], b);
                    mm[(size_t)b * lim + a] = acc;                              // :484-485
                }
                for (int t = 0; t < kk; ++t) rhs[a] += at(v, n, rt[t], a) * (rat[rt[t]] - mean);
            }
            if (kk < lim) st = 2;
            lu_inverse(lim, mm, inv);                                          // mm.inverse() :490
            p = 0.0;
            for (int a = 0; a < lim; ++a) {
                double acc = 0.0;
                for (int b = 0; b < lim; ++b) acc += inv[(size_t)b * lim + a] * rhs[b];
                p += at(v, n, 0, a) * acc;
            }
            p += mean;
        }
        J.pred[t0] = p;
        double pc = p;
        if (pc > 5) pc = 5;                                                    // :494-497
        if (pc < 1) pc = 1;
        J.err[t0] = (float)((rat_real - pc) * (rat_real - pc));               // :499
        J.status[t0] = st;
    }
}

}  // namespace

extern "C" {

// local_calc over every movie vertex that has requested pairs; outputs [nnz] aligned with items, status 4 where the
// reference writes no line.  n_threads workers take movie vertices from a shared counter (the GraphLab engine runs
// vertex programs concurrently).  Returns the number of pairs computed.
long long cpuref_local_calc(const double* W, int rows, const int64_t* offsets, const int32_t* items, const double* ratings,
                            int n_users, const uint8_t* pair_mask, int n_threads, int honest, float* err, int32_t* kk,
                            double* pred, int32_t* status, int32_t* lim, double* w_lim) {
    const int64_t nnz = offsets[n_users];
    std::vector<std::vector<std::pair<int, int64_t>>> by_movie(rows);
    for (int64_t t = 0; t < nnz; ++t) { err[t] = 0.f; kk[t] = 0; pred[t] = 0.0; status[t] = 4; lim[t] = 0; w_lim[t] = 0.0; }
    for (int u = 0; u < n_users; ++u)
        for (int64_t t = offsets[u]; t < offsets[u + 1]; ++t)
            if (items[t] >= 0 && items[t] < rows && (!pair_mask || pair_mask[t])) by_movie[items[t]].emplace_back(u, t);
    LcJob J;
    J.W = W; J.rows = rows; J.offsets = offsets; J.items = items; J.ratings = ratings; J.n_users = n_users;
    J.pair_mask = pair_mask; J.by_movie = &by_movie; J.err = err; J.kk = kk; J.pred = pred; J.status = status; J.lim = lim;
    J.w_lim = w_lim; J.honest = honest;
    std::vector<std::thread> pool;
    auto worker = [&]() {
        for (;;) {
            const int m = J.next.fetch_add(1);
            if (m >= rows) break;
            local_calc_movie(J, m);
        }
    };
    if (n_threads < 1) n_threads = 1;
    for (int t = 0; t < n_threads; ++t) pool.emplace_back(worker);
    for (auto& th : pool) th.join();
    long long done = 0;
    for (int64_t t = 0; t < nnz; ++t) done += status[t] != 4;
    return done;
}

// Outputs: sig_min[nnz]; k_out[n_users]; lam_out + lam_off[u]: max(n,2) slots, first k valid;
// vec_out + vec_off[u]: max(n,2)*n slots holding the n x k row-major kept eigenvectors.
// honest=1 executes the reference's dense inverse + 2 GEMMs.  Returns the total bytes of
// formatted text (0 when format_text == 0).
long long cpuref_precompute(const double* W, int ldw, const int64_t* offsets, const int32_t* items,
                            const int32_t* user_ids, int n_users, int n_threads, int honest,
                            int format_text, double* sig_min, int32_t* k_out, double* lam_out,
                            const int64_t* lam_off, double* vec_out, const int64_t* vec_off) {
    Job job;
    job.W = W; job.ldw = ldw; job.offsets = offsets; job.items = items; job.user_ids = user_ids;
    job.n_users = n_users; job.sig_min = sig_min; job.k_out = k_out; job.lam_out = lam_out;
    job.lam_off = lam_off; job.vec_out = vec_out; job.vec_off = vec_off; job.honest = honest;
    job.format_text = format_text;
    std::vector<std::thread> pool;                    // pool tp(n_threads) :300
    auto worker = [&]() {
        std::string text;
        for (;;) {
            const int u = job.next.fetch_add(1);      // tp.schedule(compute_eigens, ...) :311
            if (u >= n_users) break;
            compute_eigens(job, u, text);
        }
    };
    if (n_threads < 1) n_threads = 1;
    for (int t = 0; t < n_threads; ++t) pool.emplace_back(worker);
    for (auto& th : pool) th.join();                  // tp.wait() :314
    return job.text_bytes.load();
}

int cpuref_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
