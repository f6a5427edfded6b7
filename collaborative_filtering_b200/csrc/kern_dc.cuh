// Large path, stage 2: divide and conquer on the symmetric tridiagonal T (the QR-iteration half of
// Eigen's SelfAdjointEigenSolver, precompute_local.cpp:231, replaced by Cuppen / Gu-Eisenstat D&C so
// that the O(n^3) part is GEMM on the FP64 tensor cores).
//
//   dc_leaf_kernel      T is torn into 2^L leaves of <= 32 rows (cuts on multiples of 8); a warp solves a
//                       leaf with implicit QL, lane = row of the eigenvector block
//   per merge level, batched over every merge node of every user of the chunk:
//   dc_deflate_kernel   z = rows of the two child blocks at the cut, merge-sort of the poles, LAPACK
//                       dlaed2-style deflation (tiny z_i; close poles -> Givens rotation), column types
//   dc_secular_kernel   one thread per root of  1/rho + sum z_i^2/(d_i - lam) = 0, "middle way" rational
//                       interpolation with a bisection safeguard; lam is kept as (origin pole, offset)
//   dc_rank_kernel      output order of roots and deflated values; on a user's last merge also the
//                       reference's cutoff (precompute_local.cpp:252-258) -> only the kept columns follow
//   dc_zhat_kernel      Gu-Eisenstat: z recomputed from the roots (Loewner) so the vectors are orthogonal
//   dc_vectors_kernel   S[g, j] = zhat_g / (d_g - lam_j), normalised
//   dc_gemm_kernel      Qout[:, pos_j] = Qin[:, non-deflated] * S, top and bottom halves separately
//                       (column types keep the K ranges short), DMMA m8n8k4, cp.async 3-stage ring
//   dc_copy_kernel      deflated columns are copied to their sorted position
#pragma once
#include "gsi_internal.cuh"
#include "kern_trd.cuh"
#include "ptx.cuh"
#include "tc_gemm.cuh"

#define DC_LEAF 32
#define DC_EPS 1.1102230246251565e-16     // 2^-53 (LAPACK dlamch('E'))

struct DcLeaf { int job, off, sz; };
struct DcNode { int job, off, n1, n2, final_, pad_; };
struct DcState { int k, k1, k2, k3, kneed, pad_; double rho; };

struct DcParams {
    const HJob* jobs;
    const DcNode* nodes;      // nodes of the current level start at node0
    DcState* state;           // [all nodes]
    int node0, nnodes;
    int in_b;                 // 0: input Q/lam in the "a" buffers, output to "b"; 1: the other way round
    double* Qa; double* Qb; double* S;
    double* lamA; double* lamB;
    const double* e;          // off-diagonal of T
    // per-node scratch, all indexed r_off + off + i
    double* dk; double* zk; double* zhat; double* tau; double* lamk; double* dval;
    int32_t* orig; int32_t* gmap; int32_t* colsrc; int32_t* dsrc; int32_t* pos_nd; int32_t* pos_df;
    double* ds; double* zs; int32_t* src; double* rot;   // deflation scratch (rot: 2 doubles + packed pair per rotation, 4 slots)
    // cutoff
    const unsigned int* sigmax;   // [jobs] float bits of max_i ||row_i||
    int32_t* kuser;               // [jobs] kept eigenpairs (written by the final merge)
    const uint8_t* tc_node;       // optional [all nodes]: 1 = the merge GEMMs of this node run on the tcgen05 engine (tc_gemm.cu)
};

// merge GEMM of a node half on the tensor-core engine: planned on the host (upper bounds), resolved on the device (deflation counts)
struct DcTcTask { int node, bottom; int64_t ap_off, bp_off, e_off; };

// ---------------------------------------------------------------------------------------------
// leaves
// ---------------------------------------------------------------------------------------------
// grid ceil(nleaves/4), block 128: warp per leaf
__global__ void __launch_bounds__(128) dc_leaf_kernel(const HJob* __restrict__ jobs, const DcLeaf* __restrict__ leaves, int nleaves,
                                                      const double* __restrict__ dvec, const double* __restrict__ evec,
                                                      double* __restrict__ lamA, double* __restrict__ Qa) {
    __shared__ double zsm[4][32 * 33];
    __shared__ double dsm[4][32], esm[4][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int li = blockIdx.x * 4 + warp;
    if (li >= nleaves) return;
    const DcLeaf lf = leaves[li];
    const HJob jb = jobs[lf.job];
    const int sz = lf.sz, off = lf.off, n = jb.n;
    const double* d = dvec + jb.r_off;
    const double* e = evec + jb.r_off;
    double* z = zsm[warp];
    double* dd = dsm[warp];
    double* ee = esm[warp];
    if (lane < sz) {
        double v = d[off + lane];
        if (lane == 0 && off > 0) v -= fabs(e[off - 1]);
        if (lane == sz - 1 && off + sz < n) v -= fabs(e[off + sz - 1]);
        dd[lane] = v;
        ee[lane] = (lane < sz - 1) ? e[off + lane] : 0.0;
    }
    for (int c = 0; c < sz; ++c) z[lane * 33 + c] = (lane == c) ? 1.0 : 0.0;
    __syncwarp();
    // implicit QL (tqli); every lane runs the same scalar recurrence on the shared d/e (identical
    // values from every lane; reads and writes of an iteration are separated by __syncwarp), lane = row
    // of Z for the rotations
    for (int l = 0; l < sz; ++l) {
        int iter = 0;
        for (;;) {
            __syncwarp();
            int m = l;
            for (; m < sz - 1; ++m) {
                const double s = fabs(dd[m]) + fabs(dd[m + 1]);
                if (fabs(ee[m]) <= DC_EPS * s) break;
            }
            if (m == l || iter++ >= 60) break;
            const double dl = dd[l], el = ee[l];
            double g = (dd[l + 1] - dl) / (2.0 * el);
            double r = hypot(g, 1.0);
            g = dd[m] - dl + el / (g + copysign(r, g));
            double s = 1.0, c = 1.0, p = 0.0;
            bool under = false;
            for (int i = m - 1; i >= l; --i) {
                const double ei = ee[i], di1 = dd[i + 1], di = dd[i];
                __syncwarp();
                const double f = s * ei, b = c * ei;
                r = hypot(f, g);
                ee[i + 1] = r;
                if (r == 0.0) { dd[i + 1] = di1 - p; ee[m] = 0.0; under = true; break; }
                s = f / r; c = g / r;
                g = di1 - p;
                r = (di - g) * s + 2.0 * c * b;
                p = s * r;
                dd[i + 1] = g + p;
                g = c * r - b;
                const double zf = z[lane * 33 + i + 1], z0 = z[lane * 33 + i];
                z[lane * 33 + i + 1] = s * z0 + c * zf;
                z[lane * 33 + i] = c * z0 - s * zf;
            }
            if (under) continue;
            __syncwarp();
            dd[l] = dl - p; ee[l] = g; ee[m] = 0.0;
        }
    }
    __syncwarp();
    // ascending order by rank
    int rank = 0;
    const double mine = (lane < sz) ? dd[lane] : 0.0;
    for (int c = 0; c < sz; ++c) {
        const double o = dd[c];
        rank += (o < mine || (o == mine && c < lane)) ? 1 : 0;
    }
    double* lam = lamA + jb.r_off;
    double* Q = Qa + jb.m_off;
    const int ld = jb.np;
    if (lane < sz) lam[off + rank] = mine;
    for (int c = 0; c < sz; ++c) {
        const int rc = __shfl_sync(0xffffffffu, rank, c);
        if (lane < sz) Q[(size_t)(off + rc) * ld + off + lane] = z[lane * 33 + c];
    }
}

// ---------------------------------------------------------------------------------------------
// deflation: one CTA (256 threads) per merge node
// ---------------------------------------------------------------------------------------------
// dynamic shared memory: the merged pole list (d, z, source index, type) of the node -- the deflation scan and the map
// construction are serial chains over it (one thread), which must not run on global-memory latency
static inline size_t dc_deflate_smem_bytes(int mmax) { return (size_t)mmax * (8 + 8 + 4 + 4) + 64; }
#define DC_DEFLATE_SMEM_MAX_M 9600        // merges above this many poles (the Netflix tail: users up to 17,653 items) keep the lists in global memory
__global__ void __launch_bounds__(256) dc_deflate_kernel(DcParams P, int lists_in_global) {
    extern __shared__ __align__(16) unsigned char dfl_smem[];
    const DcNode nd = P.nodes[P.node0 + blockIdx.x];
    DcState& st = P.state[P.node0 + blockIdx.x];
    const HJob jb = P.jobs[nd.job];
    const int off = nd.off, n1 = nd.n1, m = nd.n1 + nd.n2, ld = jb.np, tid = threadIdx.x;
    const size_t vb = (size_t)jb.r_off + off;
    double* Qin = (P.in_b ? P.Qb : P.Qa) + jb.m_off + (size_t)off * ld + off;      // block (off, off)
    const double* lamIn = (P.in_b ? P.lamB : P.lamA) + vb;
    double* ds = P.ds + vb; double* zs = P.zs + vb; int32_t* src = P.src + vb;
    // lists_in_global: the node's own global arrays serve as the working copy (pos_df, written by dc_rank later, holds the types);
    // the serial scans then run through L1 / L2 -- a few ms on the one or two merges of a level that are this big
    double* s_ds = lists_in_global ? ds : (double*)dfl_smem;                                   // [m]
    double* s_zs = lists_in_global ? zs : s_ds + m;                                            // [m]
    int32_t* s_src = lists_in_global ? src : (int32_t*)(s_zs + m);                             // [m]
    int32_t* s_type = lists_in_global ? P.pos_df + vb : s_src + m;                             // [m] 1 top, 2 dense, 3 bottom
    const double beta = P.e[jb.r_off + off + n1 - 1];
    const double sgn = beta >= 0.0 ? 1.0 : -1.0;
    const double rho = 2.0 * fabs(beta);
    __shared__ double red[16];
    __shared__ int nrot_s;
    // (a) merge-sort the two sorted pole lists, gather z
    double dmax = 0.0, zmax = 0.0;
    for (int i = tid; i < m; i += 256) {
        const double v = lamIn[i];
        int lo, hi, rank;
        if (i < n1) {   // #{d2 < v}
            lo = n1; hi = m;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (lamIn[mid] < v) lo = mid + 1; else hi = mid; }
            rank = i + (lo - n1);
        } else {        // #{d1 <= v}
            lo = 0; hi = n1;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (lamIn[mid] <= v) lo = mid + 1; else hi = mid; }
            rank = (i - n1) + lo;
        }
        const double zi = (i < n1) ? Qin[(size_t)i * ld + (n1 - 1)] : sgn * Qin[(size_t)i * ld + n1];
        const double zz = zi * 0.70710678118654752440;
        s_ds[rank] = v; s_zs[rank] = zz; s_src[rank] = i; s_type[rank] = (i < n1) ? 1 : 3;
        src[rank] = i;
        dmax = fmax(dmax, fabs(v)); zmax = fmax(zmax, fabs(zz));
    }
    // CTA max
    for (int o = 16; o > 0; o >>= 1) {
        dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
        zmax = fmax(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
    }
    if ((tid & 31) == 0) { red[tid >> 5] = dmax; red[8 + (tid >> 5)] = zmax; }
    __syncthreads();
    dmax = 0.0; zmax = 0.0;
    for (int w = 0; w < 8; ++w) { dmax = fmax(dmax, red[w]); zmax = fmax(zmax, red[8 + w]); }
    const double tol = 8.0 * DC_EPS * fmax(dmax, zmax);
    // (b) serial deflation scan (thread 0, on the shared-memory copy).  zs[i] = 0 marks a deflated pole afterwards.
    double* rot = P.rot + 4 * vb;
    if (tid == 0) {
        int nrot = 0;
        if (rho * zmax <= tol) {
            for (int i = 0; i < m; ++i) s_zs[i] = 0.0;
        } else {
            int pj = -1;
            for (int i = 0; i < m; ++i) {
                if (rho * fabs(s_zs[i]) <= tol) { s_zs[i] = 0.0; continue; }
                if (pj >= 0) {
                    double s = s_zs[pj], c = s_zs[i];
                    const double tt = hypot(c, s);
                    const double t = s_ds[i] - s_ds[pj];
                    c /= tt; s = -s / tt;
                    if (fabs(t * c * s) <= tol) {
                        s_zs[i] = tt; s_zs[pj] = 0.0;
                        const double t2 = s_ds[pj] * c * c + s_ds[i] * s * s;
                        s_ds[i] = s_ds[pj] * s * s + s_ds[i] * c * c;
                        s_ds[pj] = t2;
                        if (s_type[pj] != s_type[i]) { s_type[i] = 2; s_type[pj] = 2; }
                        rot[4 * nrot] = c; rot[4 * nrot + 1] = s;
                        rot[4 * nrot + 2] = (double)s_src[pj]; rot[4 * nrot + 3] = (double)s_src[i];
                        ++nrot;
                    }
                }
                pj = i;
            }
        }
        nrot_s = nrot;
    }
    __syncthreads();
    if (!lists_in_global)
        for (int i = tid; i < m; i += 256) { ds[i] = s_ds[i]; zs[i] = s_zs[i]; }   // the global copies other kernels read
    // (c) apply the rotations to the columns of Qin, in order
    const int nrot = nrot_s;
    for (int r = 0; r < nrot; ++r) {
        const double c = rot[4 * r], s = rot[4 * r + 1];
        double* x = Qin + (size_t)((int)rot[4 * r + 2]) * ld;
        double* y = Qin + (size_t)((int)rot[4 * r + 3]) * ld;
        for (int i = tid; i < m; i += 256) {
            const double xv = x[i], yv = y[i];
            x[i] = c * xv + s * yv;
            y[i] = c * yv - s * xv;
        }
        __syncthreads();
    }
    // (d) maps (thread 0): secular order s, GEMM order g (type 1, 2, 3), deflated list
    if (tid == 0) {
        int k1 = 0, k2 = 0, k3 = 0;
        for (int i = 0; i < m; ++i)
            if (s_zs[i] != 0.0) { const int t = s_type[i]; k1 += (t == 1); k2 += (t == 2); k3 += (t == 3); }
        const int k = k1 + k2 + k3;
        int g1 = 0, g2 = k1, g3 = k1 + k2, s = 0, t = 0;
        double* dk = P.dk + vb; double* zk = P.zk + vb; double* dval = P.dval + vb;
        int32_t* colsrc = P.colsrc + vb; int32_t* dsrc = P.dsrc + vb; int32_t* gmap = P.gmap + vb;
        for (int i = 0; i < m; ++i) {
            if (s_zs[i] != 0.0) {
                const int ty = s_type[i];
                const int g = (ty == 1) ? g1++ : (ty == 2 ? g2++ : g3++);
                dk[s] = s_ds[i]; zk[s] = s_zs[i]; colsrc[g] = s_src[i];
                gmap[s] = g;
                ++s;
            } else { dsrc[t] = s_src[i]; dval[t] = s_ds[i]; ++t; }
        }
        st.k = k; st.k1 = k1; st.k2 = k2; st.k3 = k3; st.kneed = k; st.rho = rho;
    }
}

// ---------------------------------------------------------------------------------------------
// secular equation: grid (nnodes, ceil(mmax * LPR / 128)), block 128.  LPR lanes (1, 8 or 32, adjacent)
// share a root and split the poles, so that the few big merges near the top of the tree still fill
// the machine; every sum is reduced in a fixed order.
// ---------------------------------------------------------------------------------------------
template <int LPR>
__global__ void __launch_bounds__(128) dc_secular_kernel(DcParams P) {
    constexpr int RPC = 128 / LPR;                           // roots per CTA
    const DcNode nd = P.nodes[P.node0 + blockIdx.x];
    const DcState st = P.state[P.node0 + blockIdx.x];
    const int k = st.k;
    if ((int)blockIdx.y * RPC >= k) return;
    const HJob jb = P.jobs[nd.job];
    const size_t vb = (size_t)jb.r_off + nd.off;
    const double* dk = P.dk + vb;
    const double* zk = P.zk + vb;
    const int tid = threadIdx.x, sub = tid % LPR, j = blockIdx.y * RPC + tid / LPR;
    const bool active = j < k;
    const double rho = st.rho, rhoinv = 1.0 / rho;
    __shared__ double sd[128], sz2[128];
    if (k == 1) {
        if (j == 0 && sub == 0) { P.orig[vb] = 0; P.tau[vb] = rho * zk[0] * zk[0]; P.lamk[vb] = dk[0] + rho * zk[0] * zk[0]; }
        return;
    }
    const bool last = (j == k - 1);
    const int jc = active ? j : k - 1;
    const double dj = dk[jc];
    // upper end of the root's interval: next pole, or d_k + rho * ||z||^2 (z has unit norm) for the last root
    const double gap = (last || !active) ? rho : dk[jc + 1] - dj;

    // evaluates psi (poles <= jc), phi (poles > jc) and derivatives at lam = origin + tau
    auto evaluate = [&](double origin, double tau, double& psi, double& phi, double& dpsi, double& dphi) {
        psi = 0.0; phi = 0.0; dpsi = 0.0; dphi = 0.0;
        for (int c0 = 0; c0 < k; c0 += 128) {
            __syncthreads();
            if (c0 + tid < k) { sd[tid] = dk[c0 + tid]; const double z = zk[c0 + tid]; sz2[tid] = z * z; }
            __syncthreads();
            const int cnt = min(128, k - c0);
            for (int i = sub; i < cnt; i += LPR) {
                const double inv = 1.0 / ((sd[i] - origin) - tau);
                const double t = sz2[i] * inv;
                const double t2 = t * inv;
                if (c0 + i <= jc) { psi += t; dpsi += t2; } else { phi += t; dphi += t2; }
            }
        }
#pragma unroll
        for (int o = 1; o < LPR; o <<= 1) {
            psi += __shfl_xor_sync(0xffffffffu, psi, o);
            phi += __shfl_xor_sync(0xffffffffu, phi, o);
            dpsi += __shfl_xor_sync(0xffffffffu, dpsi, o);
            dphi += __shfl_xor_sync(0xffffffffu, dphi, o);
        }
    };

    double psi, phi, dpsi, dphi;
    const double mid = 0.5 * gap;
    evaluate(dj, mid, psi, phi, dpsi, dphi);
    const double wmid = rhoinv + psi + phi;
    const bool use_left = (wmid >= 0.0) || last;
    const int og = use_left ? jc : jc + 1;
    const double origin = use_left ? dj : dk[min(jc + 1, k - 1)];
    double lo, hi, tau;
    if (last) {
        if (wmid >= 0.0) { lo = 0.0; hi = mid; tau = 0.5 * mid; } else { lo = mid; hi = gap; tau = 0.75 * gap; }
    } else if (use_left) { lo = 0.0; hi = mid; tau = 0.5 * mid; }
    else { lo = -mid; hi = 0.0; tau = -0.5 * mid; }
    const double p1 = dj - origin;                           // pole below the root, relative to the origin
    const double p2 = last ? 0.0 : (dk[min(jc + 1, k - 1)] - origin);
    bool done = !active;
    for (int it = 0; it < 100; ++it) {
        if (__syncthreads_and(done)) break;
        evaluate(origin, tau, psi, phi, dpsi, dphi);
        if (done) continue;
        const double w = rhoinv + psi + phi;
        const double erretm = 8.0 * (fabs(psi) + fabs(phi)) + rhoinv + fabs(tau) * (dpsi + dphi);
        if (fabs(w) <= DC_EPS * erretm) { done = true; continue; }
        if (w < 0.0) lo = fmax(lo, tau); else hi = fmin(hi, tau);
        const double D1 = p1 - tau, D2 = p2 - tau;
        double eta;
        if (last) {
            const double c1 = rhoinv + psi - dpsi * D1;
            eta = D1 + dpsi * D1 * D1 / c1;
        } else {
            const double c = w - D1 * dpsi - D2 * dphi;
            const double a = (D1 + D2) * w - D1 * D2 * (dpsi + dphi);
            const double b = D1 * D2 * w;
            if (c == 0.0) eta = b / a;
            else {
                const double disc = sqrt(fabs(a * a - 4.0 * b * c));
                eta = (a <= 0.0) ? (a - disc) / (2.0 * c) : 2.0 * b / (a + disc);
            }
        }
        if (!isfinite(eta) || w * eta >= 0.0) eta = -w / (dpsi + dphi);
        double cand = tau + eta;
        if (!isfinite(cand) || cand <= lo || cand >= hi) cand = 0.5 * (lo + hi);
        if ((hi - lo) <= 4.0 * DC_EPS * fmax(fabs(lo), fabs(hi))) { done = true; continue; }
        tau = cand;
    }
    if (active && sub == 0) { P.orig[vb + j] = og; P.tau[vb + j] = tau; P.lamk[vb + j] = origin + tau; }
}

// ---------------------------------------------------------------------------------------------
// output order + cutoff: one CTA (256) per node
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dc_rank_kernel(DcParams P) {
    const DcNode nd = P.nodes[P.node0 + blockIdx.x];
    DcState& st = P.state[P.node0 + blockIdx.x];
    const HJob jb = P.jobs[nd.job];
    const int m = nd.n1 + nd.n2, k = st.k, nd_ = m - k, tid = threadIdx.x;
    const size_t vb = (size_t)jb.r_off + nd.off;
    const double* lamk = P.lamk + vb;
    const double* dval = P.dval + vb;
    double* lamOut = (P.in_b ? P.lamA : P.lamB) + vb;
    int32_t* pos_nd = P.pos_nd + vb; int32_t* pos_df = P.pos_df + vb;
    __shared__ int cnt_s, need_s;
    if (tid == 0) { cnt_s = 0; need_s = 0; }
    __syncthreads();
    double thr = 1e300;
    if (nd.final_) thr = (double)__double2float_rn(__dadd_rn((double)__uint_as_float(P.sigmax[nd.job]), 0.01));
    int below = 0;
    // total order: (value, deflated before root, index)
    for (int s = tid; s < k; s += 256) {
        const double v = lamk[s];
        int c = 0;
        for (int t = 0; t < nd_; ++t) c += (dval[t] <= v) ? 1 : 0;
        pos_nd[s] = s + c;
        lamOut[s + c] = v;
        below += !(v > thr);
    }
    for (int t = tid; t < nd_; t += 256) {
        const double v = dval[t];
        int c = 0;
        for (int u = 0; u < nd_; ++u) { const double o = dval[u]; c += (o < v || (o == v && u < t)) ? 1 : 0; }
        // roots below v: lamk ascending -> binary search #{lamk < v}
        int lo = 0, hi = k;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (lamk[mid] < v) lo = mid + 1; else hi = mid; }
        pos_df[t] = c + lo;
        lamOut[c + lo] = v;
        below += !(v > thr);
    }
    if (!nd.final_) return;
    if (below) atomicAdd(&cnt_s, below);
    __syncthreads();
    const int lim = min(max(cnt_s, 2), m);
    int need = 0;
    for (int s = tid; s < k; s += 256) need += (pos_nd[s] < lim) ? 1 : 0;
    if (need) atomicAdd(&need_s, need);
    __syncthreads();
    if (tid == 0) { st.kneed = need_s; P.kuser[nd.job] = lim; }
}

// ---------------------------------------------------------------------------------------------
// Loewner z-hat: grid (nnodes, ceil(mmax/8)), block 256: warp per pole i
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dc_zhat_kernel(DcParams P) {
    const DcNode nd = P.nodes[P.node0 + blockIdx.x];
    const DcState st = P.state[P.node0 + blockIdx.x];
    const int k = st.k;
    const int i = blockIdx.y * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= k) return;
    const HJob jb = P.jobs[nd.job];
    const size_t vb = (size_t)jb.r_off + nd.off;
    const double* dk = P.dk + vb;
    const double* tau = P.tau + vb;
    const int32_t* orig = P.orig + vb;
    const double di = dk[i];
    double prod = 1.0;
    for (int j = lane; j < k; j += 32) {
        const double num = (di - dk[orig[j]]) - tau[j];      // d_i - lam_j
        prod *= (j == i) ? num : num / (di - dk[j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) prod *= __shfl_xor_sync(0xffffffffu, prod, o);
    if (lane == 0) P.zhat[vb + i] = copysign(sqrt(fabs(prod)), P.zk[vb + i]);
}

// ---------------------------------------------------------------------------------------------
// S[g(i), j] = zhat_i / (d_i - lam_j) / norm: grid (nnodes, ceil(mmax/8)), block 256: warp per root j
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dc_vectors_kernel(DcParams P) {
    const DcNode nd = P.nodes[P.node0 + blockIdx.x];
    const DcState st = P.state[P.node0 + blockIdx.x];
    const int k = st.k;
    const int j = blockIdx.y * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= st.kneed) return;
    const HJob jb = P.jobs[nd.job];
    const size_t vb = (size_t)jb.r_off + nd.off;
    const double* dk = P.dk + vb;
    const double* zh = P.zhat + vb;
    const int32_t* gmap = P.gmap + vb;
    double* Scol = P.S + jb.m_off + (size_t)(nd.off + j) * jb.np + nd.off;
    if (k == 1) { if (lane == 0) Scol[0] = 1.0; return; }
    const double oj = dk[P.orig[vb + j]], tj = P.tau[vb + j];
    double ss = 0.0;
    for (int i = lane; i < k; i += 32) {
        const double v = zh[i] / ((dk[i] - oj) - tj);
        ss = fma(v, v, ss);
    }
    ss = warp_sum_d(ss);
    const double inv = 1.0 / sqrt(ss);
    for (int i = lane; i < k; i += 32) Scol[gmap[i]] = zh[i] / ((dk[i] - oj) - tj) * inv;
}

// ---------------------------------------------------------------------------------------------
// deflated columns: grid (nnodes, ceil(mmax/8)), block 256: warp per deflated column
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dc_copy_kernel(DcParams P) {
    const DcNode nd = P.nodes[P.node0 + blockIdx.x];
    const DcState st = P.state[P.node0 + blockIdx.x];
    const int m = nd.n1 + nd.n2, ndf = m - st.k;
    const int t = blockIdx.y * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (t >= ndf) return;
    const HJob jb = P.jobs[nd.job];
    const size_t vb = (size_t)jb.r_off + nd.off;
    const int pos = P.pos_df[vb + t];
    if (nd.final_ && pos >= P.kuser[nd.job]) return;
    const int ld = jb.np;
    const double* src = (P.in_b ? P.Qb : P.Qa) + jb.m_off + (size_t)(nd.off + P.dsrc[vb + t]) * ld + nd.off;
    double* dst = (P.in_b ? P.Qa : P.Qb) + jb.m_off + (size_t)(nd.off + pos) * ld + nd.off;
    for (int i = lane; i < m; i += 32) dst[i] = src[i];
}

// ---------------------------------------------------------------------------------------------
// GEMM  Qout[rows, pos_j] = sum_g Qin[rows, colsrc[g]] * S[g, j]   (two tasks per node: top, bottom)
// ---------------------------------------------------------------------------------------------
#define DCG_BM 128
#define DCG_BN 64
#define DCG_BK 16
#define DCG_SA (DCG_BM + 4)
#define DCG_SB (DCG_BK + 4)
#define DCG_STAGES 3
#define DCG_STAGE_DBL (DCG_BK * DCG_SA + DCG_BN * DCG_SB)

// tile prefix sums of the level: one CTA.  tile_off[2*nnodes + 1]
__global__ void __launch_bounds__(1024) dc_plan_kernel(DcParams P, int32_t* __restrict__ tile_off) {
    __shared__ int s[1024];
    const int ntask = 2 * P.nnodes, tid = threadIdx.x;
    const int per = (ntask + 1023) / 1024;
    const int b = min(ntask, tid * per), e = min(ntask, b + per);
    int acc = 0;
    for (int t = b; t < e; ++t) {
        const DcNode nd = P.nodes[P.node0 + (t >> 1)];
        const DcState st = P.state[P.node0 + (t >> 1)];
        const int M = (t & 1) ? nd.n2 : nd.n1, N = st.kneed;
        if (!(P.tc_node && P.tc_node[P.node0 + (t >> 1)])) acc += ((M + DCG_BM - 1) / DCG_BM) * ((N + DCG_BN - 1) / DCG_BN);
    }
    s[tid] = acc;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        int x = 0;
        if (tid >= o) x = s[tid - o];
        __syncthreads();
        s[tid] += x;
        __syncthreads();
    }
    int run = s[tid] - acc;
    for (int t = b; t < e; ++t) {
        tile_off[t] = run;
        const DcNode nd = P.nodes[P.node0 + (t >> 1)];
        const DcState st = P.state[P.node0 + (t >> 1)];
        const int M = (t & 1) ? nd.n2 : nd.n1, N = st.kneed;
        if (!(P.tc_node && P.tc_node[P.node0 + (t >> 1)])) run += ((M + DCG_BM - 1) / DCG_BM) * ((N + DCG_BN - 1) / DCG_BN);
    }
    if (tid == 1023) tile_off[ntask] = s[1023];
}

// The same product as dc_gemm_kernel below, described for the tcgen05 engine: Qout[rows, pos_nd[j]] = Qin[rows, colsrc[kb .. ke)] *
// S[kb .. ke, j], j < kneed.  One thread per task; entry `nt` is the engine's sentinel.
__global__ void dc_tc_resolve_kernel(DcParams P, const DcTcTask* __restrict__ ht, int nt, int8_t* __restrict__ planes,
                                     int32_t* __restrict__ expo, TcTask* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > nt) return;
    TcTask T;
    memset(&T, 0, sizeof T);
    if (t < nt) {
        const DcTcTask h = ht[t];
        const DcNode nd = P.nodes[h.node];
        const DcState st = P.state[h.node];
        const HJob jb = P.jobs[nd.job];
        const bool bottom = h.bottom != 0;
        const int M = bottom ? nd.n2 : nd.n1;
        const int kb = bottom ? st.k1 : 0, ke = bottom ? st.k : st.k1 + st.k2;
        const int rowbeg = bottom ? nd.n1 : 0, ld = jb.np;
        const size_t vb = (size_t)jb.r_off + nd.off;
        const size_t blk = (size_t)jb.m_off + (size_t)nd.off * ld + nd.off;
        T.A = (P.in_b ? P.Qb : P.Qa) + blk + rowbeg; T.lda = ld; T.gather = P.colsrc + vb + kb;
        T.B = P.S + blk + kb; T.ldb = ld;
        T.C = (P.in_b ? P.Qa : P.Qb) + blk + rowbeg; T.ldc = ld; T.scatter = P.pos_nd + vb;
        T.M = M; T.N = st.kneed; T.K = max(ke - kb, 0);
        T.Ap = planes + h.ap_off; T.Bp = planes + h.bp_off; T.ea = expo + h.e_off; T.eb = T.ea + M;
    }
    out[t] = T;
}

__global__ void __launch_bounds__(256, 2) dc_gemm_kernel(DcParams P, const int32_t* __restrict__ tile_off) {
    extern __shared__ __align__(16) double dcg_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2;        // warp tile: rows [32 wm, +32), cols [32 wn, +32)
    const int fk = lane & 3, fr = lane >> 2;
    const int ntask = 2 * P.nnodes;
    const int total = tile_off[ntask];
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        int lo = 0, hi = ntask - 1;                 // last task with tile_off[task] <= tile
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (tile_off[mid] <= tile) lo = mid; else hi = mid - 1; }
        const int task = lo, local = tile - tile_off[task];
        const DcNode nd = P.nodes[P.node0 + (task >> 1)];
        const DcState st = P.state[P.node0 + (task >> 1)];
        const HJob jb = P.jobs[nd.job];
        const bool bottom = task & 1;
        const int M = bottom ? nd.n2 : nd.n1, N = st.kneed;
        const int kb = bottom ? st.k1 : 0, ke = bottom ? st.k : st.k1 + st.k2;     // K range in GEMM order
        const int rowbeg = bottom ? nd.n1 : 0, ld = jb.np;
        const int tiles_m = (M + DCG_BM - 1) / DCG_BM;
        const int m0 = (local % tiles_m) * DCG_BM, n0 = (local / tiles_m) * DCG_BN;
        const size_t vb = (size_t)jb.r_off + nd.off;
        const double* Qin = (P.in_b ? P.Qb : P.Qa) + jb.m_off + (size_t)nd.off * ld + nd.off + rowbeg + m0;
        double* Qout = (P.in_b ? P.Qa : P.Qb) + jb.m_off + (size_t)nd.off * ld + nd.off + rowbeg + m0;
        const double* Sblk = P.S + jb.m_off + (size_t)(nd.off + n0) * ld + nd.off;
        const int32_t* colsrc = P.colsrc + vb;
        const int32_t* pos_nd = P.pos_nd + vb;
        const int nk = (ke - kb + DCG_BK - 1) / DCG_BK;

        auto load_stage = [&](int kt, int stg) {
            double* As = dcg_smem + (size_t)stg * DCG_STAGE_DBL;
            double* Bs = As + DCG_BK * DCG_SA;
            const int k0 = kb + kt * DCG_BK;
            // A: 16 k-columns x 128 rows, 16-byte chunks (rows are even-aligned)
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int ch = tid + it * 256, kk = ch >> 6, r2 = (ch & 63) * 2;
                const int g = k0 + kk;
                const bool ok = (g < ke) && (m0 + r2 < M);
                const double* src = ok ? Qin + (size_t)colsrc[g] * ld + r2 : Qin;
                cp_async16_zfill(As + kk * DCG_SA + r2, src, ok);
            }
            // B: 64 columns x 16 k, 8-byte copies (K ranges start anywhere)
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int el = tid + it * 256, nn = el >> 4, kk = el & 15;
                const int g = k0 + kk;
                const bool ok = (g < ke) && (n0 + nn < N);
                const double* src = ok ? Sblk + (size_t)nn * ld + g : Sblk;
                const int sz = ok ? 8 : 0;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem_u32(Bs + nn * DCG_SB + kk)), "l"(src), "r"(sz) : "memory");
            }
        };

        double acc[4][4][2];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }
        __syncthreads();       // previous tile's readers are done with the ring
        for (int s = 0; s < DCG_STAGES - 1; ++s) { if (s < nk) load_stage(s, s); cp_async_commit(); }
        for (int kt = 0; kt < nk; ++kt) {
            cp_async_wait<DCG_STAGES - 2>();
            __syncthreads();
            if (kt + DCG_STAGES - 1 < nk) load_stage(kt + DCG_STAGES - 1, (kt + DCG_STAGES - 1) % DCG_STAGES);
            cp_async_commit();
            const double* As = dcg_smem + (size_t)(kt % DCG_STAGES) * DCG_STAGE_DBL;
            const double* Bs = As + DCG_BK * DCG_SA;
#pragma unroll
            for (int k4 = 0; k4 < DCG_BK; k4 += 4) {
                double af[4], bf[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) af[a] = Bs[(32 * wn + 8 * a + fr) * DCG_SB + k4 + fk];      // mma rows = n
#pragma unroll
                for (int b = 0; b < 4; ++b) bf[b] = As[(k4 + fk) * DCG_SA + 32 * wm + 8 * b + fr];      // mma cols = m
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) dmma(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
            }
        }
        cp_async_wait<0>();
        // D[n][m]: lane holds rows m = 2 fk + {0,1} of column n = fr
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int nn = n0 + 32 * wn + 8 * a + fr;
            if (nn < N) {
                double* dst = Qout + (size_t)(pos_nd[nn] - 0) * ld;
                // column position is relative to the node block: Qout already points at (off, off + rowbeg + m0)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int mm = 32 * wm + 8 * b + 2 * fk;
                    if (m0 + mm + 1 < M) *(double2*)(dst + mm) = make_double2(acc[a][b][0], acc[a][b][1]);
                    else if (m0 + mm < M) dst[mm] = acc[a][b][0];
                }
            }
        }
    }
}
