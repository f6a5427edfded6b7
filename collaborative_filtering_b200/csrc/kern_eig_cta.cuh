// CTA-resident path for users with n <= 160 rated items: ONE kernel does the whole record --
//   gather W (precompute_local.cpp:185-192) -> degree (:196-208) -> normalised Laplacian
//   (:211-222) -> sig_min (:236-249) -> full symmetric eigensolve (:231-233) -> cutoff (:252-261)
// with the n x n matrix living in shared memory from the gather to the last store.
//
// Eigensolver: one-sided (Hestenes) Jacobi on G = sym(lower(L)) + I.  L's spectrum is in [0,2],
// so G is positive definite with spectrum in [1,3]: at convergence the columns of G V are
// orthogonal, ||g_j|| = lambda_j + 1 and g_j / ||g_j|| is the eigenvector.  Only ONE n x n array
// is needed (fp64: n <= 160 in 227 KB).  Pairs follow the round-robin ("circle") ordering:
// n/2 disjoint pairs per round, one warp per pair, lanes own rows {lane, lane+32, ...}.
#pragma once
#include "gsi_internal.cuh"

struct SParams {
    const double* W;
    int w_rows;
    const int32_t* items;         // all rated ids of the batch (device)
    const int64_t* job_item_off;  // per job: offset into items / sig_min
    const int32_t* job_n;
    const int64_t* job_vec_off;   // per job: offset into vec_pad (n * max(n,2) doubles reserved)
    const int64_t* job_lam_off;   // per job: offset into lam_pad (max(n,2) doubles reserved)
    double* sig_min;
    int32_t* job_k;
    int32_t* job_sweeps;
    double* lam_pad;
    double* vec_pad;
    int job_base;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ unsigned long long dbits(double x) { return (unsigned long long)__double_as_longlong(x); }

// pair `pi` of round `r` in the circle ordering over npad (even) slots
__device__ __forceinline__ void circle_pair(int npad, int r, int pi, int& p, int& q) {
    const int m = npad - 1;
    if (pi == 0) { p = m; q = r; }
    else { p = r + pi; if (p >= m) p -= m; q = r - pi; if (q < 0) q += m; }
    if (p > q) { int t = p; p = q; q = t; }
}

// Rotation that orthogonalises two columns with squared norms a, b and inner product g.
__device__ __forceinline__ void jacobi_cs(double a, double b, double g, double& c, double& s, double& t) {
    const double zeta = (b - a) / (2.0 * g);
    t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    c = 1.0 / sqrt(1.0 + t * t);
    s = c * t;
}

template <int EPT>
__global__ void __launch_bounds__(EPT <= 2 ? 256 : (EPT == 3 ? 512 : 1024))
eig_cta_kernel(SParams P) {
    extern __shared__ double smem[];
    const int job = blockIdx.x + P.job_base;
    const int n = P.job_n[job];
    const int ld = n | 1;                    // odd leading dimension: conflict-free both ways
    const int npad = n + (n & 1);
    double* G = smem;                        // n x n, column-major, G[i + j*ld]
    double* dg = G + (size_t)n * ld;         // degree, later eigenvalues
    double* sc = dg + npad;                  // D^-1/2, later sign/norm scale
    double* nrm = sc + npad;                 // squared column norms
    int* ids = (int*)(nrm + npad);
    int* perm = ids + npad;
    __shared__ unsigned int sh_sigmax;
    __shared__ unsigned long long sh_smax;
    __shared__ int sh_cnt;

    const int tid = threadIdx.x, T = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const int64_t ioff = P.job_item_off[job];

    for (int t = tid; t < n; t += T) ids[t] = P.items[ioff + t];
    if (tid == 0) { sh_sigmax = 0u; sh_cnt = 0; }
    __syncthreads();

    // ---- P3 gather: warp per row of W (one table row is contiguous), lanes over columns ----
    for (int i = warp; i < n; i += nwarps) {
        const unsigned mi = (unsigned)ids[i];
        const bool iok = mi < (unsigned)P.w_rows;
        const double* row = P.W + (size_t)mi * P.w_rows;
        for (int j = lane; j < n; j += 32) {
            const unsigned mj = (unsigned)ids[j];
            G[i + j * ld] = (iok && mj < (unsigned)P.w_rows) ? __ldg(row + mj) : 0.0;
        }
    }
    __syncthreads();
    // ---- P4 degree: sequential j order per row, exactly as the reference's scalar loop ----
    for (int i = tid; i < n; i += T) {
        double d = 0.0;
        for (int j = 0; j < n; ++j) d = __dadd_rn(d, G[i + j * ld]);
        if (d == 0.0) d = 1.0;
        dg[i] = d;
        sc[i] = __dsqrt_rn(__ddiv_rn(1.0, d));   // sqrt of the inverse, not 1/sqrt (:216-220)
    }
    __syncthreads();
    // ---- P5: ll2_ij = fl(fl(s_i * (dd_ij - ww_ij)) * s_j) ----
    for (int idx = tid; idx < n * n; idx += T) {
        const int i = idx % n, j = idx / n;
        const double w = G[i + j * ld];
        const double ll = (i == j) ? __dsub_rn(dg[i], w) : __dsub_rn(0.0, w);
        G[i + j * ld] = __dmul_rn(__dmul_rn(sc[i], ll), sc[j]);
    }
    __syncthreads();
    // ---- P7 sig_min: float accumulator, (float)((double)acc + x*x), full row, j ascending ----
    for (int i = tid; i < n; i += T) {
        float acc = 0.f;
        for (int j = 0; j < n; ++j) {
            const double x = G[i + j * ld];
            acc = __double2float_rn(__dadd_rn((double)acc, __dmul_rn(x, x)));
        }
        const float sig = __fsqrt_rn(acc);
        P.sig_min[ioff + i] = __dadd_rn((double)sig, 0.01);
        atomicMax(&sh_sigmax, __float_as_uint(sig));   // sig >= 0: uint order == float order
    }
    __syncthreads();
    const float sig_min_max = __double2float_rn(__dadd_rn((double)__uint_as_float(sh_sigmax), 0.01));
    // ---- solver input: lower triangle mirrored (SelfAdjointEigenSolver reads lower only), + I ----
    for (int idx = tid; idx < n * n; idx += T) {
        const int i = idx % n, j = idx / n;
        if (i < j) G[i + j * ld] = G[j + i * ld];
    }
    __syncthreads();
    for (int i = tid; i < n; i += T) G[i + i * ld] += 1.0;
    __syncthreads();

    // ---- one-sided Jacobi sweeps ----
    const int npairs = npad >> 1;
    int sweeps = 0;
    for (; sweeps < GSI_MAX_SWEEPS; ++sweeps) {
        for (int c = warp; c < n; c += nwarps) {
            double a = 0.0;
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                const int i = lane + 32 * e;
                if (i < n) { const double v = G[i + c * ld]; a = fma(v, v, a); }
            }
            a = warp_sum(a);
            if (lane == 0) nrm[c] = a;
        }
        if (tid == 0) sh_smax = 0ull;
        __syncthreads();
        double wmax = 0.0;
        for (int r = 0; r < npad - 1; ++r) {
            for (int pi = warp; pi < npairs; pi += nwarps) {
                int p, q;
                circle_pair(npad, r, pi, p, q);
                if (q >= n) continue;            // phantom slot of an odd n
                double gp[EPT], gq[EPT];
                double g = 0.0;
#pragma unroll
                for (int e = 0; e < EPT; ++e) {
                    const int i = lane + 32 * e;
                    gp[e] = (i < n) ? G[i + p * ld] : 0.0;
                    gq[e] = (i < n) ? G[i + q * ld] : 0.0;
                    g = fma(gp[e], gq[e], g);
                }
                g = warp_sum(g);
                const double a = nrm[p], b = nrm[q];
                const double ab = a * b, g2 = g * g;
                if (ab > 0.0) wmax = fmax(wmax, g2 / ab);
                if (g2 > GSI_ROT2 * ab) {
                    double c, s, t;
                    jacobi_cs(a, b, g, c, s, t);
#pragma unroll
                    for (int e = 0; e < EPT; ++e) {
                        const int i = lane + 32 * e;
                        if (i < n) {
                            G[i + p * ld] = c * gp[e] - s * gq[e];
                            G[i + q * ld] = s * gp[e] + c * gq[e];
                        }
                    }
                    if (lane == 0) { nrm[p] = a - t * g; nrm[q] = b + t * g; }
                }
            }
            __syncthreads();
        }
        if (lane == 0) atomicMax(&sh_smax, dbits(wmax));
        __syncthreads();
        const unsigned long long smax = sh_smax;
        __syncthreads();
        if (smax <= dbits(GSI_STOP2)) { ++sweeps; break; }
    }

    // ---- eigenvalues, sign convention (largest |component| positive, first on ties) ----
    for (int c = warp; c < n; c += nwarps) {
        double a = 0.0, best = -1.0;
        int arg = 0;
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const int i = lane + 32 * e;
            if (i < n) {
                const double v = G[i + c * ld];
                a = fma(v, v, a);
                if (fabs(v) > best) { best = fabs(v); arg = i; }
            }
        }
        a = warp_sum(a);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
        }
        if (lane == 0) {
            const double nr = sqrt(a);
            dg[c] = nr - 1.0;                                   // eigenvalue
            sc[c] = (G[arg + c * ld] < 0.0) ? -nr : nr;         // signed norm
        }
    }
    __syncthreads();
    // ---- ascending order by counting rank; P8 cutoff: lim = #{lambda <= sig_min_max}, >= 2 ----
    for (int c = tid; c < n; c += T) {
        const double v = dg[c];
        int rank = 0;
        for (int o = 0; o < n; ++o) {
            const double w = dg[o];
            rank += (w < v || (w == v && o < c)) ? 1 : 0;
        }
        perm[rank] = c;
        if (!(v > (double)sig_min_max)) atomicAdd(&sh_cnt, 1);
    }
    __syncthreads();
    const int k = max(sh_cnt, 2);
    if (tid == 0) { P.job_k[job] = k; P.job_sweeps[job] = sweeps; }
    double* lam = P.lam_pad + P.job_lam_off[job];
    double* vec = P.vec_pad + P.job_vec_off[job];
    for (int r = tid; r < k; r += T) lam[r] = (r < n) ? dg[perm[r]] : 0.0;
    for (int idx = tid; idx < n * k; idx += T) {
        const int i = idx / k, r = idx - i * k;
        double v = 0.0;
        if (r < n) { const int c = perm[r]; v = G[i + c * ld] / sc[c]; }
        vec[idx] = v;
    }
}

static inline size_t eig_cta_smem_bytes(int n) {
    const int ld = n | 1, npad = n + (n & 1);
    return ((size_t)n * ld + 3 * (size_t)npad) * sizeof(double) + 2 * (size_t)npad * sizeof(int);
}
