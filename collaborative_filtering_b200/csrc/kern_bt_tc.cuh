// Large path, stage 3 on the tensor-core engine: the back-transformation U = H_0 ... H_{n-2} Z of the big users with
// aggregated compact-WY panels of 512 reflectors, so that both products of a panel have a long K and run FP64-equivalent
// on tcgen05 (tc_gemm.cu) instead of K = 64 on the FP64 pipe (kern_bt.cuh).  Per super-panel P = reflectors j0 .. j0+nb-1:
//     H_j0 ... H_j0+nb-1 = I - V T V^T,   V = n x nb unit lower trapezoid,  T = nb x nb upper triangular (dlarft, forward / columnwise)
//   up front, for every super-panel of every big user at once
//     btt_extract_kernel   V of the whole user as a plain column-major matrix (from the tile-major reflector storage) in the Q
//                          buffer the divide & conquer left free
//     G = V^T V            (tc_gemm, K = rows)
//     btt_formt_kernel     T from G and tau, in place: 64 x 64 diagonal blocks by the column recurrence, block column q above the
//                          diagonal as  T(0:q, q) = -T(0:q, 0:q) (G(0:q, q) T_qq)  -- 64^3 block products in shared memory
//     VT = V T             (tc_gemm, K = nb) into the S buffer
//   then, last super-panel first (every user's s-th panel from the end in one batch):
//     X = V^T Z            (tc_gemm, M = nb, N = k, K = rows)
//     Z -= VT X            (tc_gemm, M = rows, N = k, K = nb)
// k (kept eigenvectors) is only known on the device: btt_set_n_kernel copies it into the task list.
#pragma once
#include "gsi_internal.cuh"
#include "kern_trd.cuh"
#include "kern_sbr.cuh"      // sbr_mm64
#include "tc_gemm.cuh"

#define BTT_NB 512

struct BttSp { int job, j0, nb, pad_; int64_t g_off; };      // g_off: doubles into the G / T buffer (512 x 512 per super-panel, ld 512)

// grid (NTmax * NTmax, nusers), block 256: tile (I, J) of user `users[blockIdx.y]` -> column-major Vfull (ld = np); tiles above the
// diagonal and columns without a reflector (j >= n - 1) are zero
__global__ void __launch_bounds__(256) btt_extract_kernel(const HJob* __restrict__ jobs, int nusers, int NTmax, const double* __restrict__ A,
                                                          double* __restrict__ Qa, double* __restrict__ Qb) {
    const HJob jb = jobs[blockIdx.y];
    const int NT = jb.np >> 6, I = blockIdx.x % NTmax, J = blockIdx.x / NTmax;
    if (I >= NT || J >= NT) return;
    double* V = ((jb.levels & 1) ? Qa : Qb) + jb.m_off;                      // the buffer that does NOT hold Z
    const double* tile = A + jb.m_off + (((size_t)J * NT + I) << 12);
    for (int e = threadIdx.x; e < 4096; e += 256) {
        const int c = e >> 6, r = e & 63, j = J * 64 + c;
        V[(size_t)j * jb.np + I * 64 + r] = (I >= J && j < jb.n - 1) ? tile[e] : 0.0;
    }
}

// tasks[t].N = kuser[task_job[t]] for the tasks whose N is the user's kept-eigenvector count (task_job >= 0)
__global__ void btt_set_n_kernel(TcTask* __restrict__ tasks, const int32_t* __restrict__ task_job, int ntasks, const int32_t* __restrict__ kuser) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ntasks && task_job[t] >= 0) tasks[t].N = kuser[task_job[t]];
}

// grid (super-panels), block 256, smem 3 x 64 x 65 doubles: G (512 x 512, ld 512) -> T in place
static inline size_t btt_formt_smem_bytes() { return (size_t)3 * 64 * 65 * sizeof(double); }
__global__ void __launch_bounds__(256) btt_formt_kernel(const HJob* __restrict__ jobs, const BttSp* __restrict__ sps, const double* __restrict__ tauall,
                                                        double* __restrict__ GT) {
    extern __shared__ __align__(16) double fsm[];
    const BttSp sp = sps[blockIdx.x];
    const HJob jb = jobs[sp.job];
    double* G = GT + sp.g_off;
    const double* tau = tauall + jb.r_off + sp.j0;
    double* As = fsm;                  // column-major 64 x 64 tiles, ld 65
    double* Bs = As + 64 * 65;
    double* Ts = Bs + 64 * 65;         // T_qq of the current block column
    const int tid = threadIdx.x, tr = tid >> 4, tc = tid & 15;
    const int nq = (sp.nb + 63) >> 6;
    auto load_tile = [&](double* dst, int br, int bc) {          // dst(r, c) = G(64 br + r, 64 bc + c)
        for (int e = tid; e < 4096; e += 256) { const int c = e >> 6, r = e & 63; dst[c * 65 + r] = G[(size_t)(64 * bc + c) * BTT_NB + 64 * br + r]; }
    };
    auto store_acc = [&](double (&acc)[4][4], int br, int bc, double scale) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) G[(size_t)(64 * bc + 4 * tc + b) * BTT_NB + 64 * br + 4 * tr + a] = scale * acc[a][b];
    };
    for (int q = 0; q < nq; ++q) {
        const int b = min(64, sp.nb - 64 * q);
        // ---- T_qq: columnwise recurrence on the diagonal block (bt_formt_kernel's loop)
        __syncthreads();
        load_tile(As, q, q);
        for (int e = tid; e < 64 * 65; e += 256) Ts[e] = 0.0;
        __syncthreads();
        for (int j = 0; j < b; ++j) {
            const double tj = tau[64 * q + j];
            if (tid < j) {
                double s = 0.0;
                for (int l = tid; l < j; ++l) s = fma(Ts[l * 65 + tid], As[j * 65 + l], s);
                Ts[j * 65 + tid] = -tj * s;
            } else if (tid == j) Ts[j * 65 + j] = tj;
            __syncthreads();
        }
        for (int e = tid; e < 4096; e += 256) { const int c = e >> 6, r = e & 63; G[(size_t)(64 * q + c) * BTT_NB + 64 * q + r] = Ts[c * 65 + r]; }
        // ---- H_r = G(r, q) T_qq, r < q, stored where G(r, q) was
        for (int r = 0; r < q; ++r) {
            __syncthreads();
            load_tile(As, r, q);
            __syncthreads();
            double acc[4][4] = {};
            sbr_mm64(As, 65, false, Ts, 65, acc, tr, tc);
            store_acc(acc, r, q, 1.0);
        }
        // ---- T(r, q) = - sum_{s = r}^{q-1} T(r, s) H_s, r ascending (H_r is consumed by its own row first)
        for (int r = 0; r < q; ++r) {
            double acc[4][4] = {};
            for (int s = r; s < q; ++s) {
                __syncthreads();
                load_tile(As, r, s);
                load_tile(Bs, s, q);
                __syncthreads();
                sbr_mm64(As, 65, false, Bs, 65, acc, tr, tc);
            }
            __syncthreads();                                    // everybody has read H_r before it is overwritten
            store_acc(acc, r, q, -1.0);
        }
        // ---- blocks below the diagonal of this block column: T is upper triangular
        for (int r = q + 1; r < 8; ++r)
            for (int e = tid; e < 4096; e += 256) { const int c = e >> 6, rr = e & 63; G[(size_t)(64 * q + c) * BTT_NB + 64 * r + rr] = 0.0; }
    }
    // columns beyond nb (a short last super-panel) stay as the memset left them
}
