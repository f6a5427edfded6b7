// FP64-equivalent GEMM on the 5th-generation tensor cores: C = A * B with A, B, C in
// FP64, computed as exact INT8 slice products on tcgen05.mma.kind::i8 with INT32 accumulators in tensor memory (the
// Ozaki scheme).  It serves the GEMM-rich stages of the eigensolve that replaces precompute_local.cpp:231 -- first the
// divide-and-conquer merges (dc_gemm_kernel, K = merged size) -- where the FP64 pipe (DMMA m8n8k4, 36 TF/s) is the bound.
//
//   slicing   row i of A (column j of B) is scaled by a power of two so that |x| < 1/2, rounded ONCE to a P = 7 S bit
//             fixed-point integer X and cut into S balanced base-128 digits d_t in [-64, 63]:  X = sum_t d_t 128^(S-1-t).
//             The digits are written as INT8 planes in the tensor core's canonical K-major shared-memory layout, tile by
//             tile, so that ONE bulk-copy request (cp.async.bulk -> UBLKCP, completion on an mbarrier) moves all S planes of
//             a (128 x 64) A block or a (64 x 64) B block into shared memory.
//   product   A_ik B_kj = 2^(e_i + f_j - 2P) sum_{t,u} d_t(ik) g_u(kj) 128^(2S-2-t-u).  The slice pairs with t + u = g share a
//             weight, so they accumulate -- exactly, in INT32 -- into ONE tensor-memory accumulator per g: S accumulators of
//             128 lanes x 64 columns fill the 512 columns of tensor memory, one pass over K, S (S + 1) / 2 slice products per K
//             step of 32 (pairs with t + u >= S are below the rounding of X and are dropped).  The products of one A slice t
//             with its partners u = 0 .. S-1-t land in CONSECUTIVE accumulators, so they are issued as one MMA of
//             N = 64 (S - t) columns (cut at 256): 12 instructions per K step for S = 8 instead of 36, and a third of the
//             shared-memory reads of the A block.
//   epilogue  four warps read the S accumulators of their rows (tcgen05.ld), combine them by Horner in FP64
//             (acc = acc / 128 + G_g, exact conversions), scale by the row and column factors and write C.
//
// Error: each operand entry carries an absolute error <= 2^(e-P-1) (e = its row / column exponent), the dropped pairs add
// <= S 2^-7S relative to the row / column maxima; everything else is exact integer arithmetic.  For orthogonal-matrix
// blocks and normalised secular vectors (the D&C operands) S = 8 gives |dC| ~ K 2^-57 max|A_i.| max|B_.j| -- below the
// FP64 rounding of the DMMA kernel's own K-long sums.
//
// One CTA per SM (all of tensor memory), 192 threads: warp 0 = bulk-copy producer, warp 1 = MMA issuer (one elected
// lane each), warps 2-5 = epilogue (TMEM lane quadrant = warp % 4).  Persistent over the (m tile, n tile) list.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include "gsi_internal.cuh"
#include "ptx.cuh"
#include "tc_gemm.cuh"

// ---------------------------------------------------------------------------------------------------------------------
// tcgen05 / TMEM wrappers (sm_100a PTX)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
// D[tmem] (+)= A[smem desc] * B[smem desc], INT8 x INT8 -> INT32, issued by one thread for the CTA
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive once every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns: thread t of the warp gets lane (lane base + t)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// shared-memory matrix descriptor, K-major, no swizzle ("interleave"): core matrix = 8 rows x 16 bytes, contiguous (128 B);
// SBO = bytes between core matrices of consecutive 8-row groups, LBO = bytes between the two 16-byte K chunks of one MMA
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) |
           (1ull << 46);                                     // descriptor version 1 (sm_100); base offset 0, layout type 0
}
// instruction descriptor: D = S32 (bits 4-5 = 2), A = B = signed INT8 (bits 7-9, 10-12 = 1), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t umma_idesc_i8(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------------------------------
// slicing.  Plane storage (bytes):  A: [m tile][k block][slice][k chunk (4)][row group (16)][8 rows][16 k]   8 KB per slice
//                                   B: [n tile][k block][k chunk (4)][slice][col group ( 8)][8 cols][16 k]   4 KB per slice
//   (B: the slices of one k chunk are adjacent, so that 8-column groups of consecutive slices are 128 bytes apart: one MMA descriptor
//    spans several slices, see the issue loop)
// ---------------------------------------------------------------------------------------------------------------------
// exponent e with |x| 2^-e < 1/2 for every x of the line (max |x| = 0 -> e = 0)
__device__ __forceinline__ int tc_exponent(double amax) {
    if (!(amax > 0.0)) return 0;
    int e;
    frexp(amax, &e);                                          // amax = f 2^e, f in [1/2, 1)
    return e + 1;
}

// Every kernel works on a device-resident list of TcTask (tc_gemm.cuh): blockIdx.z (or a tile prefix sum for the GEMM) picks the
// task, whose sizes may have been written by an earlier kernel of the same stream (the D&C's deflation counts).  Grids are sized by
// host-side upper bounds; blocks beyond a task's real extent leave at once.

// The exponent of a line is the largest element exponent (frexp is monotone in |x|), so the scan is an integer atomicMax over
// coalesced tiles.  ea / eb start at TC_EXP_NONE (tc_exp_init_kernel); a line of zeros keeps it and is treated as e = 0.
#define TC_EXP_NONE (-100000)
__device__ __forceinline__ int tc_elem_exp(double x) {
    if (!(fabs(x) > 0.0)) return TC_EXP_NONE;
    int e;
    frexp(x, &e);
    return e + 1;
}
__device__ __forceinline__ int tc_line_exp(int e) { return e == TC_EXP_NONE ? 0 : e; }

// grid (ceil((Mmax + Nmax) / 256), 1, tasks)
__global__ void __launch_bounds__(256) tc_exp_init_kernel(const TcTask* __restrict__ tasks) {
    const TcTask T = tasks[blockIdx.z];
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < T.M) T.ea[i] = TC_EXP_NONE;
    else if (i - T.M < T.N) T.eb[i - T.M] = TC_EXP_NONE;
}

// grid (k blocks max, m tiles max, tasks), block 256: thread = (row r = tid & 127, k half), 32 k's each
__global__ void __launch_bounds__(256) tc_row_exp_kernel(const TcTask* __restrict__ tasks) {
    const TcTask T = tasks[blockIdx.z];
    const int kb = blockIdx.x, i = blockIdx.y * 128 + (threadIdx.x & 127), half = threadIdx.x >> 7;
    if (kb * 64 >= T.K || i >= T.M || (T.flags & TC_UNIT_A)) return;
    int e = TC_EXP_NONE;
    for (int q = 0; q < 32; ++q) {
        const int k = kb * 64 + half * 32 + q;
        if (k < T.K) e = max(e, tc_elem_exp((T.flags & TC_TRANS_A) ? T.A[(size_t)i * T.lda + k] : T.A[(size_t)(T.gather ? T.gather[k] : k) * T.lda + i]));
    }
    if (e != TC_EXP_NONE) atomicMax(T.ea + i, e);
}

// grid (k blocks max, n tiles max, tasks), block 256: thread = (k chunk = tid >> 6, column = tid & 63), 16 consecutive k
__global__ void __launch_bounds__(256) tc_col_exp_kernel(const TcTask* __restrict__ tasks) {
    const TcTask T = tasks[blockIdx.z];
    const int kb = blockIdx.x, j = blockIdx.y * 64 + (threadIdx.x & 63), chunk = threadIdx.x >> 6;
    if (kb * 64 >= T.K || j >= T.N || (T.flags & TC_UNIT_B)) return;
    int e = TC_EXP_NONE;
    const double* col = T.B + (size_t)j * T.ldb;
    for (int q = 0; q < 16; ++q) {
        const int k = kb * 64 + chunk * 16 + q;
        if (k < T.K) e = max(e, tc_elem_exp(col[k]));
    }
    if (e != TC_EXP_NONE) atomicMax(T.eb + j, e);
}

template <int S>
__device__ __forceinline__ void tc_digits(double x, int e, int8_t (&d)[S]) {
    // X = round(x 2^(7S - e)), |X| <= 2^(7S-1);  balanced base-128 digits, least significant first
    long long X = __double2ll_rn(ldexp(x, 7 * S - e));
    const long long lim = (1ll << (7 * S - 1)) - 1;
    X = X > lim ? lim : (X < -lim ? -lim : X);
#pragma unroll
    for (int t = S - 1; t >= 1; --t) {
        const int r = (int)(((X + 64) & 127) - 64);            // in [-64, 63], X - r divisible by 128
        d[t] = (int8_t)r;
        X = (X - r) >> 7;
    }
    d[0] = (int8_t)X;                                          // what is left: |X| <= 65
}

// grid (k blocks max, m tiles max, tasks), block 256: one (128 x 64) block of A -> S planes of 8 KB.  Thread = (row r = tid & 127,
// k half); every thread handles 32 k's of its row; reads are coalesced over rows.
template <int S>
__global__ void __launch_bounds__(256) tc_slice_a_kernel(const TcTask* __restrict__ tasks) {
    const TcTask T = tasks[blockIdx.z];
    const int nkb = (T.K + 63) >> 6, kb = blockIdx.x, mt = blockIdx.y, r = threadIdx.x & 127, half = threadIdx.x >> 7;
    if (kb >= nkb || mt * 128 >= T.M) return;
    const int i = mt * 128 + r;
    int8_t* base = T.Ap + ((size_t)mt * nkb + kb) * (size_t)(S * 8192);
    const int e = (T.flags & TC_UNIT_A) ? 1 : ((i < T.M) ? tc_line_exp(T.ea[i]) : 0);
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {                                // two 16-byte k chunks per thread
        const int chunk = half * 2 + c;
        __align__(16) int8_t dig[S][16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int k = kb * 64 + chunk * 16 + q;
            double x = 0.0;
            if (i < T.M && k < T.K) x = (T.flags & TC_TRANS_A) ? T.A[(size_t)i * T.lda + k] : T.A[(size_t)(T.gather ? T.gather[k] : k) * T.lda + i];
            int8_t d[S];
            tc_digits<S>(x, e, d);
#pragma unroll
            for (int t = 0; t < S; ++t) dig[t][q] = d[t];
        }
#pragma unroll
        for (int t = 0; t < S; ++t) {
            int8_t* dst = base + (size_t)t * 8192 + chunk * 2048 + (r >> 3) * 128 + (r & 7) * 16;
            *(int4*)dst = *(const int4*)dig[t];
        }
    }
}

// grid (k blocks max, n tiles max, tasks), block 256: one (64 k x 64 columns) block of B -> S planes of 4 KB.  Thread = (k chunk =
// tid >> 6, column = tid & 63): 16 consecutive k of one column (contiguous in memory: B is column-major K x N)
template <int S>
__global__ void __launch_bounds__(256) tc_slice_b_kernel(const TcTask* __restrict__ tasks) {
    const TcTask T = tasks[blockIdx.z];
    const int nkb = (T.K + 63) >> 6, kb = blockIdx.x, nt = blockIdx.y, chunk = threadIdx.x >> 6, c = threadIdx.x & 63;
    if (kb >= nkb || nt * 64 >= T.N) return;
    const int j = nt * 64 + c;
    int8_t* base = T.Bp + ((size_t)nt * nkb + kb) * (size_t)(S * 4096);
    const int e = (T.flags & TC_UNIT_B) ? 1 : ((j < T.N) ? tc_line_exp(T.eb[j]) : 0);
    __align__(16) int8_t dig[S][16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int k = kb * 64 + chunk * 16 + q;
        double x = 0.0;
        if (j < T.N && k < T.K) x = T.B[(size_t)j * T.ldb + k];
        int8_t d[S];
        tc_digits<S>(x, e, d);
#pragma unroll
        for (int t = 0; t < S; ++t) dig[t][q] = d[t];
    }
#pragma unroll
    for (int t = 0; t < S; ++t) {
        int8_t* dst = base + (size_t)chunk * (S * 1024) + t * 1024 + (c >> 3) * 128 + (c & 7) * 16;
        *(int4*)dst = *(const int4*)dig[t];
    }
}

// tile prefix sums of the task list (one CTA; tasks[ntasks].tile0 = total).  Called after the tasks' sizes are final.
__global__ void __launch_bounds__(256) tc_tile_scan_kernel(TcTask* __restrict__ tasks, int ntasks) {
    __shared__ int s[256];
    const int tid = threadIdx.x, per = (ntasks + 255) / 256;
    const int b = min(ntasks, tid * per), e = min(ntasks, b + per);
    int acc = 0;
    for (int t = b; t < e; ++t) acc += ((tasks[t].M + 127) >> 7) * ((tasks[t].N + 63) >> 6);
    s[tid] = acc;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        const int x = (tid >= o) ? s[tid - o] : 0;
        __syncthreads();
        s[tid] += x;
        __syncthreads();
    }
    int run = s[tid] - acc;
    for (int t = b; t < e; ++t) {
        tasks[t].tile0 = run;
        run += ((tasks[t].M + 127) >> 7) * ((tasks[t].N + 63) >> 6);
    }
    if (tid == 255) tasks[ntasks].tile0 = s[255];
}

// ---------------------------------------------------------------------------------------------------------------------
// the GEMM
// ---------------------------------------------------------------------------------------------------------------------
#define TC_STAGES 2
template <int S> struct TcSmem {
    static constexpr int A_BYTES = S * 8192, B_BYTES = S * 4096, STAGE = A_BYTES + B_BYTES;
    static constexpr int TOTAL = TC_STAGES * STAGE + 1024;      // + barriers / tmem pointer (and alignment slack)
};

struct TcTile { int task, mt, nt, nkb; };
__device__ __forceinline__ TcTile tc_find_tile(const TcTask* __restrict__ tasks, int ntasks, int tile) {
    int lo = 0, hi = ntasks - 1;                                // last task with tile0 <= tile (empty tasks share their successor's tile0)
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (tasks[mid].tile0 <= tile) lo = mid; else hi = mid - 1; }
    TcTile r;
    r.task = lo;
    const int local = tile - tasks[lo].tile0, mtiles = (tasks[lo].M + 127) >> 7;
    r.mt = local % mtiles; r.nt = local / mtiles; r.nkb = (tasks[lo].K + 63) >> 6;
    return r;
}

template <int S>
__global__ void __launch_bounds__(192, 1) tc_gemm_kernel(const TcTask* __restrict__ tasks, int ntasks) {
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    using SM = TcSmem<S>;
    uint8_t* stage0 = tc_smem;
    uint64_t* bars = (uint64_t*)(tc_smem + TC_STAGES * SM::STAGE);
    uint64_t* full = bars;                    // [TC_STAGES] bulk copies landed
    uint64_t* empty = bars + TC_STAGES;       // [TC_STAGES] MMAs that read the stage are complete
    uint64_t* acc_full = bars + 2 * TC_STAGES;    // accumulators of the tile are complete
    uint64_t* acc_empty = acc_full + 1;           // the epilogue has read them
    uint32_t* tmem_ptr = (uint32_t*)(acc_empty + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total = tasks[ntasks].tile0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 128);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;

    if (warp == 0) {
        // ===== producer: one bulk copy per operand block (all S planes are contiguous) =====
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                const TcTile tl = tc_find_tile(tasks, ntasks, tile);
                GSI_BOUNDS(tl.task >= 0 && tl.task < ntasks && tl.mt * 128 < tasks[tl.task].M && tl.nt * 64 < tasks[tl.task].N);
                const int8_t* Ap = tasks[tl.task].Ap + (size_t)tl.mt * tl.nkb * SM::A_BYTES;
                const int8_t* Bp = tasks[tl.task].Bp + (size_t)tl.nt * tl.nkb * SM::B_BYTES;
                for (int kb = 0; kb < tl.nkb; ++kb, ++it) {
                    const int s = it % TC_STAGES, ph = (it / TC_STAGES) & 1;
                    mbar_wait(empty + s, ph ^ 1);
                    uint8_t* sa = stage0 + (size_t)s * SM::STAGE;
                    mbar_expect_tx(full + s, SM::STAGE);
                    bulk_g2s(sa, Ap + (size_t)kb * SM::A_BYTES, SM::A_BYTES, full + s);
                    bulk_g2s(sa + SM::A_BYTES, Bp + (size_t)kb * SM::B_BYTES, SM::B_BYTES, full + s);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int it = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tcount) {
                const int nkb = tc_find_tile(tasks, ntasks, tile).nkb;
                mbar_wait(acc_empty, (tcount & 1) ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % TC_STAGES, ph = (it / TC_STAGES) & 1;
                    mbar_wait(full + s, ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(stage0 + (size_t)s * SM::STAGE), sb = sa + SM::A_BYTES;
                    // For a fixed A slice t the partners u = 0 .. S-1-t write the accumulators t .. S-1: CONSECUTIVE column blocks of
                    // tensor memory, and the B planes of one k chunk are stored slice after slice -- so the S - t products are ONE
                    // MMA of N = 64 (S - t) columns (cut at the instruction's limit of 256): the A block is read from shared memory
                    // 12 times per K step instead of 36 (S = 8), which takes the operand reads off the shared-memory roofline.
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {             // two K steps of 32 per 64-byte k block
#pragma unroll
                        for (int t = 0; t < S; ++t) {
                            const uint64_t ad = umma_desc(sa + t * 8192 + kk * 4096, 2048, 128);
#pragma unroll
                            for (int c0 = 0; c0 < 64 * (S - t); c0 += 256) {
                                const int ncols = (64 * (S - t) - c0) < 256 ? (64 * (S - t) - c0) : 256;
                                const uint64_t bd = umma_desc(sb + kk * (2 * S * 1024) + (c0 >> 3) * 128, S * 1024, 128);
                                umma_i8(tmem + (uint32_t)(64 * t + c0), ad, bd, umma_idesc_i8(128, ncols), (kb > 0 || kk > 0 || t > 0) ? 1u : 0u);
                            }
                        }
                    }
                    umma_commit(empty + s);                      // frees the stage once these MMAs have read it
                }
                umma_commit(acc_full);
            }
        }
    } else {
        // ===== epilogue: warps 2..5, TMEM lanes 32 (warp % 4) .. + 31 = rows of the tile =====
        const int quad = warp & 3, row = quad * 32 + lane;
        int tcount = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tcount) {
            const TcTile tl = tc_find_tile(tasks, ntasks, tile);
            const TcTask T = tasks[tl.task];
            const int i = tl.mt * 128 + row;
            mbar_wait(acc_full, tcount & 1);
            tc_fence_after();
            // C_ij = 2^(e_i + f_j - 2P) 128^(2S-2) sum_g G_g 128^-g,  P = 7S:  2^(e_i + f_j - 14) per unit of the Horner sum
            const double rs = (i < T.M) ? ldexp(1.0, ((T.flags & TC_UNIT_A) ? 1 : tc_line_exp(T.ea[i])) - 7 - 7 * S) : 0.0;
            // C -= product: the 64 old values of this thread's row are requested before the accumulators are read, so that their
            // latency hides behind the tensor-memory loads instead of serialising load / store pairs column by column
            const bool sub = (T.flags & TC_SUB_C) != 0;
            double cold[64];
            if (sub) {
#pragma unroll
                for (int q = 0; q < 64; ++q) {
                    const int j = tl.nt * 64 + q;
                    cold[q] = (i < T.M && j < T.N) ? T.C[(size_t)(T.scatter ? T.scatter[j] : j) * T.ldc + i] : 0.0;
                }
            }
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 16) {
                double acc[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) acc[q] = 0.0;
#pragma unroll
                for (int g = S - 1; g >= 0; --g) {
                    uint32_t v[16];
                    tmem_ld16(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(64 * g + c0), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 16; ++q) acc[q] = fma(acc[q], 0.0078125, (double)(int)v[q]);
                }
                if (i < T.M) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const int j = tl.nt * 64 + c0 + q;
                        if (j < T.N) {
                            const int eb = (T.flags & TC_UNIT_B) ? 1 : tc_line_exp(T.eb[j]);
                            const double val = (tl.nkb > 0) ? acc[q] * rs * ldexp(1.0, eb - 7 + 7 * S) : 0.0;
                            const int jo = T.scatter ? T.scatter[j] : j;
                            T.C[(size_t)jo * T.ldc + i] = sub ? cold[c0 + q] - val : val;
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(acc_empty);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ---------------------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------------------
size_t tc_gemm_plane_bytes_a(int M, int K, int S) { return (size_t)((M + 127) / 128) * ((K + 63) / 64) * S * 8192; }
size_t tc_gemm_plane_bytes_b(int K, int N, int S) { return (size_t)((N + 63) / 64) * ((K + 63) / 64) * S * 4096; }

template <int S>
static cudaError_t tc_slice_run(const TcBatch& b, cudaStream_t st) {
    const int nkb = (b.Kmax + 63) / 64, mt = (b.Mmax + 127) / 128, nt = (b.Nmax + 63) / 64;
    tc_exp_init_kernel<<<dim3((b.Mmax + b.Nmax + 255) / 256, 1, b.ntasks), 256, 0, st>>>(b.tasks);
    tc_row_exp_kernel<<<dim3(nkb, mt, b.ntasks), 256, 0, st>>>(b.tasks);
    tc_col_exp_kernel<<<dim3(nkb, nt, b.ntasks), 256, 0, st>>>(b.tasks);
    tc_slice_a_kernel<S><<<dim3(nkb, mt, b.ntasks), 256, 0, st>>>(b.tasks);
    tc_slice_b_kernel<S><<<dim3(nkb, nt, b.ntasks), 256, 0, st>>>(b.tasks);
    return cudaGetLastError();
}

template <int S>
static cudaError_t tc_mma_run(const TcBatch& b, cudaStream_t st, int sms) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<S>::TOTAL);
    if (e != cudaSuccess) return e;
    tc_tile_scan_kernel<<<1, 256, 0, st>>>(b.tasks, b.ntasks);
    const int64_t bound = (int64_t)b.ntasks * ((b.Mmax + 127) / 128) * ((b.Nmax + 63) / 64);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(sms, bound));
    tc_gemm_kernel<S><<<grid, 192, TcSmem<S>::TOTAL, st>>>(b.tasks, b.ntasks);
    return cudaGetLastError();
}

cudaError_t tc_gemm_slice(const TcBatch& b, cudaStream_t st) {
    if (b.ntasks <= 0 || b.Mmax <= 0 || b.Nmax <= 0 || b.Kmax <= 0) return cudaSuccess;
    switch (b.S) {
        case 6: return tc_slice_run<6>(b, st);
        case 7: return tc_slice_run<7>(b, st);
        case 8: return tc_slice_run<8>(b, st);
        default: return cudaErrorInvalidValue;
    }
}
cudaError_t tc_gemm_mma(const TcBatch& b, cudaStream_t st, int sms) {
    if (b.ntasks <= 0 || b.Mmax <= 0 || b.Nmax <= 0 || b.Kmax <= 0) return cudaSuccess;
    switch (b.S) {
        case 6: return tc_mma_run<6>(b, st, sms);
        case 7: return tc_mma_run<7>(b, st, sms);
        case 8: return tc_mma_run<8>(b, st, sms);
        default: return cudaErrorInvalidValue;
    }
}
cudaError_t tc_gemm_batch(const TcBatch& b, cudaStream_t st, int sms) {
    cudaError_t e = tc_gemm_slice(b, st);
    return e != cudaSuccess ? e : tc_gemm_mma(b, st, sms);
}
