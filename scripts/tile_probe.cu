// What bounds the symv tile loop of trd_kernel?  One CTA (256 threads) per SM streams 64 x 64 fp64 tiles (32 KB) and does, per
// tile, a growing subset of what the kernel does.  Development aid (profiles/r01d_summary.md).
//   mode 0  bulk copy -> wait -> one 8-byte load per thread -> __syncthreads -> refill                  (copy engine alone)
//   mode 1  + the whole tile through LDS.128 (8 per thread)
//   mode 2  + x_J (4 broadcast LDS.128), x_I (1 LDS.128), 32 DFMA per thread (direct + transposed products)
//   mode 3  + direct partials to shared memory and the deferred 8-partial sum (9 LDS.64 + 8 DADD + 1 STS per thread)
//   mode 4  = mode 2, but only thread 0 polls the mbarrier (of the NEXT tile, before the barrier)
//   mode 5  no shared-memory staging: LDG.128 straight to registers, next tile prefetched while this one is used (+ mode 2 math)
//   mode 6  = mode 2 with 16-byte tile loads replaced by 8-byte ones (16 LDS.64 per thread)
//   mode 7  = mode 3 without __syncthreads: full/empty mbarriers per stage (a warp releases a stage as soon as the tile is in its
//             registers, thread 0 refills), partial sets handed over through mbarriers with 8 arrivals, sums deferred by one tile
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../collaborative_filtering_b200/csrc/ptx.cuh"

#define TILE_DBL 4096
#define STAGES 5

template <int MODE>
__global__ void __launch_bounds__(256, 1) tile_kernel(const double* __restrict__ src, int ntiles, double* out) {
    extern __shared__ __align__(128) unsigned char smraw[];
    double* stage = (double*)smraw;                            // [STAGES][4096]
    double* xs = stage + STAGES * TILE_DBL;                    // [1024] x vector
    double* dset = xs + 1024;                                  // [4][8][66]
    double* ys = dset + 4 * 8 * 66;                            // [1024]
    __shared__ uint64_t full[STAGES], empty[STAGES], dfull[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); } for (int s = 0; s < 4; ++s) mbar_init(&dfull[s], 8); mbar_fence_init(); }
    for (int i = tid; i < 1024; i += 256) { xs[i] = 1.0 + i; ys[i] = 0.0; }
    for (int i = tid; i < 4 * 8 * 66; i += 256) dset[i] = 0.0;
    __syncthreads();
    const double* base = src + (size_t)blockIdx.x * ntiles * TILE_DBL;
    double acc = 0.0;
    double tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (MODE == 7) {
        int issued = 0;
        if (tid == 0)
            for (; issued < min(STAGES, ntiles); ++issued) {
                mbar_expect_tx(&full[issued], TILE_DBL * 8);
                bulk_g2s(stage + (size_t)issued * TILE_DBL, base + (size_t)issued * TILE_DBL, TILE_DBL * 8, &full[issued]);
            }
        for (int k = 0; k < ntiles; ++k) {
            const int st = k % STAGES;
            mbar_wait(&full[st], (k / STAGES) & 1);
            const double* tile = stage + (size_t)st * TILE_DBL;
            const int I = k & 15, J = (k >> 4) & 15;
            const double2* tp = (const double2*)tile + (8 * warp) * 32 + lane;
            double2 a[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) a[q] = tp[q * 32];
            double xj[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) { const double2 v = ((const double2*)(xs + J * 64))[4 * warp + q]; xj[2 * q] = v.x; xj[2 * q + 1] = v.y; }
            const double2 xi = ((const double2*)(xs + I * 64))[lane];
            // the tile is in registers: release the stage (one arrival per warp)
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[st])) : "memory");
            // thread 0 refills the stage released one tile ago
            if (tid == 0 && k >= 1 && k - 1 + STAGES < ntiles) {
                const int sp = (k - 1) % STAGES;
                mbar_wait(&empty[sp], ((k - 1) / STAGES) & 1);
                mbar_expect_tx(&full[sp], TILE_DBL * 8);
                bulk_g2s(stage + (size_t)sp * TILE_DBL, base + (size_t)(k - 1 + STAGES) * TILE_DBL, TILE_DBL * 8, &full[sp]);
            }
            if (k >= 1) {                                      // deferred sum of tile k-1: every warp has delivered its partial
                mbar_wait(&dfull[(k - 1) & 3], ((k - 1) >> 2) & 1);
                const double* dp = dset + ((k - 1) & 3) * 528 + 8 * warp + (lane & 7);
                double* yp = ys + ((k + 15) & 15) * 64 + 8 * warp + (lane & 7);
                *yp = *yp + (((dp[0] + dp[66]) + (dp[132] + dp[198])) + ((dp[264] + dp[330]) + (dp[396] + dp[462])));
            }
            double2 de = make_double2(0, 0);
#pragma unroll
            for (int q = 0; q < 8; ++q) { de.x = fma(a[q].x, xj[q], de.x); de.y = fma(a[q].y, xj[q], de.y); }
#pragma unroll
            for (int q = 0; q < 8; ++q) tacc[q] = fma(a[q].x, xi.x, fma(a[q].y, xi.y, tacc[q]));
            *(double2*)(dset + (k & 3) * 528 + warp * 66 + 2 * lane) = de;
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&dfull[k & 3])) : "memory");
        }
    } else
    if (MODE == 5) {
        const double2* g = (const double2*)base + (8 * warp) * 32 + lane;
        double2 nxt[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) nxt[q] = __ldcg(g + q * 32);
        for (int k = 0; k < ntiles; ++k) {
            double2 a[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) a[q] = nxt[q];
            if (k + 1 < ntiles) {
                const double2* gn = g + (size_t)(k + 1) * (TILE_DBL / 2);
#pragma unroll
                for (int q = 0; q < 8; ++q) nxt[q] = __ldcg(gn + q * 32);
            }
            const int I = k & 15, J = (k >> 4) & 15;
            double xj[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) { const double2 v = ((const double2*)(xs + J * 64))[4 * warp + q]; xj[2 * q] = v.x; xj[2 * q + 1] = v.y; }
            const double2 xi = ((const double2*)(xs + I * 64))[lane];
            double2 de = make_double2(0, 0);
#pragma unroll
            for (int q = 0; q < 8; ++q) { de.x = fma(a[q].x, xj[q], de.x); de.y = fma(a[q].y, xj[q], de.y); }
#pragma unroll
            for (int q = 0; q < 8; ++q) tacc[q] = fma(a[q].x, xi.x, fma(a[q].y, xi.y, tacc[q]));
            acc += de.x + de.y;
        }
    } else {
        int issued = 0;
        if (tid == 0)
            for (; issued < min(STAGES, ntiles); ++issued) {
                mbar_expect_tx(&full[issued], TILE_DBL * 8);
                bulk_g2s(stage + (size_t)issued * TILE_DBL, base + (size_t)issued * TILE_DBL, TILE_DBL * 8, &full[issued]);
            }
        if (MODE == 4) { if (tid == 0) mbar_wait(&full[0], 0); __syncthreads(); }
        for (int k = 0; k < ntiles; ++k) {
            const int st = k % STAGES;
            if (MODE != 4) mbar_wait(&full[st], (k / STAGES) & 1);
            const double* tile = stage + (size_t)st * TILE_DBL;
            const int I = k & 15, J = (k >> 4) & 15;
            if (MODE == 0) acc += tile[tid];
            if (MODE == 1) {
                const double2* tp = (const double2*)tile + (8 * warp) * 32 + lane;
#pragma unroll
                for (int q = 0; q < 8; ++q) { const double2 v = tp[q * 32]; acc += v.x + v.y; }
            }
            if (MODE == 2 || MODE == 3 || MODE == 4) {
                const double2* tp = (const double2*)tile + (8 * warp) * 32 + lane;
                double2 a[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) a[q] = tp[q * 32];
                double xj[8];
#pragma unroll
                for (int q = 0; q < 4; ++q) { const double2 v = ((const double2*)(xs + J * 64))[4 * warp + q]; xj[2 * q] = v.x; xj[2 * q + 1] = v.y; }
                const double2 xi = ((const double2*)(xs + I * 64))[lane];
                if (MODE == 3) {
                    const double* dp = dset + ((k + 1) & 1) * 528 + 8 * warp + (lane & 7);
                    double* yp = ys + ((k + 15) & 15) * 64 + 8 * warp + (lane & 7);
                    *yp = *yp + (((dp[0] + dp[66]) + (dp[132] + dp[198])) + ((dp[264] + dp[330]) + (dp[396] + dp[462])));
                }
                double2 de = make_double2(0, 0);
#pragma unroll
                for (int q = 0; q < 8; ++q) { de.x = fma(a[q].x, xj[q], de.x); de.y = fma(a[q].y, xj[q], de.y); }
#pragma unroll
                for (int q = 0; q < 8; ++q) tacc[q] = fma(a[q].x, xi.x, fma(a[q].y, xi.y, tacc[q]));
                if (MODE == 3) *(double2*)(dset + (k & 1) * 528 + warp * 66 + 2 * lane) = de;
                else acc += de.x + de.y;
            }
            if (MODE == 6) {
                const double* tp = tile + (8 * warp) * 64 + lane;
                double a0[8], a1[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) { a0[q] = tp[q * 64]; a1[q] = tp[q * 64 + 32]; }
                double xj[8];
#pragma unroll
                for (int q = 0; q < 4; ++q) { const double2 v = ((const double2*)(xs + J * 64))[4 * warp + q]; xj[2 * q] = v.x; xj[2 * q + 1] = v.y; }
                const double xi0 = xs[I * 64 + lane], xi1 = xs[I * 64 + 32 + lane];
                double d0 = 0, d1 = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) { d0 = fma(a0[q], xj[q], d0); d1 = fma(a1[q], xj[q], d1); }
#pragma unroll
                for (int q = 0; q < 8; ++q) tacc[q] = fma(a0[q], xi0, fma(a1[q], xi1, tacc[q]));
                acc += d0 + d1;
            }
            if (MODE == 4 && tid == 0 && k + 1 < ntiles) mbar_wait(&full[(k + 1) % STAGES], ((k + 1) / STAGES) & 1);
            __syncthreads();
            if (tid == 0 && k + STAGES < ntiles) {
                mbar_expect_tx(&full[st], TILE_DBL * 8);
                bulk_g2s(stage + (size_t)st * TILE_DBL, base + (size_t)(k + STAGES) * TILE_DBL, TILE_DBL * 8, &full[st]);
            }
        }
    }
    for (int q = 0; q < 8; ++q) acc += tacc[q];
    out[blockIdx.x * 256 + tid] = acc + ys[tid];
}

template <int MODE>
static void run(const double* src, double* out, int ctas, int ntiles) {
    const size_t smem = (size_t)(STAGES * TILE_DBL + 1024 + 4 * 8 * 66 + 1024) * 8;
    cudaFuncSetAttribute(tile_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a);
        tile_kernel<MODE><<<ctas, 256, smem>>>(src, ntiles, out);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); best = fminf(best, ms);
    }
    const double per = (double)ntiles * TILE_DBL * 8;
    printf("mode %d ctas=%3d  %7.1f GB/s per SM  %8.1f GB/s total  %.3f us per tile  (%s)\n", MODE, ctas, per / best / 1e6, ctas * per / best / 1e6,
           best * 1e3 / ntiles, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int ntiles = 8192;                                   // 256 MB per CTA
    const size_t total = (size_t)148 * ntiles * TILE_DBL * 8;
    double* src; double* out;
    cudaMalloc(&src, total); cudaMalloc(&out, 148 * 256 * 8);
    cudaMemset(src, 0, total);
    const int ctas_list[] = {1, 8, 148};
    for (int ctas : ctas_list) {
        run<0>(src, out, ctas, ntiles); run<1>(src, out, ctas, ntiles); run<2>(src, out, ctas, ntiles); run<3>(src, out, ctas, ntiles);
        run<4>(src, out, ctas, ntiles); run<5>(src, out, ctas, ntiles); run<6>(src, out, ctas, ntiles); run<7>(src, out, ctas, ntiles);
    }
    return 0;
}
