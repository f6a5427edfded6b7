// Per-SM streaming bandwidth probe: bulk-copy (cp.async.bulk + mbarrier) vs plain LDG.128, for different
// request sizes / stage counts / number of participating CTAs.  Development aid.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../collaborative_filtering_b200/csrc/ptx.cuh"

__global__ void __launch_bounds__(256, 1) bulk_stream(const double* __restrict__ src, size_t per_cta_bytes, int req_bytes, int stages, double* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ uint64_t full[16];
    const int tid = threadIdx.x;
    if (tid == 0) { for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1); mbar_fence_init(); }
    __syncthreads();
    const char* base = (const char*)src + (size_t)blockIdx.x * per_cta_bytes;
    const int nreq = (int)(per_cta_bytes / req_bytes);
    int issued = 0;
    if (tid == 0)
        for (; issued < min(stages, nreq); ++issued) {
            mbar_expect_tx(&full[issued], req_bytes);
            bulk_g2s(sm + (size_t)issued * req_bytes, base + (size_t)issued * req_bytes, req_bytes, &full[issued]);
        }
    double acc = 0;
    for (int k = 0; k < nreq; ++k) {
        const int st = k % stages;
        mbar_wait(&full[st], (k / stages) & 1);
        const double* p = (const double*)(sm + (size_t)st * req_bytes);
        acc += p[tid];
        __syncthreads();
        if (tid == 0 && k + stages < nreq) {
            mbar_expect_tx(&full[st], req_bytes);
            bulk_g2s(sm + (size_t)st * req_bytes, base + (size_t)(k + stages) * req_bytes, req_bytes, &full[st]);
        }
    }
    out[blockIdx.x * 256 + tid] = acc;
}

__global__ void __launch_bounds__(256, 1) ldg_stream(const double2* __restrict__ src, size_t per_cta_bytes, double* out) {
    const double2* base = src + (size_t)blockIdx.x * per_cta_bytes / 16;
    const size_t n = per_cta_bytes / 16;
    double acc = 0;
    for (size_t i = threadIdx.x; i < n; i += 256 * 8) {
        double2 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (i + u * 256 < n) ? __ldcg(base + i + u * 256) : make_double2(0, 0);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].y;
    }
    out[blockIdx.x * 256 + threadIdx.x] = acc;
}

int main() {
    const size_t total = (size_t)4 << 30;
    double* src; double* out;
    cudaMalloc(&src, total); cudaMalloc(&out, 148 * 256 * 8);
    cudaMemset(src, 0, total);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaFuncSetAttribute(bulk_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int ctas_list[] = {1, 8, 37, 74, 148};
    for (int ci = 0; ci < 5; ++ci) {
        const int ctas = ctas_list[ci];
        const size_t per = ((total / 148) / (256 * 1024)) * (256 * 1024);
        struct { int req, stages; } cfg[] = {{32768, 4}, {32768, 6}, {16384, 8}, {16384, 12}, {8192, 16}, {65536, 3}, {4096, 16}};
        for (auto c : cfg) {
            float best = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(a);
                bulk_stream<<<ctas, 256, (size_t)c.req * c.stages>>>(src, per, c.req, c.stages, out);
                cudaEventRecord(b); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b); best = fminf(best, ms);
            }
            printf("bulk ctas=%3d req=%6d stages=%2d  %8.1f GB/s total  %6.1f GB/s per SM  (%s)\n", ctas, c.req, c.stages,
                   ctas * (double)per / best / 1e6, (double)per / best / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(a);
            ldg_stream<<<ctas, 256>>>((const double2*)src, per, out);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); best = fminf(best, ms);
        }
        printf("ldg  ctas=%3d                        %8.1f GB/s total  %6.1f GB/s per SM\n", ctas, ctas * (double)per / best / 1e6, (double)per / best / 1e6);
    }
    return 0;
}
