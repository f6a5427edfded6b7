// Internal declarations shared by the translation units of libgsi.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/gsi.h"

#define GSI_WARP 32
// In-kernel bounds checks for builds where compute-sanitizer is not available (it is closed on the pool this was developed on):
// `make debug` compiles libgsi_debug.so with -DGSI_DEBUG_BOUNDS, a failed check prints the site and traps (the launch then fails
// with an error the C ABI reports).  Off in the product build: the macro expands to nothing.
#ifdef GSI_DEBUG_BOUNDS
#include <stdio.h>
#define GSI_BOUNDS(cond)                                                                                          \
    do {                                                                                                          \
        if (!(cond)) {                                                                                            \
            printf("gsi bounds check failed: %s  (%s:%d, block %d,%d,%d thread %d)\n", #cond, __FILE__, __LINE__, \
                   (int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z, (int)threadIdx.x);                          \
            __trap();                                                                                             \
        }                                                                                                         \
    } while (0)
#else
#define GSI_BOUNDS(cond) ((void)0)
#endif
// One-sided Jacobi thresholds (fp64).  A pair is rotated when gamma^2 > ROT2 * alpha * beta; a
// sweep whose largest pre-rotation ratio^2 is <= STOP2 is the last one (quadratic convergence:
// a 3e-8 ratio before the sweep leaves ~1e-15 after it).
#define GSI_ROT2 1e-30
#define GSI_STOP2 9e-16
#define GSI_MAX_SWEEPS 40

// CTA-resident solver covers n <= 32 * GSI_S_MAX_EPT
#define GSI_S_MAX_EPT 5
#define GSI_S_MAX_N (32 * GSI_S_MAX_EPT)
// measured crossover (scripts/sweep_degree.py, profiles/r01b_degree_sweep.md): the Householder + D&C path is faster
// than the CTA-resident Jacobi kernel from n ~ 80 on (2.1x at n = 128)
#define GSI_SMALL_DEFAULT 44     // measured crossover (profiles/r01k_degree_sweep.md): above it Householder + D&C wins

// block Jacobi (large path): panels of M columns = two blocks of M/2; M is 64 (default) or 32
// largest n on the Householder path: its tridiagonalisation kernel keeps two np-long vectors in shared memory
#define GSI_HH_MAX_N 9216
#define GSI_BJ_ROWS 512  // rows of a panel handled by one CTA of the Gram / update kernels

struct gsi_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;   // record copies of a user group overlap the back-transform of the next one (host path)
    std::string err;
    int sm_count = 148;
    int64_t ws_limit = (int64_t)8 << 30;
    int bj_m = 64;
    int small_max = GSI_SMALL_DEFAULT;   // users with n <= small_max take the CTA-resident Jacobi kernel (GSI_SMALL_MAX, 32..160)
    bool trace = false;        // GSI_TRACE=1: per-launch timing lines on stderr (synchronising; diagnostics only)
    bool large_bj = false;     // GSI_LARGE=bj: block-Jacobi large path instead of Householder + D&C
    // weights
    double* d_w = nullptr;
    int w_rows = 0;
    bool own_w = false;
    // timing
    bool timing = false;
    double t_ms[GSI_T_COUNT] = {0};
    int64_t t_launch[GSI_T_COUNT] = {0};
    int64_t t_samples[GSI_T_COUNT] = {0};
    std::vector<cudaEvent_t> ev_pool;
    struct Span { int cls; cudaEvent_t a, b; };
    std::vector<Span> spans;
};

int gsi_fail(gsi_ctx* ctx, int code, const char* fmt, ...);

#define GSI_CUDA(ctx, call)                                                                   \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return gsi_fail((ctx), e_ == cudaErrorMemoryAllocation ? GSI_ERR_NOMEM : GSI_ERR_CUDA, \
                            "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)

// timing spans: begin/end around a group of launches of one class on ctx->stream
struct GsiSpan {
    gsi_ctx* ctx; int cls; cudaEvent_t a = nullptr, b = nullptr; int64_t launches;
    // counts `n_launches` launches; when timing is on, the span's time covers `n_timed` of them
    GsiSpan(gsi_ctx* c, int cls_, int64_t n_launches = 1, int64_t n_timed = -1);
    void end();
};
void gsi_count_launch(gsi_ctx* ctx, int cls, int64_t n);
