// Interface of tc_gemm.cu: FP64-equivalent GEMM on tcgen05 (INT8 slices, INT32 accumulators in tensor memory).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct TcGemmParams {            // device-side view
    const int8_t* Ap; const int8_t* Bp;     // slice planes (tc_gemm.cu: "slicing")
    const int32_t* ea; const int32_t* eb;   // row exponents of A, column exponents of B
    double* C; int64_t ldc;                 // column-major M x N
    const int32_t* scatter;                 // optional: column j of the product goes to column scatter[j] of C
    int M, N, K;
};

struct TcGemmArgs {              // C[:, scatter[j]] = A[:, gather[0..K)] * B[0..K, j], all column-major FP64
    const double* A; int64_t lda;           // M x (>= max gather) ; element (i, k) = A[gather[k] * lda + i]
    const double* B; int64_t ldb;           // K x N
    double* C; int64_t ldc;
    const int32_t* gather;                  // optional [K]
    const int32_t* scatter;                 // optional [N]
    int M, N, K, S;                         // S = slices per operand (6, 7 or 8: 42 / 49 / 56 bit fixed point)
    int8_t* Ap; int8_t* Bp;                 // workspaces: tc_gemm_plane_bytes_a / _b
    int32_t* ea; int32_t* eb;               // workspaces: M and N ints
};

size_t tc_gemm_plane_bytes_a(int M, int K, int S);
size_t tc_gemm_plane_bytes_b(int K, int N, int S);
// enqueues exponent scan, slicing and the GEMM on `st`; sms = CTAs of the persistent GEMM kernel
cudaError_t tc_gemm_fp64(const TcGemmArgs& a, cudaStream_t st, int sms);
cudaError_t tc_gemm_slice_only(const TcGemmArgs& a, cudaStream_t st);            // the two halves, for separate timing
cudaError_t tc_gemm_mma_only(const TcGemmArgs& a, cudaStream_t st, int sms);
