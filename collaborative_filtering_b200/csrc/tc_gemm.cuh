// Interface of tc_gemm.cu: FP64-equivalent GEMM on tcgen05 (INT8 slices, INT32 accumulators in tensor memory).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// One product C[:, scatter[j]] = A[:, gather[0..K)] * B[0..K, j], all column-major FP64.  The list lives in DEVICE memory: M, N, K
// (and the pointers) may be written by a kernel earlier in the stream -- the divide-and-conquer fills them from its deflation
// counts.  The array has ntasks + 1 entries; tile0 is filled by the engine (prefix sum of the tiles, total in the last entry).
struct TcTask {
    const double* A; const double* B; double* C;
    const int32_t* gather;                  // optional [K]: column of A that is the k-th column of the product's left operand
    const int32_t* scatter;                 // optional [N]: column of C that receives column j of the product
    int64_t lda, ldb, ldc;
    int8_t* Ap; int8_t* Bp;                 // slice planes, sized by tc_gemm_plane_bytes_a / _b of the task's UPPER bounds
    int32_t* ea; int32_t* eb;               // row exponents of A [M], column exponents of B [N]
    int M, N, K, tile0;
    int flags, pad_;                        // TC_TRANS_A: element (i, k) of the left operand is A[i * lda + k] (gather unused); TC_SUB_C: C -= product
};
#define TC_TRANS_A 1
#define TC_SUB_C 2
#define TC_UNIT_A 4                         // every entry of the left / right operand is at most 1 in magnitude (orthonormal columns,
#define TC_UNIT_B 8                         // Householder vectors with unit head): fixed exponent 1, no exponent scan

struct TcBatch {
    TcTask* tasks; int ntasks;              // device pointer
    int Mmax, Nmax, Kmax;                   // host-side upper bounds over the tasks (grid sizes)
    int S;                                  // slices per operand entry: 6, 7 or 8 digits of 7 bits
};

size_t tc_gemm_plane_bytes_a(int M, int K, int S);
size_t tc_gemm_plane_bytes_b(int K, int N, int S);
cudaError_t tc_gemm_batch(const TcBatch& b, cudaStream_t st, int sms);   // exponent scan + slicing + GEMM, enqueued on `st`
cudaError_t tc_gemm_slice(const TcBatch& b, cudaStream_t st);            // the two halves, for separate timing
cudaError_t tc_gemm_mma(const TcBatch& b, cudaStream_t st, int sms);     // sms = CTAs of the persistent kernel
