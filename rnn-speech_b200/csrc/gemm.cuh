// Internal GEMM interface used by the acoustic-model kernels (row-major fp32).
#pragma once
#include "common.cuh"

namespace rs {

// C[M,N] = (accumulate ? C : 0) + op(A) * op(B) (+ bias[N] broadcast over rows)
//   transA == 0: A is [M,K] (lda >= K)      transA == 1: A is [K,M] (lda >= M)
//   transB == 0: B is [K,N] (ldb >= N)      transB == 1: B is [N,K] (ldb >= K)
// Plain fp32 FFMA tiles (128x128x16, 8x8 per thread): the generic path for any
// shape; the tcgen05 path in gemm_tc.cu takes over for the large aligned shapes.
int sgemm(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
          float* C, int ldc, const float* bias, int accumulate, cudaStream_t st);

// out[n] (+)= sum_m A[m, n]   (column sums of a row-major [M,N] matrix)
int colsum(const float* A, int M, int N, int lda, float* out, int accumulate, cudaStream_t st);

}  // namespace rs
