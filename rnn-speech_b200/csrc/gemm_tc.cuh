// tcgen05 GEMMs for the batched projections of the acoustic model.
#pragma once
#include "common.cuh"
#include <cuda_bf16.h>

namespace rs {

// A "split" activation / weight matrix: x ~= hi + lo, two bf16 planes with one layout.
struct SplitMat {
  const __nv_bfloat16* hi;
  const __nv_bfloat16* lo;   // may be nullptr (plain bf16)
  int rows, cols, ld;        // row-major, ld in elements (multiple of 8)
};

enum GemmOut {
  GEMM_OUT_F32 = 0,        // C[m*ldc + n] fp32 (+bias) (+= when accumulate)
  GEMM_OUT_REC = 1,        // gate pre-activations in the recurrent kernel's layout (see lstm_rec_tc.cu)
  GEMM_OUT_SPLIT = 2       // bf16 hi/lo planes, row-major ldc
};

struct GemmTcOut {
  int mode;
  float* C;
  __nv_bfloat16* Chi;
  __nv_bfloat16* Clo;
  int ldc;
  const float* bias;       // [N] or nullptr
  int accumulate;          // GEMM_OUT_F32 only
  // GEMM_OUT_REC (transposed product): row m = permuted gate row (unit/U)*4U + (unit%U)*4 + g (the A operand
  //   is the pack_wrec() layout of the weights), col n = t*B + b -> C[(t*4H + m) * Bpad + b];
  //   bias is indexed in the original order g*H + unit
  int recB, recBpad, recH, recU;
  int max_ctas;            // 0 = one CTA per SM; otherwise cap the persistent grid (side-stream GEMMs)
  // Elastic grid: with `elastic` set the launch has one CTA per SM, but only the first max_ctas of them work unless
  // elastic[0] != 0 when the first CTA looks (the pipelined schedule raises it once its recurrent launches are done,
  // so that the GEMMs still queued behind them take the whole machine).  elastic[1 + elastic_id] records the
  // decision for the launch (zeroed by the caller before the step).
  int* elastic;
  int elastic_id;
  int coresident;          // 1: the 129 KB / 256-TMEM-column variant that shares SMs with the backward recurrent CTAs
  int tiles_per_cta;       // > 0: grid = ceil(tiles / tiles_per_cta) short-lived CTAs instead of a persistent grid, so
                           // that the block scheduler can place them wherever (and whenever) SMs are free
  // GEMM_OUT_SPLIT only: dropout of the result before the split (common.cuh dropout_keep; element index m*N + n).
  // drop_thr = 0 (a zero-initialised struct) or 0xffffffff: no dropout
  uint64_t drop_key;
  uint32_t drop_stream, drop_thr;
  float drop_inv;
};

// C[M,N] = A[M,K] * B[N,K]^T, both operands K-major (row-major with K contiguous).
// products = 3: bf16x3 over whichever lo planes are present (A_hi*B_hi + A_hi*B_lo + A_lo*B_hi),
// 1: plain bf16 (hi planes only).
// Shape rules: lda/ldb multiples of 8 elements, 16-byte aligned bases; any M, N, K
// (tails are zero-filled by TMA and masked in the epilogue).
int gemm_tc_nt(const SplitMat& A, const SplitMat& B, int M, int N, int K, int products, const GemmTcOut& out,
               cudaStream_t st);

// C[M,N] = A^T B with both operands stored K-outermost ("MN-major"): A is [K][M], B is [K][N], row-major, ld in
// elements (SplitMat.rows = K).  The weight-gradient products x^T dgates read x, h and dgates exactly as the
// forward / backward kernels wrote them ((t, b) as the row index): no transposed copies.  fp32 output only.
int gemm_tc_tn(const SplitMat& A, const SplitMat& B, int M, int N, int K, int products, const GemmTcOut& out,
               cudaStream_t st);
// out[c] (+)= sum_r (hi[r][c] + lo[r][c]) over a [R, C] plane pair (lo may be nullptr): bias gradients from dgates.
// Bit-reproducible.  scratch: colsum_scratch_bytes(C) bytes whose leading 1024 counter words are zero before the first
// call (the kernel leaves them zero); calls that may run concurrently need separate scratch, consecutive calls (any C up
// to the one the scratch was sized for) may share it.
size_t colsum_scratch_bytes(int C);
int colsum_planes(const __nv_bfloat16* hi, const __nv_bfloat16* lo, int R, int C, int ld, float* out, int accumulate,
                  void* scratch, cudaStream_t st);

// out planes <- split(in * scale) elementwise; optional dropout masks as in lstm.cu
int split_planes(const float* in, __nv_bfloat16* hi, __nv_bfloat16* lo, int64_t n, cudaStream_t st);
// transpose + split: in [R,C] row-major (ld_in) -> planes [C,R] row-major (ld_out)
int split_planes_transposed(const float* in, int R, int C, int ld_in, __nv_bfloat16* hi, __nv_bfloat16* lo, int ld_out,
                            cudaStream_t st);
// bf16 matrix transpose: in [R,C] (ld_in) -> out [C,R] (ld_out)
int transpose_bf16(const __nv_bfloat16* in, int R, int C, int ld_in, __nv_bfloat16* out, int ld_out, cudaStream_t st);
// out[r] (+)= sum_c (hi[r][c] + lo[r][c])  over a [R, C] plane pair (lo may be nullptr)
int rowsum_planes(const __nv_bfloat16* hi, const __nv_bfloat16* lo, int R, int C, int ld, float* out, int accumulate,
                  cudaStream_t st);

}  // namespace rs
