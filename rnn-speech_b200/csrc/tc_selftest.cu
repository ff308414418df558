// Single-CTA tcgen05 self-test: D[128,N] = A[128,K] * B[N,K]^T through the same
// descriptor / swizzle / TMEM helpers the production kernels use (tc_common.cuh).
// Exists so that tests/test_gpu_tc.py can validate the encodings in isolation.
#include "common.cuh"
#include "tc_common.cuh"

namespace rs {
namespace {

// dynamic smem (1024-aligned): A_hi | A_lo | B_hi | B_lo ; K-blocks of 64 elements
__global__ void __launch_bounds__(128, 1)
tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int N, int K,
                   int split) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkb = K / 64;
  const uint32_t a_kb_bytes = 128 * 128, b_kb_bytes = (uint32_t)N * 128;
  unsigned char* a_hi = smem;
  unsigned char* a_lo = a_hi + nkb * a_kb_bytes;
  unsigned char* b_hi = a_lo + nkb * a_kb_bytes;
  unsigned char* b_lo = b_hi + nkb * b_kb_bytes;
  uint32_t ncols = 32;
  while ((int)ncols < N) ncols <<= 1;

  if (warp == 0) tc::tmem_alloc(&tmem_slot, ncols);
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_mbar_init();
  }
  for (int e = tid; e < 128 * K; e += 128) {
    const int r = e / K, k = e % K;
    __nv_bfloat16 hi, lo;
    tc::split_bf16(A[e], hi, lo);
    const uint32_t off = (k / 64) * a_kb_bytes + tc::swz_off(r, k % 64);
    *reinterpret_cast<__nv_bfloat16*>(a_hi + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(a_lo + off) = lo;
  }
  for (int e = tid; e < N * K; e += 128) {
    const int r = e / K, k = e % K;
    __nv_bfloat16 hi, lo;
    tc::split_bf16(B[e], hi, lo);
    const uint32_t off = (k / 64) * b_kb_bytes + tc::swz_off(r, k % 64);
    *reinterpret_cast<__nv_bfloat16*>(b_hi + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(b_lo + off) = lo;
  }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (tid == 0) {
    const uint32_t idesc = tc::instr_desc_bf16(128, N);
    bool acc = false;
    for (int kb = 0; kb < nkb; ++kb) {
      const uint64_t dah = tc::smem_desc_sw128(tc::smem_u32(a_hi + kb * a_kb_bytes));
      const uint64_t dal = tc::smem_desc_sw128(tc::smem_u32(a_lo + kb * a_kb_bytes));
      const uint64_t dbh = tc::smem_desc_sw128(tc::smem_u32(b_hi + kb * b_kb_bytes));
      const uint64_t dbl = tc::smem_desc_sw128(tc::smem_u32(b_lo + kb * b_kb_bytes));
      for (int k = 0; k < 4; ++k) {        // UMMA_K = 16 bf16 = 32 bytes = 2 descriptor units
        tc::mma_bf16_ss(tmem, dah + 2 * k, dbh + 2 * k, idesc, acc);
        acc = true;
        if (split) {
          tc::mma_bf16_ss(tmem, dah + 2 * k, dbl + 2 * k, idesc, true);
          tc::mma_bf16_ss(tmem, dal + 2 * k, dbh + 2 * k, idesc, true);
        }
      }
    }
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  for (int c = 0; c < N / 16; ++c) {
    float v[16];
    tc::tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)(c * 16), v);
    tc::tmem_ld_wait();
    const int row = 32 * warp + lane;
#pragma unroll
    for (int i = 0; i < 16; ++i) D[(size_t)row * N + c * 16 + i] = v[i];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, ncols);
}

}  // namespace
}  // namespace rs

using namespace rs;

// A_d [128,K] fp32, B_d [N,K] fp32, D_d [128,N] fp32; N multiple of 16 in [16,256], K multiple of 64.
extern "C" int rs_tc_selftest(const float* A_d, const float* B_d, float* D_d, int N, int K, int split, void* stream) {
  RS_REQUIRE(A_d && B_d && D_d, RS_ERR_INVALID, "rs_tc_selftest: NULL argument");
  RS_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && K >= 64 && K % 64 == 0, RS_ERR_INVALID,
             "rs_tc_selftest: N=%d K=%d unsupported", N, K);
  size_t smem = (size_t)(K / 64) * (128 * 128 + (size_t)N * 128) * 2 + 1024;
  RS_REQUIRE(smem <= 220 * 1024, RS_ERR_UNSUPPORTED, "rs_tc_selftest: needs %zu B smem", smem);
  RS_CHECK_CUDA(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A_d, B_d, D_d, N, K, split);
  RS_CHECK_LAUNCH();
  return RS_OK;
}
