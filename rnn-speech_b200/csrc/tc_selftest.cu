// Compiled into librnnspeech_b200_diag.so only (-DRS_DIAG): self-tests and micro-benchmarks of the tcgen05 plumbing.
#ifdef RS_DIAG
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

// ---------------------------------------------------------------------------------------
// Issue-rate microbenchmark: one CTA issues `count` tcgen05.mma (kind::f16, SS) of shape
// M x N x 16 over `ntiles` distinct resident A/B K-slices and `nacc` accumulators, and
// reports the elapsed SM cycles from first issue to completion (mbarrier commit).
// Used to size the recurrent kernels (DESIGN.md "Recurrent step budget").
// ---------------------------------------------------------------------------------------
namespace rs {
namespace {
template <int M, int N, int NACC, bool WARP>
__device__ __forceinline__ void mma_bench_body(uint32_t tmem, uint32_t sa, uint32_t sb, int count, uint64_t* bar,
                                               long long* out_cycles) {
  // 8 resident K-slices (2 K-blocks x 4), descriptors precomputed; `count` multiple of 8
  const uint32_t idesc = tc::instr_desc_bf16(M, N);
  uint64_t da[8], db[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    da[i] = tc::smem_desc_sw128(sa + (i >> 2) * 128 * 128) + 2 * (i & 3);
    db[i] = tc::smem_desc_sw128(sb + (i >> 2) * N * 128) + 2 * (i & 3);
  }
  const long long t0 = clock64();
  for (int it = 0; it < count; it += 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t d = tmem + (uint32_t)((i % NACC) * N);
      if (WARP) tc::mma_bf16_ss_warp(d, da[i], db[i], idesc, (uint32_t)(it > 0 || i >= NACC));
      else tc::mma_bf16_ss(d, da[i], db[i], idesc, it > 0 || i >= NACC);
    }
  }
  const long long t1 = clock64();
  if (WARP) tc::mma_commit_warp(bar); else tc::mma_commit(bar);
  tc::mbar_wait(bar, 0);
  const long long t2 = clock64();
  if ((threadIdx.x & 31) == 0) { out_cycles[0] = t1 - t0; out_cycles[1] = t2 - t0; }
}

__global__ void __launch_bounds__(128, 1)
tc_mma_bench_kernel(int M, int N, int count, int variant, int nacc, long long* out_cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (2 * 128 * 128 + 2 * 256 * 128) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 256);
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t sa = tc::smem_u32(smem), sb = sa + 2 * 128 * 128;
  const bool wmode = variant == 1;
  if (warp == 1 && (wmode || (tid & 31) == 0)) {
#define RS_CASE(MM, NN, AA) \
    if (M == MM && N == NN && nacc == AA) { \
      if (wmode) mma_bench_body<MM, NN, AA, true>(tmem, sa, sb, count, &bar, out_cycles); \
      else mma_bench_body<MM, NN, AA, false>(tmem, sa, sb, count, &bar, out_cycles); }
    RS_CASE(64, 32, 1) RS_CASE(64, 32, 4) RS_CASE(128, 32, 1) RS_CASE(128, 32, 4)
    RS_CASE(64, 16, 1) RS_CASE(64, 16, 4) RS_CASE(128, 128, 1) RS_CASE(128, 256, 1) RS_CASE(64, 64, 1) RS_CASE(64, 64, 4)
#undef RS_CASE
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}
}  // namespace
}  // namespace rs

// out_cycles_d: int64[2] = {issue cycles, issue-to-completion cycles}
// variant 0: issued by one thread inside a divergent region; 1: converged warp + elect.sync
extern "C" int rs_tc_mma_bench(int M, int N, int count, int variant, int nacc, void* out_cycles_d, void* stream) {
  RS_REQUIRE((M == 64 || M == 128) && N >= 16 && N <= 256 && N % 16 == 0 && count > 0 && count % 8 == 0 && nacc > 0 &&
                 N * nacc <= 256, RS_ERR_INVALID, "rs_tc_mma_bench: bad arguments");
  size_t smem = 2 * 128 * 128 + 2 * 256 * 128 + 1024;
  RS_CHECK_CUDA(cudaFuncSetAttribute(tc_mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_mma_bench_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(M, N, count, variant, nacc, (long long*)out_cycles_d);
  RS_CHECK_LAUNCH();
  return RS_OK;
}

// ---------------------------------------------------------------------------------------
// "TS" self-test: the A operand lives in TMEM (written with tcgen05.st), stacked the way the
// TMEM-resident recurrent kernels stack it: sub-partition q holds rows 16q..16q+15 of A_hi in
// lanes 32q..32q+15 and the same rows of A_lo in lanes 32q+16..32q+31.  B is the stacked
// [B_hi (32 rows); B_lo (32 rows)] N = 64 tile in shared memory.  The raw accumulator
// D[128 lanes][64] comes back so the test can pin every quadrant (hi*hi, hi*lo, lo*hi, lo*lo).
// Also reports the cycles of `reps` back-to-back passes over K (issue -> completion).
// ---------------------------------------------------------------------------------------
namespace rs {
namespace {
__global__ void __launch_bounds__(128, 1)
tc_ts_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int K, int variant,
                      int reps, long long* out_cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkb = K / 64;
  constexpr uint32_t B_KB_BYTES = 64 * 128;
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  for (int e = tid; e < 32 * K; e += 128) {
    const int r = e / K, k = e % K;
    __nv_bfloat16 hi, lo;
    tc::split_bf16(B[e], hi, lo);
    unsigned char* tile = smem + (size_t)(k / 64) * B_KB_BYTES;
    *reinterpret_cast<__nv_bfloat16*>(tile + tc::swz_off(r, k % 64)) = hi;
    *reinterpret_cast<__nv_bfloat16*>(tile + tc::swz_off(32 + r, k % 64)) = lo;
  }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t tmemD = tmem + ((variant >> 8) ? (uint32_t)(variant >> 8) : 448u);     // variant >> 8: accumulator column
  {
    // this thread's TMEM lane = tid: row 16*warp + (lane & 15), hi plane for lane < 16, lo otherwise
    const int row = 16 * warp + (lane & 15);
    const bool is_lo = lane >= 16;
    for (int c0 = 0; c0 < K / 2; c0 += 8) {
      uint32_t r[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        __nv_bfloat16 h0, l0, h1, l1;
        tc::split_bf16(A[(size_t)row * K + 2 * (c0 + i)], h0, l0);
        tc::split_bf16(A[(size_t)row * K + 2 * (c0 + i) + 1], h1, l1);
        const __nv_bfloat16 e0 = is_lo ? l0 : h0, e1 = is_lo ? l1 : h1;
        r[i] = (variant & 1) ? tc::pack_bf16(e1, e0) : tc::pack_bf16(e0, e1);
      }
      tc::tmem_st8(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0, r);
    }
    tc::tmem_st_wait();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (warp == 1) {
    const uint32_t idesc = tc::instr_desc_bf16(128, 64);
    const uint64_t db0 = tc::smem_desc_sw128(tc::smem_u32(smem));
    long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep)
      for (int kb = 0; kb < nkb; ++kb)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc::mma_bf16_ts_warp(tmemD, tmem + (uint32_t)(kb * 32 + k * 8), db0 + (uint64_t)kb * (B_KB_BYTES >> 4) + 2 * k, idesc,
                               (uint32_t)((kb | k) != 0));
    long long t1 = clock64();
    tc::mma_commit_warp(&bar);
    tc::mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (lane == 0 && out_cycles) { out_cycles[0] = t1 - t0; out_cycles[1] = t2 - t0; }
  }
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  for (int c = 0; c < 4; ++c) {
    float v[16];
    tc::tmem_ld16(tmemD + ((uint32_t)(32 * warp) << 16) + (uint32_t)(c * 16), v);
    tc::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) D[(size_t)tid * 64 + c * 16 + i] = v[i];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}
}  // namespace
}  // namespace rs

// A_d [64,K], B_d [32,K] fp32; D_d [128,64] fp32 raw accumulator (TMEM lane, column); K multiple of 64, <= 768.
// variant bit0: swap the two bf16 halves of each TMEM column.  out_cycles_d: int64[2] or NULL.
extern "C" int rs_tc_ts_selftest(const float* A_d, const float* B_d, float* D_d, int K, int variant, int reps,
                                 void* out_cycles_d, void* stream) {
  RS_REQUIRE(A_d && B_d && D_d, RS_ERR_INVALID, "rs_tc_ts_selftest: NULL argument");
  RS_REQUIRE(K >= 64 && K % 64 == 0 && K <= 768 && reps >= 1, RS_ERR_INVALID, "rs_tc_ts_selftest: K=%d unsupported", K);
  size_t smem = (size_t)(K / 64) * 64 * 128 + 1024;
  RS_CHECK_CUDA(cudaFuncSetAttribute(tc_ts_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_ts_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A_d, B_d, D_d, K, variant, reps, (long long*)out_cycles_d);
  RS_CHECK_LAUNCH();
  return RS_OK;
}

#endif  // RS_DIAG
