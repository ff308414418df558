// sm_100a tensor-core plumbing: tcgen05 MMA with TMEM accumulators, shared-memory
// matrix descriptors, mbarriers, TMA bulk tensor loads.  Inline PTX only.
//
// Conventions used by every kernel in this library
//   * operands are bf16, K-major, in the canonical SWIZZLE_128B layout: a tile of R rows
//     by 64 K-elements is R x 128 bytes, 8-row groups of 1024 bytes, 16-byte chunk index
//     XOR (row & 7).  Tiles start on 1024-byte boundaries.  TMA writes this layout
//     directly (CU_TENSOR_MAP_SWIZZLE_128B); st_swz() writes it from registers.
//   * D[M=128, N] fp32 lives in TMEM: lane = row, column = n.
//   * fp32-grade products from bf16 tensor cores ("bf16x3"): x = hi + lo with
//     hi = bf16(x), lo = bf16(x - hi);  a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi
//     (relative error ~2^-16 per product, fp32 accumulate).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace rs {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- bf16 split -------------------------------------------------------------------
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// byte offset of element (row r, k-element k in [0,64)) inside a SWIZZLE_128B K-major tile
__device__ __forceinline__ uint32_t swz_off(int r, int k) {
  const int chunk = (k >> 3) ^ (r & 7);
  return (uint32_t)(r * 128 + chunk * 16 + (k & 7) * 2);
}

// ---- mbarrier ---------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}

// ---- proxies / fences ---------------------------------------------------------------
// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma / TMA reads of smem)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// generic-proxy global writes <-> async-proxy (TMA) accesses of global memory
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM ---------------------------------------------------------------------------
// one full warp; ncols power of two >= 32; the base address is written to *slot (smem)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// lane l of the warp receives TMEM lane (lane field of taddr + l), 16 consecutive columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// lane l of the warp writes 8 consecutive 32-bit columns of TMEM lane (lane field of taddr + l)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- descriptors --------------------------------------------------------------------
// K-major SWIZZLE_128B operand tile (rows x 64 bf16, 128 B per row, 1024 B per 8-row group)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);   // start address, 16-byte units
  d |= (uint64_t)1 << 16;                              // leading byte offset (unused with swizzle)
  d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset: 8-row group pitch
  d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
  return d;
}
// MN-major SWIZZLE_128B operand tile: the operand is stored with its M (or N) index contiguous -- 64 of them per
// 128-byte row, one row per K index, 8-row groups of 1024 bytes (what a TMA box {64 mn, 64 k} of a [K][MN] row-major
// matrix lands as).  Canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units: LBO = distance between
// 64-element MN blocks, SBO = distance between 8-row K groups.
__device__ __forceinline__ uint64_t smem_desc_sw128_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16, A = B = bf16 (K-major), D = fp32, shape M x N x 16
__host__ __device__ constexpr uint32_t instr_desc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// the same with both operands MN-major (bit 15: A, bit 16: B)
__host__ __device__ constexpr uint32_t instr_desc_bf16_mn(int M, int N) {
  return instr_desc_bf16(M, N) | (1u << 15) | (1u << 16);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; one thread issues
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// Same, for a fully converged warp: every lane executes this, one elected lane issues.
// (Issuing from inside an `if (lane == 0)` region makes the compiler wrap every UTCHMMA in an
// ELECT/BRA.U.ANY retry loop, which costs more than the MMA itself for small tiles.)
__device__ __forceinline__ void mma_bf16_ss_warp(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand resident in TMEM ("TS" form): A[M=128, K=16] bf16 occupies lanes 0..127 x 8 columns
// (column c of a row holds K elements 2c | 2c+1 << 16); a_tmem = address of its first column.
__device__ __forceinline__ void mma_bf16_ts_warp(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Four consecutive K = 16 steps of one 64-element K-block in ONE asm block (one elect, no
// per-MMA compiler glue): A advances 8 TMEM columns / 32 bytes per step, B 32 bytes.
__device__ __forceinline__ void mma4_bf16_ts_warp(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                                  uint32_t accumulate_first) {
  asm volatile(
      "{\n\t.reg .pred p, q, t;\n\t"
      ".reg .b32 a1, a2, a3;\n\t"
      ".reg .b64 b1, b2, b3;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.eq.b32 t, %3, %3;\n\t"
      "add.u32 a1, %1, 8;\n\t"
      "add.u32 a2, %1, 16;\n\t"
      "add.u32 a3, %1, 24;\n\t"
      "add.u64 b1, %2, 2;\n\t"
      "add.u64 b2, %2, 4;\n\t"
      "add.u64 b3, %2, 6;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], b1, %3, t;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], b2, %3, t;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], b3, %3, t;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate_first)
      : "memory");
}
__device__ __forceinline__ void mma4_bf16_ss_warp(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                  uint32_t accumulate_first) {
  asm volatile(
      "{\n\t.reg .pred p, q, t;\n\t"
      ".reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.eq.b32 t, %3, %3;\n\t"
      "add.u64 a1, %1, 2;\n\t"
      "add.u64 a2, %1, 4;\n\t"
      "add.u64 a3, %1, 6;\n\t"
      "add.u64 b1, %2, 2;\n\t"
      "add.u64 b2, %2, 4;\n\t"
      "add.u64 b3, %2, 6;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, t;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, t;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, t;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate_first)
      : "memory");
}
__device__ __forceinline__ void mma_commit_warp(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(smem_u32(bar))
      : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when complete
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- TMA ------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// 2-D tiled load: box lands at smem_dst in the tensor map's swizzle; completes tx bytes on bar
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// 4-D tiled load (coordinates innermost first)
__device__ __forceinline__ void tma_load_4d_warp(void* smem_dst, const void* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n\t}"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// 2-D tiled store shared -> global (bulk async-group completion); one thread issues
__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tmap), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* tmap, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(tmap), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk groups of this thread complete (writes performed), not just their shared-memory reads
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// converged-warp variants: every lane executes, one elected lane issues
__device__ __forceinline__ void tma_load_2d_warp(void* smem_dst, const void* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_warp(uint64_t* bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}"
      ::"r"(smem_u32(bar)), "r"(bytes)
      : "memory");
}

}  // namespace tc
}  // namespace rs
