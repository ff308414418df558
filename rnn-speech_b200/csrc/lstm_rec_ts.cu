// Recurrent LSTM kernels with the recurrent weights resident in TENSOR MEMORY (sm_100a).
//
// Same contract as lstm_rec_tc.cu (tf.nn.dynamic_rnn over BasicLSTMCell and its BPTT,
// /root/reference/models/AcousticModel.py:227-237, :277-278; oracle/model.py), different
// machine mapping.  In lstm_rec_tc.cu the weights are the shared-memory operand of every
// tcgen05.mma, and re-reading them from shared memory each step is what bounds the step
// (one M64 N32 K16 MMA costs ~25 cycles of operand traffic, 144 of them per step).  Here the
// weights are the A operand in TMEM, written once per launch with tcgen05.st:
//
// Forward.  CTA j owns hidden units [16j, 16j+16) = 64 gate rows (row m = 4*unit + gate).  TMEM
// lane 32q+i (i < 16) holds W_hi row 16q+i, lane 32q+16+i holds W_lo of the same row; K runs
// along the columns (two bf16 per 32-bit column).  h_{t-1} arrives by TMA as ONE stacked B
// tile per K-block: rows 0..Bpad-1 = h_hi, rows Bpad..2Bpad-1 = h_lo (N = 2 Bpad).  One M128
// MMA per 16 K-elements then yields all four partial products; the epilogue adds
// hi*hi + hi*lo (same lane, two column ranges) + lo*hi (lane + 16, one shuffle) -- the
// bf16x3 product of tc_common.cuh -- in 48 MMAs of ~33 cycles instead of 144 of ~25.
// K-blocks that do not fit in the 512 TMEM columns (H > 896) stay in shared memory as an
// ordinary SS operand with the same row stacking.
//
// Backward.  dh_{t-1}^T[k, b] = sum_n Wh[k, n] dgates_t[b, n]: M = 128 hidden units k per
// CLUSTER of 8 CTAs, K = 4H split 8 ways over the cluster (CTA s holds Wh[128 rows, its H/2
// columns] in TMEM: 24 MMAs of M128 at H = 768 instead of 192).  The product is bf16x3 like forward's
// (the reference's BPTT is fp32, models/AcousticModel.py:386-401): dgates_t streams in as a stacked B tile
// (rows 0..Bpad-1 = hi plane, Bpad..2Bpad-1 = lo plane), Wh_hi x [dg_hi | dg_lo] is one N = 2 Bpad MMA per 16
// K-elements and Wh_lo -- a second resident A block, in tensor memory as far as the 512 columns go, in shared
// memory beyond -- x dg_hi accumulates onto the first Bpad columns.  The eight partial
// accumulators are reduce-scattered through distributed shared memory with st.async
// (complete_tx on the owner's mbarrier): CTA s of the cluster ends up with the 16 units
// [128i+16s, +16) and does their cell backward.  Each CTA streams only its K-segment of
// dgates_t (24 KB instead of 196 KB).
//
// Both kernels publish their per-step result (h_t planes / dgates_t planes) through a shared
// memory staging tile and 16-byte coalesced global stores, then signal the grid barrier.
// The values forward saves for backward live in a private "blob" whose layout is the thread
// layout of the two epilogues (both own the same cells), so every access is a coalesced float4.
#include "lstm_rec_tc.cuh"
#include "tc_common.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

namespace rs {
namespace {

constexpr int NTHREADS = 320;   // warps 0-7: epilogue (sub-partition w%4, column half w/4); 8: TMA; 9: MMA
constexpr int MAXG = 2;         // 16-column groups per epilogue thread (Bpad <= 64)
constexpr int MAXSLOTS = 16;
constexpr int GKB = 4;          // K-blocks per mbarrier (forward)
constexpr int TSU = 16;         // hidden units per CTA
constexpr int BLOB_ITEMS = 5;   // i, j, f, o, c_t
constexpr int CL = 8;           // backward cluster size (K split)

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// timeline stamps of CTA 0 (rows [0,T)) and of the last CTA (rows [T,2T)); 16 events per step
#define RS_STAMP(dbgp, step, ev) do { if constexpr (STAMP) if ((dbgp) && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) \
    (dbgp)[((size_t)(blockIdx.x ? a.T : 0) + (size_t)(step)) * 16 + (ev)] = gtime(); } while (0)
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// grid barrier wait.  variant bit1: relaxed polls + one acquire fence instead of acquire polls
__device__ __forceinline__ void red_relaxed_add(unsigned* p, unsigned v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// variant bit2: relaxed polls and NO acquire fence (the data is only read through TMA, which is
// issued after -- control-dependent on -- the poll that saw the count)
__device__ __forceinline__ void wait_counter(const unsigned* ctr, unsigned target, int variant) {
  if (variant & 4) {
    while (ld_relaxed_u32(ctr) < target) {}
  } else if (variant & 2) {
    while (ld_relaxed_u32(ctr) < target) {}
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
  } else {
    while (ld_acquire_u32(ctr) < target) {}
  }
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ float pick4(int sel, float a0, float a1, float a2, float a3) {
  const float lo = (sel & 1) ? a1 : a0;
  const float hi = (sel & 1) ? a3 : a2;
  return (sel & 2) ? hi : lo;
}

// ---- cluster / distributed shared memory ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 16-byte store into a peer CTA's shared memory; completes 16 tx bytes on the peer's mbarrier
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, float4 v, uint32_t remote_mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(remote_addr), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)),
                 "r"(__float_as_uint(v.w)), "r"(remote_mbar)
               : "memory");
}

// bulk copy of this CTA's shared memory into a peer's, completing transaction bytes on the peer's mbarrier
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t remote_addr, uint32_t local_addr, uint32_t bytes, uint32_t remote_mbar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(remote_addr), "r"(local_addr), "r"(bytes), "r"(remote_mbar) : "memory");
}
// plain 16-byte store into a peer CTA's shared memory (ordered by a later release at cluster scope)
__device__ __forceinline__ void st_cluster_v4(uint32_t remote_addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(remote_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_release(uint32_t remote_mbar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(tc::smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}

// n TMEM-resident K-blocks (64 elements each = four K16 MMAs) issued back to back: A block i at a0 + 32 i columns,
// B block i at b0 + i * bstep.  Fully unrolled for the usual counts so that no per-block address arithmetic sits
// between two tcgen05.mma.
template <int N>
__device__ __forceinline__ void issue_ts_n(uint32_t d, uint32_t a0, uint64_t b0, uint64_t bstep, uint32_t idesc,
                                           uint32_t acc_first) {
#pragma unroll
  for (int i = 0; i < N; ++i)
    tc::mma4_bf16_ts_warp(d, a0 + (uint32_t)(i * 32), b0 + (uint64_t)i * bstep, idesc, i ? 1u : acc_first);
}
// (only the count of the benchmark shape is unrolled in each kernel: every extra case is a few hundred more instructions
//  in kernels whose hot loops should stay in the instruction cache -- with eight cases the gain was gone)
template <int NFAST>
__device__ __forceinline__ void issue_ts_blocks(int n, uint32_t d, uint32_t a0, uint64_t b0, uint64_t bstep,
                                                uint32_t idesc, uint32_t acc_first) {
  if (n == NFAST) issue_ts_n<NFAST>(d, a0, b0, bstep, idesc, acc_first);
  else
    for (int i = 0; i < n; ++i)
      tc::mma4_bf16_ts_warp(d, a0 + (uint32_t)(i * 32), b0 + (uint64_t)i * bstep, idesc, i ? 1u : acc_first);
}

struct KFwd {
  RecTcFwdArgs a;
  int H, B, Bpad, nslice, slots, nkb, nkb_t, ngroups, ngl, gkb;
  int variant;                  // RS_TS_VARIANT switches, see ts_variant()
  uint32_t kb_bytes;            // one streamed K-block: 2 planes x Bpad rows x 128 B
};

// blob float2 index of (step t, slice j, group gl, item, warp, lane)
__device__ __forceinline__ size_t blob_idx(int t, int nslice, int j, int ngl, int gl, int item, int warp, int lane) {
  return ((((((size_t)t * nslice + j) * ngl + gl) * BLOB_ITEMS + item) * 8 + warp) * 32 + lane);
}

// Epilogue thread layout shared by forward and backward (8 warps): q = warp & 3 is the TMEM
// sub-partition, hf = warp >> 2 picks the 16-column (batch) groups gi = hf, hf + 2.  Lane:
// l16 = lane & 15 -> gate row m = 16q + l16 (unit ul = m >> 2 of the slice, gate g = lane & 3);
// up = lane >> 4 -> columns [8 up, 8 up + 8) of the group.  After the quad exchange the lane owns
// the two cells (unit ul, batch b = 16 gi + 8 up + 2 g + k), k = 0, 1.

// STAMP: the instantiation with the timeline stamps compiled in is launched only when a debug buffer is attached
// VARIANT >= 0: the exchange protocol fixed at compile time (the default's dead branches drop out of the hot loops);
// -1: taken from the arguments at run time
// BPAD > 0: the padded batch fixed at compile time (the batch-group loops of the epilogue unroll without tests)
// HH > 0: hidden size fixed at compile time, with all of K in tensor memory as one TMA group (nkb = nkb_t = gkb = HH/64)
template <bool STAMP, int VARIANT, int BPAD, int HH>
__global__ void __launch_bounds__(NTHREADS, 1)
rec_ts_fwd_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmS,
                  const __grid_constant__ CUtensorMap tmD, KFwd p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[MAXSLOTS], empty_bar[MAXSLOTS], tfull_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(128) __nv_bfloat16 sH[64][2][TSU];          // staged h_t [b][plane][unit]
  __shared__ __align__(128) __nv_bfloat16 sD[64][2][TSU];          // staged dropout(out_t) for the next consumer
  const RecTcFwdArgs& a = p.a;
  const int variant = VARIANT >= 0 ? VARIANT : p.variant;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x;
  const int H = HH > 0 ? HH : p.H, B = p.B, Bpad = BPAD > 0 ? BPAD : p.Bpad, T = a.T;
  const int nkb = HH > 0 ? HH / 64 : p.nkb, nkb_t = HH > 0 ? HH / 64 : p.nkb_t, ngroups = HH > 0 ? 1 : p.ngroups, gkb = HH > 0 ? HH / 64 : p.gkb;
  const int nslots = HH > 0 ? 1 : p.slots, nslice = HH > 0 ? HH / TSU : p.nslice;
  const uint32_t kb_bytes = BPAD > 0 ? 2u * BPAD * 128u : p.kb_bytes;
  const int nkb_s = nkb - nkb_t;
  unsigned char* sA = smem;                                         // [nkb_s][128 rows x 128 B] resident SS K-blocks
  unsigned char* sRing = smem + (size_t)nkb_s * 16384;              // [slots][gkb][2 planes][Bpad*128]
  const uint32_t slot_bytes = (uint32_t)gkb * kb_bytes;
  const bool full_flight = nslots >= ngroups;
  const uint32_t colD = (uint32_t)nkb_t * 32;

  if (threadIdx.x == 0) {
    for (int s = 0; s < nslots; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    tc::mbar_init(&tfull_bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 8) tc::tmem_alloc(&tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp < 4) {
    // ---- resident weights: TMEM lane = threadIdx.x; gate row 16q + (lane & 15), hi plane for lane < 16
    const int row = j * 64 + 16 * warp + (lane & 15);
    const __nv_bfloat16* src = ((lane < 16) ? a.wrec_hi : a.wrec_lo) + (size_t)row * H;
    for (int c0 = 0; c0 < nkb_t * 32; c0 += 8) {
      const uint4 x0 = __ldg(reinterpret_cast<const uint4*>(src + 2 * c0));
      const uint4 x1 = __ldg(reinterpret_cast<const uint4*>(src + 2 * c0 + 8));
      const uint32_t r[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      tc::tmem_st8(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0, r);
    }
    for (int kb = nkb_t; kb < nkb; ++kb) {
      unsigned char* tile = sA + (size_t)(kb - nkb_t) * 16384 + (size_t)threadIdx.x * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(tile + ((c ^ (threadIdx.x & 7)) << 4)) = __ldg(reinterpret_cast<const uint4*>(src + kb * 64 + c * 8));
    }
    tc::tmem_st_wait();
    tc::fence_proxy_async_smem();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) tc::tma_prefetch_desc(&tmH);
    __syncwarp();
    const unsigned per_step = gridDim.x * (unsigned)(Bpad / 8);
    uint32_t git = 0;
    for (int ti = 0; ti < T; ++ti) {
      const int t = a.t0 + ti;
      if (ti > 0) {
        wait_counter(a.barrier, per_step * (unsigned)ti, variant);
        __syncwarp();
        if (!(variant & 8)) tc::fence_proxy_async_all();     // other CTAs' generic-proxy stores of h_{t-1} -> TMA reads
      }
      if (lane == 0) RS_STAMP(a.dbg, ti, 0);
      __syncwarp();
      for (int grp = 0; grp < ngroups; ++grp, ++git) {
        const int s = git % nslots;
        const uint32_t ph = (git / nslots) & 1;
        if (!full_flight) tc::mbar_wait(&empty_bar[s], ph ^ 1);
        tc::mbar_arrive_expect_tx_warp(&full_bar[s], slot_bytes);
        // one box = gkb stacked tiles [kb][plane][Bpad rows][64]; row t*B = slot t = h_{t-1}
        tc::tma_load_4d_warp(sRing + (size_t)s * slot_bytes, &tmH, 0, t * B, 0, grp * gkb, &full_bar[s]);
      }
      if (lane == 0) RS_STAMP(a.dbg, ti, 1);
      __syncwarp();
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer (converged warp)
    const uint32_t idesc = tc::instr_desc_bf16(128, 2 * Bpad);
    const uint64_t dA0 = tc::smem_desc_sw128(tc::smem_u32(sA));
    const uint64_t dring0 = tc::smem_desc_sw128(tc::smem_u32(sRing));
    const uint64_t kb_u = kb_bytes >> 4;
    const uint32_t tmemD = tmem + colD;
    uint32_t git = 0;
    for (int t = 0; t < T; ++t) {
      for (int grp = 0; grp < ngroups; ++grp, ++git) {
        const int s = git % nslots;
        const uint32_t ph = (git / nslots) & 1;
        tc::mbar_wait(&full_bar[s], ph);
        tc::tc_fence_after();
        if (lane == 0 && grp == 0) RS_STAMP(a.dbg, t, 2);
        if (lane == 0 && grp >= 1 && grp <= 3) RS_STAMP(a.dbg, t, 10 + grp);       // 11..13: later groups landed
        __syncwarp();
        {
          // TMEM-resident blocks of this group first (unrolled for the usual counts: the descriptors of all blocks are
          // ready before the first issue -- a rolled loop spent ~16 cycles per MMA on them, 4.64 -> 4.26 ms per layer
          // at cfg-2), then the blocks that live in shared memory
          const int kb0 = grp * gkb;
          const int nts = min(max(nkb_t - kb0, 0), gkb);
          const uint64_t dbase = dring0 + (uint64_t)s * (slot_bytes >> 4);
          issue_ts_blocks<12>(nts, tmemD, tmem + (uint32_t)(kb0 * 32), dbase, kb_u, idesc, (uint32_t)(kb0 != 0));   // 12: H = 768
          for (int i = nts; i < gkb; ++i) {
            const int kb = kb0 + i;
            tc::mma4_bf16_ss_warp(tmemD, dA0 + (uint64_t)(kb - nkb_t) * (16384 >> 4), dbase + (uint64_t)i * kb_u, idesc,
                                  (uint32_t)(kb != 0));
          }
        }
        if (lane == 0 && grp == 0) RS_STAMP(a.dbg, t, 14);                          // group 0 MMAs issued
        __syncwarp();
        if (!full_flight) tc::mma_commit_warp(&empty_bar[s]);
      }
      tc::mma_commit_warp(&tfull_bar);
      if (lane == 0) RS_STAMP(a.dbg, t, 3);
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3, hf = warp >> 2, l16 = lane & 15, up = lane >> 4;
    const int m = q * 16 + l16;                 // gate row (hi copy in lanes 0..15, lo copy in 16..31)
    const int g = lane & 3;                     // gate of this row: 0 i, 1 j, 2 f, 3 o
    const int ul = m >> 2;                      // unit inside the slice
    const int unit = j * TSU + ul;
    const int ng = Bpad / 16;
    const uint32_t tmemD = tmem + colD + ((uint32_t)(q * 32) << 16);
    float c[MAXG][2], hl[MAXG][2];
    int lenr[MAXG][2];
#pragma unroll
    for (int gl = 0; gl < MAXG; ++gl)
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int b = (hf + 2 * gl) * 16 + 8 * up + 2 * g + k;
        const bool ok = (hf + 2 * gl) < ng && b < B;
        c[gl][k] = ok ? a.c0[(size_t)b * H + unit] : 0.f;
        hl[gl][k] = ok ? a.h0[(size_t)b * H + unit] : 0.f;
        lenr[gl][k] = ok ? a.len[b] : 0;
      }
    const float fbias = (g == 2) ? 1.0f : 0.0f;          // forget_bias
    const float pre = (g == 1) ? 2.0f : 1.0f;            // tanh(x) = 2*sigmoid(2x) - 1
    const float post_m = (g == 1) ? 2.0f : 1.0f, post_a = (g == 1) ? -1.0f : 0.0f;
    float2* blob = reinterpret_cast<float2*>(a.gates);
    const unsigned nstore = (unsigned)(4 * Bpad);         // 16-byte chunks of the staged h tile

    for (int ti = 0; ti < T; ++ti) {
      const int t = a.t0 + ti;
      // hoisted input projection for this step (independent of the recurrence: issue early)
      float4 gxv[MAXG][2];
#pragma unroll
      for (int gl = 0; gl < MAXG; ++gl) {
        const int gi = hf + 2 * gl;
        if (gi < ng) {
          const float4* gp = reinterpret_cast<const float4*>(
              a.gx + (((size_t)t * nslice + j) * 64 + m) * Bpad + gi * 16 + 8 * up);
          gxv[gl][0] = __ldg(gp);
          gxv[gl][1] = __ldg(gp + 1);
        } else {
          gxv[gl][0] = gxv[gl][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      float2 keep[MAXG][BLOB_ITEMS];
      float hout[MAXG][2];
      tc::mbar_wait(&tfull_bar, (uint32_t)(ti & 1));
      tc::tc_fence_after();
      if (threadIdx.x == 0) RS_STAMP(a.dbg, ti, 4);
#pragma unroll
      for (int gl = 0; gl < MAXG; ++gl) {
        const int gi = hf + 2 * gl;
        if (gi < ng) {                           // warp-uniform
          float v[16], w[16];
          tc::tmem_ld16(tmemD + (uint32_t)(gi * 16), v);             // hi rows: W_hi h_hi ; lo rows: W_lo h_hi
          tc::tmem_ld16(tmemD + (uint32_t)(Bpad + gi * 16), w);      // hi rows: W_hi h_lo
          tc::tmem_ld_wait();
          // the hi-row lane finishes columns 0..7, its lo-row partner (lane ^ 16) columns 8..15
          float act[8];
          const float xs[8] = {gxv[gl][0].x, gxv[gl][0].y, gxv[gl][0].z, gxv[gl][0].w,
                               gxv[gl][1].x, gxv[gl][1].y, gxv[gl][1].z, gxv[gl][1].w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float s_a = v[i] + w[i], s_b = v[8 + i] + w[8 + i];
            const float got = __shfl_xor_sync(0xffffffffu, up ? v[i] : s_b, 16);
            const float z = ((up ? (got + v[8 + i]) : (s_a + got)) + xs[i] + fbias) * pre;
            act[i] = fmaf(fast_sigmoid(z), post_m, post_a);
          }
          // quad exchange: this lane owns columns 2g + k; gate tau comes from lane g ^ (g ^ tau)
          float own[2], rcv[3][2];
#pragma unroll
          for (int k = 0; k < 2; ++k) own[k] = pick4(g, act[k], act[2 + k], act[4 + k], act[6 + k]);
#pragma unroll
          for (int d = 1; d < 4; ++d)
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const float snd = pick4(g ^ d, act[k], act[2 + k], act[4 + k], act[6 + k]);
              rcv[d - 1][k] = __shfl_xor_sync(0xffffffffu, snd, d);
            }
          float gv[4][2], cn[2];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const float ig = pick4(g ^ 0, own[k], rcv[0][k], rcv[1][k], rcv[2][k]);
            const float jg = pick4(g ^ 1, own[k], rcv[0][k], rcv[1][k], rcv[2][k]);
            const float fg = pick4(g ^ 2, own[k], rcv[0][k], rcv[1][k], rcv[2][k]);
            const float og = pick4(g ^ 3, own[k], rcv[0][k], rcv[1][k], rcv[2][k]);
            const int b = gi * 16 + 8 * up + 2 * g + k;
            const float c_new = c[gl][k] * fg + ig * jg;
            const float h_new = fast_tanh(c_new) * og;
            const bool valid = t < lenr[gl][k];
            if (valid) { c[gl][k] = c_new; hl[gl][k] = h_new; }
            __nv_bfloat16 hh, hlo;
            tc::split_bf16(valid ? h_new : 0.f, hh, hlo);
            sH[b][0][ul] = hh;
            sH[b][1][ul] = hlo;
            hout[gl][k] = valid ? h_new : 0.f;
            gv[0][k] = ig; gv[1][k] = jg; gv[2][k] = fg; gv[3][k] = og; cn[k] = c_new;
          }
#pragma unroll
          for (int it = 0; it < 4; ++it) keep[gl][it] = make_float2(gv[it][0], gv[it][1]);
          keep[gl][4] = make_float2(cn[0], cn[1]);
        }
      }
      if (threadIdx.x == 0) RS_STAMP(a.dbg, t, 5);
      tc::tc_fence_before();
      if (variant & 16) tc::fence_proxy_async_smem();
      // (thread 0 arrives here after its previous bulk groups -- the hop store of step t-1 among them -- have read
      //  their shared-memory source, so sD may be rewritten behind this barrier)
      if (a.drop_hi && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      epi_bar_sync();
      if (variant & 16) {
        // publish h_t (both planes) with one TMA store; the bulk-group wait returns when the writes are performed
        if (threadIdx.x == 0) {
          tc::tma_store_3d(&tmS, &sH[0][0][0], j * TSU, 0, (t + 1) * B);
          tc::bulk_commit();
          RS_STAMP(a.dbg, t, 9);
          tc::bulk_wait_all();
          RS_STAMP(a.dbg, t, 10);
          if (ti + 1 < T) {
            if ((variant & 32) && !(variant & 64)) { if (variant & 256) tc::fence_proxy_async_global(); else tc::fence_proxy_async_all(); }
            if ((variant & 32) && !(variant & 128)) red_release_add(a.barrier, (unsigned)(Bpad / 8));
            else red_relaxed_add(a.barrier, (unsigned)(Bpad / 8));
          }
        }
      } else if (threadIdx.x < nstore) {
        // publish h_t: 16-byte coalesced stores of the staged tile, one release per storing warp
        const int b = (int)threadIdx.x >> 2, pl = ((int)threadIdx.x >> 1) & 1, half = (int)threadIdx.x & 1;
        if (b < B) {
          const uint4 x = *reinterpret_cast<const uint4*>(&sH[b][pl][half * 8]);
          __nv_bfloat16* dst = (pl ? a.h_lo : a.h_hi) + ((size_t)(t + 1) * B + b) * a.h_ld + j * TSU + half * 8;
          *reinterpret_cast<uint4*>(dst) = x;
        }
        if (threadIdx.x == 0) RS_STAMP(a.dbg, t, 9);
        if (!(variant & 1)) tc::fence_proxy_async_all();
        __syncwarp();
        if (threadIdx.x == 0) RS_STAMP(a.dbg, t, 10);
        if (lane == 0 && ti + 1 < T) red_release_add(a.barrier, 1u);
      }
      if (threadIdx.x == 0) RS_STAMP(a.dbg, t, 6);
      // reserve for backward (not on the critical path of the recurrence)
      if (blob) {
#pragma unroll
        for (int gl = 0; gl < MAXG; ++gl) {
          if (hf + 2 * gl < ng) {
#pragma unroll
            for (int it = 0; it < BLOB_ITEMS; ++it) blob[blob_idx(t, nslice, j, p.ngl, gl, it, warp, lane)] = keep[gl][it];
          }
        }
      }
      // fused hop (off the critical path of the recurrence): dropout(out_t) planes for the next layer / the output dense
      if (a.drop_hi) {
#pragma unroll
        for (int gl = 0; gl < MAXG; ++gl) {
          const int gi = hf + 2 * gl;
          if (gi < ng) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int b = gi * 16 + 8 * up + 2 * g + k;
              float v = hout[gl][k];
              const unsigned long long idx = ((unsigned long long)t * B + b) * H + unit;
              if (a.drop_thr_a != 0xffffffffu) v = dropout_keep(a.drop_key, a.drop_sa, idx, a.drop_thr_a) ? v * a.drop_inv_a : 0.f;
              if (a.drop_thr_b != 0xffffffffu) v = dropout_keep(a.drop_key, a.drop_sb, idx, a.drop_thr_b) ? v * a.drop_inv_b : 0.f;
              __nv_bfloat16 dh_, dl_;
              tc::split_bf16(v, dh_, dl_);
              sD[b][0][ul] = dh_;
              sD[b][1][ul] = dl_;
            }
          }
        }
        tc::fence_proxy_async_smem();
        epi_bar_sync();
        if (threadIdx.x == 0) {
          tc::tma_store_3d(&tmD, &sD[0][0][0], j * TSU, 0, t * B);
          tc::bulk_commit();
        }
      }
      if (threadIdx.x == 0) RS_STAMP(a.dbg, t, 7);
      if constexpr (STAMP)
        if (threadIdx.x == 0 && a.dbg && blockIdx.x == 0) a.dbg[(size_t)t * 16 + 15] = (unsigned long long)clock64();   // SM clock vs globaltimer
    }
    if (a.drop_hi && threadIdx.x == 0) tc::bulk_wait_all();          // the last hop store is performed before the kernel ends
#pragma unroll
    for (int gl = 0; gl < MAXG; ++gl)
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int b = (hf + 2 * gl) * 16 + 8 * up + 2 * g + k;
        if ((hf + 2 * gl) < ng && b < B) {
          if (a.cT) a.cT[(size_t)b * H + unit] = c[gl][k];
          if (a.hT) a.hT[(size_t)b * H + unit] = hl[gl][k];
        }
      }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------
// Forward, two chains.  The utterances of a mini-batch never interact inside the recurrence, so a batch of 17..32 is
// run as TWO independent recurrences over 16 rows each (chain X = batch rows [16X, 16X + 16)) that share the CTA's
// resident weights: each chain has its own producer warp, MMA warp, four epilogue warps, accumulator columns, shared
// memory slot, mbarriers and grid-barrier counter.  A step of one chain is the same dependent sequence as before
// (counter seen -> TMA box -> 48 MMAs -> epilogue -> publish -> release), but while one chain waits for its exchange
// through L2 (~60 % of a step) the other one computes: the tensor pipe, the TMA unit and the epilogue issue slots are
// idle most of a step in the one-chain kernel.  Per chain N = 2 x 16 (hi | lo planes of 16 rows), its box is 48 KB.
// Requires Bpad == 32 and one TMA group per step (all of K in one box); other shapes use rec_ts_fwd_kernel.
// ------------------------------------------------------------------------------------
constexpr int NTHREADS2 = 384;   // warps 0-3 / 4-7: epilogue of chain 0 / 1; 8, 9: TMA + MMA of chain 0; 10, 11: of chain 1
constexpr int CHB = 16;          // batch rows per chain

struct KFwd2 {
  RecTcFwdArgs a;
  int H, B, nslice, nkb, nkb_t;
  int rows1;                      // valid batch rows of chain 1 (B - 16)
};

__device__ __forceinline__ void chain_bar_sync(int chain) {
  asm volatile("bar.sync %0, 128;" ::"r"(1 + chain) : "memory");
}

template <int HH>
__global__ void __launch_bounds__(NTHREADS2, 1)
rec_ts_fwd2_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmS0,
                   const __grid_constant__ CUtensorMap tmS1, const __grid_constant__ CUtensorMap tmD0,
                   const __grid_constant__ CUtensorMap tmD1, KFwd2 p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[2], tfull_bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(128) __nv_bfloat16 sH[2][CHB][2][TSU];       // staged h_t per chain [b][plane][unit]
  __shared__ __align__(128) __nv_bfloat16 sD[2][CHB][2][TSU];       // staged dropout(out_t) per chain
  const RecTcFwdArgs& a = p.a;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x;
  const int H = HH > 0 ? HH : p.H, B = p.B, T = a.T;
  const int nkb = HH > 0 ? HH / 64 : p.nkb, nkb_t = HH > 0 ? HH / 64 : p.nkb_t, nslice = HH > 0 ? HH / TSU : p.nslice;
  const int nkb_s = nkb - nkb_t;
  constexpr uint32_t kb_bytes = 2u * CHB * 128u;                     // one K-block of a chain: hi rows, lo rows
  const uint32_t slot_bytes = (uint32_t)nkb * kb_bytes;
  unsigned char* sA = smem;                                          // [nkb_s][128 rows x 128 B] weight blocks outside TMEM
  unsigned char* sRing = smem + (size_t)nkb_s * 16384;               // [2 chains][nkb][2 planes][16 rows x 128 B]
  const uint32_t colD = (uint32_t)nkb_t * 32;

  if (threadIdx.x == 0) {
    for (int x = 0; x < 2; ++x) { tc::mbar_init(&full_bar[x], 1); tc::mbar_init(&tfull_bar[x], 1); }
    tc::fence_mbar_init();
  }
  if (warp == 8) tc::tmem_alloc(&tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp < 4) {
    // resident weights, as in rec_ts_fwd_kernel: TMEM lane = threadIdx.x; gate row 16q + (lane & 15), hi plane for lane < 16
    const int row = j * 64 + 16 * warp + (lane & 15);
    const __nv_bfloat16* src = ((lane < 16) ? a.wrec_hi : a.wrec_lo) + (size_t)row * H;
    for (int c0 = 0; c0 < nkb_t * 32; c0 += 8) {
      const uint4 x0 = __ldg(reinterpret_cast<const uint4*>(src + 2 * c0));
      const uint4 x1 = __ldg(reinterpret_cast<const uint4*>(src + 2 * c0 + 8));
      const uint32_t r[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      tc::tmem_st8(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0, r);
    }
    for (int kb = nkb_t; kb < nkb; ++kb) {
      unsigned char* tile = sA + (size_t)(kb - nkb_t) * 16384 + (size_t)threadIdx.x * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(tile + ((c ^ (threadIdx.x & 7)) << 4)) = __ldg(reinterpret_cast<const uint4*>(src + kb * 64 + c * 8));
    }
    tc::tmem_st_wait();
    tc::fence_proxy_async_smem();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();

  if (warp == 8 || warp == 10) {
    // ------------------------------------------------------------------ TMA producer of chain X
    const int X = (warp - 8) >> 1;
    if (lane == 0) tc::tma_prefetch_desc(&tmH);
    __syncwarp();
    const unsigned per_step = gridDim.x * 2u;                        // every CTA adds CHB / 8 per chain and step
    const unsigned* ctr = a.barrier + 32 * X;
    for (int ti = 0; ti < T; ++ti) {
      const int t = a.t0 + ti;
      if (ti > 0) {
        while (ld_relaxed_u32(ctr) < per_step * (unsigned)ti) {}
        __syncwarp();
      }
      tc::mbar_arrive_expect_tx_warp(&full_bar[X], slot_bytes);
      // one box = nkb stacked tiles [kb][plane][16 rows][64]; row t*B + 16X = h_{t-1} of the chain's first utterance
      tc::tma_load_4d_warp(sRing + (size_t)X * slot_bytes, &tmH, 0, t * B + CHB * X, 0, 0, &full_bar[X]);
    }
  } else if (warp == 9 || warp == 11) {
    // ------------------------------------------------------------------ MMA issuer of chain X (converged warp)
    const int X = (warp - 9) >> 1;
    const uint32_t idesc = tc::instr_desc_bf16(128, 2 * CHB);
    const uint64_t dA0 = tc::smem_desc_sw128(tc::smem_u32(sA));
    const uint64_t dB0 = tc::smem_desc_sw128(tc::smem_u32(sRing + (size_t)X * slot_bytes));
    constexpr uint64_t kb_u = kb_bytes >> 4;
    const uint32_t tmemD = tmem + colD + (uint32_t)(2 * CHB * X);
    for (int t = 0; t < T; ++t) {
      tc::mbar_wait(&full_bar[X], (uint32_t)(t & 1));
      tc::tc_fence_after();
      issue_ts_blocks<12>(nkb_t, tmemD, tmem, dB0, kb_u, idesc, 0u);                                       // 12: H = 768
      for (int kb = nkb_t; kb < nkb; ++kb)
        tc::mma4_bf16_ss_warp(tmemD, dA0 + (uint64_t)(kb - nkb_t) * (16384 >> 4), dB0 + (uint64_t)kb * kb_u, idesc, (uint32_t)(kb != 0));
      tc::mma_commit_warp(&tfull_bar[X]);
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps of chain X = warp >> 2
    const int q = warp & 3, X = warp >> 2, l16 = lane & 15, up = lane >> 4;
    const int m = q * 16 + l16;                 // gate row (hi copy in lanes 0..15, lo copy in 16..31)
    const int g = lane & 3;                     // gate of this row: 0 i, 1 j, 2 f, 3 o
    const int ul = m >> 2;                      // unit inside the slice
    const int unit = j * TSU + ul;
    const uint32_t tmemD = tmem + colD + (uint32_t)(2 * CHB * X) + ((uint32_t)(q * 32) << 16);
    const bool leader = (threadIdx.x & 127) == 0;
    const CUtensorMap* tmS = X ? &tmS1 : &tmS0;
    const CUtensorMap* tmD = X ? &tmD1 : &tmD0;
    unsigned* ctr = a.barrier + 32 * X;
    float c[2], hl[2];
    int lenr[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int b = X * CHB + 8 * up + 2 * g + k;
      const bool ok = b < B;
      c[k] = ok ? a.c0[(size_t)b * H + unit] : 0.f;
      hl[k] = ok ? a.h0[(size_t)b * H + unit] : 0.f;
      lenr[k] = ok ? a.len[b] : 0;
    }
    const float fbias = (g == 2) ? 1.0f : 0.0f;          // forget_bias
    const float pre = (g == 1) ? 2.0f : 1.0f;            // tanh(x) = 2*sigmoid(2x) - 1
    const float post_m = (g == 1) ? 2.0f : 1.0f, post_a = (g == 1) ? -1.0f : 0.0f;
    float2* blob = reinterpret_cast<float2*>(a.gates);

    for (int ti = 0; ti < T; ++ti) {
      const int t = a.t0 + ti;
      // hoisted input projection for this step (independent of the recurrence: issue early); Bpad = 32, group X
      const float4* gp = reinterpret_cast<const float4*>(a.gx + (((size_t)t * nslice + j) * 64 + m) * 32 + X * 16 + 8 * up);
      const float4 gx0 = __ldg(gp), gx1 = __ldg(gp + 1);
      tc::mbar_wait(&tfull_bar[X], (uint32_t)(ti & 1));
      tc::tc_fence_after();
      float v[16], w[16];
      tc::tmem_ld16(tmemD, v);                               // hi rows: W_hi h_hi ; lo rows: W_lo h_hi
      tc::tmem_ld16(tmemD + (uint32_t)CHB, w);               // hi rows: W_hi h_lo
      tc::tmem_ld_wait();
      float act[8];
      const float xs[8] = {gx0.x, gx0.y, gx0.z, gx0.w, gx1.x, gx1.y, gx1.z, gx1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float s_a = v[i] + w[i], s_b = v[8 + i] + w[8 + i];
        const float got = __shfl_xor_sync(0xffffffffu, up ? v[i] : s_b, 16);
        const float z = ((up ? (got + v[8 + i]) : (s_a + got)) + xs[i] + fbias) * pre;
        act[i] = fmaf(fast_sigmoid(z), post_m, post_a);
      }
      float own[2], rcv[3][2];
#pragma unroll
      for (int k = 0; k < 2; ++k) own[k] = pick4(g, act[k], act[2 + k], act[4 + k], act[6 + k]);
#pragma unroll
      for (int d = 1; d < 4; ++d)
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const float snd = pick4(g ^ d, act[k], act[2 + k], act[4 + k], act[6 + k]);
          rcv[d - 1][k] = __shfl_xor_sync(0xffffffffu, snd, d);
        }
      float2 keep[BLOB_ITEMS];
      float gv[4][2], cn[2], hout[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float ig = pick4(g ^ 0, own[k], rcv[0][k], rcv[1][k], rcv[2][k]);
        const float jg = pick4(g ^ 1, own[k], rcv[0][k], rcv[1][k], rcv[2][k]);
        const float fg = pick4(g ^ 2, own[k], rcv[0][k], rcv[1][k], rcv[2][k]);
        const float og = pick4(g ^ 3, own[k], rcv[0][k], rcv[1][k], rcv[2][k]);
        const int bl = 8 * up + 2 * g + k;                   // row inside the chain
        const float c_new = c[k] * fg + ig * jg;
        const float h_new = fast_tanh(c_new) * og;
        const bool valid = t < lenr[k];
        if (valid) { c[k] = c_new; hl[k] = h_new; }
        __nv_bfloat16 hh, hlo;
        tc::split_bf16(valid ? h_new : 0.f, hh, hlo);
        sH[X][bl][0][ul] = hh;
        sH[X][bl][1][ul] = hlo;
        hout[k] = valid ? h_new : 0.f;
        gv[0][k] = ig; gv[1][k] = jg; gv[2][k] = fg; gv[3][k] = og; cn[k] = c_new;
      }
#pragma unroll
      for (int it = 0; it < 4; ++it) keep[it] = make_float2(gv[it][0], gv[it][1]);
      keep[4] = make_float2(cn[0], cn[1]);
      tc::tc_fence_before();
      tc::fence_proxy_async_smem();
      // (the chain's leader arrives after its earlier bulk groups -- the hop store of step t-1 -- have read sD)
      if (a.drop_hi && leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      chain_bar_sync(X);
      if (leader) {
        // publish the chain's h_t (both planes) with one TMA store; completion of the bulk group + a release on the
        // chain's counter (see ts_variant(): the writer side of the exchange protocol)
        tc::tma_store_3d(tmS, &sH[X][0][0][0], j * TSU, 0, (t + 1) * B + CHB * X);
        tc::bulk_commit();
        tc::bulk_wait_all();
        if (ti + 1 < T) {
          tc::fence_proxy_async_global();
          red_release_add(ctr, 2u);
        }
      }
      // reserve for backward (not on the critical path of the recurrence)
      if (blob) {
#pragma unroll
        for (int it = 0; it < BLOB_ITEMS; ++it) blob[blob_idx(t, nslice, j, 1, 0, it, warp, lane)] = keep[it];
      }
      // fused hop: dropout(out_t) planes for the next layer / the output dense
      if (a.drop_hi) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int bl = 8 * up + 2 * g + k;
          float dv = hout[k];
          const unsigned long long idx = ((unsigned long long)t * B + (X * CHB + bl)) * H + unit;
          if (a.drop_thr_a != 0xffffffffu) dv = dropout_keep(a.drop_key, a.drop_sa, idx, a.drop_thr_a) ? dv * a.drop_inv_a : 0.f;
          if (a.drop_thr_b != 0xffffffffu) dv = dropout_keep(a.drop_key, a.drop_sb, idx, a.drop_thr_b) ? dv * a.drop_inv_b : 0.f;
          __nv_bfloat16 dh_, dl_;
          tc::split_bf16(dv, dh_, dl_);
          sD[X][bl][0][ul] = dh_;
          sD[X][bl][1][ul] = dl_;
        }
        tc::fence_proxy_async_smem();
        chain_bar_sync(X);
        if (leader) {
          tc::tma_store_3d(tmD, &sD[X][0][0][0], j * TSU, 0, t * B + CHB * X);
          tc::bulk_commit();
        }
      }
    }
    if (a.drop_hi && leader) tc::bulk_wait_all();
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int b = X * CHB + 8 * up + 2 * g + k;
      if (b < B) {
        if (a.cT) a.cT[(size_t)b * H + unit] = c[k];
        if (a.hT) a.hT[(size_t)b * H + unit] = hl[k];
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------
// Forward, two chains, VALIDATED exchange: no release, no bulk-group wait, no second pass over the data.
//
// The exchange of rec_ts_fwd2_kernel costs four dependent hops per step: TMA store + bulk-group wait, proxy fence +
// red.release (a MEMBAR behind every other store of the SM: ~0.6 us), the consumers' poll of the counter, their TMA load.
// Here the writer side is two instructions: a 16-byte `st.relaxed.gpu` per (row, plane, half) of the CTA's 16 units and a
// RELAXED `red` on the chain's counter.  Nothing orders the two, so the counter is only a hint that the tile is probably
// there; the proof is in the data.  Before a forward pass the host fills slots 1..T of the h planes with the bit pattern
// 0xFFFF -- a bf16 NaN that cvt.rn.bf16 never produces (NaN converts to 0x7FFF).  A consumer polls the counter and fetches
// the chain's 48 KB tile with ONE TMA box as before.  A tile row that still holds fill pattern anywhere -- a store overtaken
// by its producer's `red`, or not yet visible to the async proxy -- makes that batch column of the accumulator NaN in every
// gate row (NaN x w = NaN, also for w = 0): THE TENSOR CORE VALIDATES THE TILE, the epilogue only tests the eight
// pre-activations it computes anyway and votes in the barrier it needs anyway before publishing.  A NaN vote takes the slow
// path: the four epilogue warps scan the landed tile for the fill pattern; if it is there, the producer warp fetches the
// tile again and the MMA warp multiplies again (accumulate = 0 overwrites), nothing of the attempt having been kept or
// published; if not, the NaN is the model's own (diverged training) and the step goes on.
// (tests/test_gpu_model.py::test_exchange_fault_injection delays half of every publish behind its `red` and checks that the
// retries happen and that nothing changes bit for bit.)
//
// The chains take turns: see the producer warps.
// Requires Bpad <= 32 and all of K in one tile.  A batch of at most 16 rows runs as one chain (the second chain's warps idle).
// ------------------------------------------------------------------------------------
constexpr uint32_t kFill = 0xffffffffu;

__device__ __forceinline__ void st_relaxed_v4(void* p, const uint4 v) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
  return v;
}
// barrier over `nthreads` threads (named barrier `id`) + vote: true for every thread iff the predicate holds for at least one
__device__ __forceinline__ bool bar_any(int id, int nthreads, bool pred) {
  uint32_t r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %1, 0;\n\t"
      "barrier.cta.red.or.aligned.pred q, %2, %3, p;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t}"
      : "=r"(r) : "r"((uint32_t)pred), "r"(id), "r"(nthreads) : "memory");
  return r != 0;
}
// does the landed tile [bytes] at saddr hold the fill pattern in a row < rows?  Called by nthr threads (tid = index among
// them); the tile is [K-blocks][2 planes][tile_rows x 128 B], so 16-byte chunk c lies in row (c >> 3) % tile_rows.
__device__ __forceinline__ bool tile_has_fill(uint32_t saddr, uint32_t bytes, int tile_rows, int rows, int tid, int nthr) {
  uint32_t low = 1u;                                       // min of ~word: 0 iff a word is the fill pattern
  for (uint32_t c = (uint32_t)tid; c < bytes / 16; c += (uint32_t)nthr) {
    if ((int)((c >> 3) % (uint32_t)tile_rows) < rows) {
      const uint4 x = ld_shared_v4(saddr + c * 16u);
      low = min(min(low, min(~x.x, ~x.y)), min(~x.z, ~x.w));
    }
  }
  return low == 0u;
}

struct KFwd3 {
  RecTcFwdArgs a;
  int H, B, Bpad, nslice, nkb, nkb_t;
  int turns;                      // 1: the chains take turns at the TMA port and the tensor pipe (see the producer warps)
  int fault;                      // STAMP instantiation only: delay half of every publish behind its hint (test of the retry path)
  int endbar;                     // 1: the chain's four epilogue warps meet once more at the end of a step
};

// chain X stamps (STAMP instantiation, tests/gpu_diag.py xchg): event 8X + 0 counter seen, 1 fetch issued, 2 tile landed, 3 MMAs done,
// 4 cell math done, 5 published, 6 end of step, 7 attempts so far
#define RS_STAMP3(step, ev, val) do { if constexpr (STAMP) if (a.dbg && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) \
    a.dbg[((size_t)(blockIdx.x ? a.T : 0) + (size_t)(step)) * 16 + 8 * X + (ev)] = (val); } while (0)
template <int HH, bool STAMP>
__global__ void __launch_bounds__(NTHREADS2, 1)
rec_ts_fwd3_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmD0,
                   const __grid_constant__ CUtensorMap tmD1, KFwd3 p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[2], tfull_bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ uint32_t retry_req[2];                                 // fetches of the current step asked for again, per chain (running count)
  __shared__ uint32_t quit[2];                                      // the chain's last step is done
  __shared__ uint32_t steps_done[2];                                // steps of the launch whose MMAs have completed, per chain
  __shared__ uint32_t tiles_landed[2];                              // fetches of the launch that have landed, per chain
  __shared__ __align__(128) __nv_bfloat16 sH[2][CHB][2][TSU];       // staged h_t per chain [b][plane][unit]
  __shared__ __align__(128) __nv_bfloat16 sD[2][CHB][2][TSU];       // staged dropout(out_t) per chain
  const RecTcFwdArgs& a = p.a;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x;
  const int H = HH > 0 ? HH : p.H, B = p.B, T = a.T;
  const int nkb = HH > 0 ? HH / 64 : p.nkb, nkb_t = HH > 0 ? HH / 64 : p.nkb_t, nslice = HH > 0 ? HH / TSU : p.nslice;
  const int nkb_s = nkb - nkb_t;
  constexpr uint32_t kb_bytes = 2u * CHB * 128u;                     // one K-block of a chain: hi rows, lo rows
  const uint32_t slot_bytes = (uint32_t)nkb * kb_bytes;
  unsigned char* sA = smem;                                          // [nkb_s][128 rows x 128 B] weight blocks outside TMEM
  unsigned char* sRing = smem + (size_t)nkb_s * 16384;               // [2 chains][nkb][2 planes][16 rows x 128 B]
  const uint32_t colD = (uint32_t)nkb_t * 32;

  if (threadIdx.x == 0) {
    for (int x = 0; x < 2; ++x) { tc::mbar_init(&full_bar[x], 1); tc::mbar_init(&tfull_bar[x], 1); retry_req[x] = 0; quit[x] = 0; steps_done[x] = 0; tiles_landed[x] = 0; }
    tc::fence_mbar_init();
  }
  if (warp == 8) tc::tmem_alloc(&tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp < 4) {
    // resident weights, as in rec_ts_fwd_kernel: TMEM lane = threadIdx.x; gate row 16q + (lane & 15), hi plane for lane < 16
    const int row = j * 64 + 16 * warp + (lane & 15);
    const __nv_bfloat16* src = ((lane < 16) ? a.wrec_hi : a.wrec_lo) + (size_t)row * H;
    for (int c0 = 0; c0 < nkb_t * 32; c0 += 8) {
      const uint4 x0 = __ldg(reinterpret_cast<const uint4*>(src + 2 * c0));
      const uint4 x1 = __ldg(reinterpret_cast<const uint4*>(src + 2 * c0 + 8));
      const uint32_t r[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      tc::tmem_st8(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0, r);
    }
    for (int kb = nkb_t; kb < nkb; ++kb) {
      unsigned char* tile = sA + (size_t)(kb - nkb_t) * 16384 + (size_t)threadIdx.x * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(tile + ((c ^ (threadIdx.x & 7)) << 4)) = __ldg(reinterpret_cast<const uint4*>(src + kb * 64 + c * 8));
    }
    tc::tmem_st_wait();
    tc::fence_proxy_async_smem();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();

  // a batch of at most 16 rows is ONE chain: the roles of chain 1 have nothing to do (and nobody to take turns with)
  const bool two = B > CHB;
  const int my_chain = warp < 8 ? (warp >> 2) : ((warp - 8) >> 1);
  if (my_chain == 1 && !two) {
    // (falls through to the common exit)
  } else if (warp == 8 || warp == 10) {
    // ------------------------------------------------------------------ TMA producer of chain X
    const int X = (warp - 8) >> 1;
    if (lane == 0) tc::tma_prefetch_desc(&tmH);
    __syncwarp();
    const unsigned per_step = gridDim.x * 2u;                        // every CTA adds 2 per chain and step
    const unsigned* ctr = a.barrier + 32 * X;
    uint32_t served = 0;                                             // retry requests answered
    // one box = nkb stacked tiles [kb][plane][16 rows][64]; row t*B + 16X = h_{t-1} of the chain's first utterance
    auto fetch = [&](int t) {
      tc::mbar_arrive_expect_tx_warp(&full_bar[X], slot_bytes);
      tc::tma_load_4d_warp(sRing + (size_t)X * slot_bytes, &tmH, 0, t * B + CHB * X, 0, 0, &full_bar[X]);
    };
    // the epilogue found fill pattern in the tile of step t (after that attempt's MMAs had completed): fetch it again
    auto serve = [&](int t) {
      if (*reinterpret_cast<volatile uint32_t*>(&retry_req[X]) != served) { ++served; __syncwarp(); fetch(t); }
    };
    for (int ti = 0; ti < T; ++ti) {
      const int t = a.t0 + ti;
      if (ti > 0) {
        uint32_t spins = 0;
        while (ld_relaxed_u32(ctr) < per_step * (unsigned)ti) { serve(t - 1); if (++spins > (1u << 26)) __trap(); }
        __syncwarp();
      }
      if (lane == 0) RS_STAMP3(ti, 0, gtime());
      __syncwarp();
      if (p.turns && two) {
        // Two chains that start together stay together (measured: 6 ns apart after 900 steps, and they fall back into
        // lockstep from either side of half a period): both tiles cross the SM's port at the same time and 2 x 48 MMAs
        // queue on one tensor pipe, so that a step of either takes as long as the step of a 32-row batch.  So they take
        // turns at the port: chain 1 fetches the tile of step ti when chain 0's tile of step ti has landed (turns == 1) or
        // its MMAs are complete (turns == 2), chain 0 the tile of step ti + 1 when chain 1's of step ti has -- one chain
        // computes while the other one exchanges.  (Retries only make the counts run ahead: a weaker constraint.)
        const uint32_t need = (uint32_t)(ti + X);
        const volatile uint32_t* other = p.turns == 2 ? &steps_done[X ^ 1] : &tiles_landed[X ^ 1];
        while (*other < need) {}
        __syncwarp();
      }
      fetch(t);
      if (lane == 0) RS_STAMP3(ti, 1, gtime());
      __syncwarp();
    }
    while (*reinterpret_cast<volatile uint32_t*>(&quit[X]) == 0) serve(a.t0 + T - 1);
  } else if (warp == 9 || warp == 11) {
    // ------------------------------------------------------------------ MMA issuer of chain X (converged warp): one batch per fetch
    const int X = (warp - 9) >> 1;
    const uint32_t idesc = tc::instr_desc_bf16(128, 2 * CHB);
    const uint64_t dA0 = tc::smem_desc_sw128(tc::smem_u32(sA));
    const uint64_t dB0 = tc::smem_desc_sw128(tc::smem_u32(sRing + (size_t)X * slot_bytes));
    constexpr uint64_t kb_u = kb_bytes >> 4;
    const uint32_t tmemD = tmem + colD + (uint32_t)(2 * CHB * X);
    for (uint32_t n = 0;; ++n) {
      tc::mbar_wait(&full_bar[X], n & 1);
      if (*reinterpret_cast<volatile uint32_t*>(&quit[X]) != 0) break;
      if (lane == 0) *reinterpret_cast<volatile uint32_t*>(&tiles_landed[X]) = n + 1;
      tc::tc_fence_after();
      if constexpr (STAMP) { if (lane == 0 && n < (uint32_t)T) RS_STAMP3(n, 2, gtime()); __syncwarp(); }   // tile landed (batch n = step n without retries)
      issue_ts_blocks<12>(nkb_t, tmemD, tmem, dB0, kb_u, idesc, 0u);                                       // 12: H = 768
      for (int kb = nkb_t; kb < nkb; ++kb)
        tc::mma4_bf16_ss_warp(tmemD, dA0 + (uint64_t)(kb - nkb_t) * (16384 >> 4), dB0 + (uint64_t)kb * kb_u, idesc, (uint32_t)(kb != 0));
      tc::mma_commit_warp(&tfull_bar[X]);
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps of chain X = warp >> 2
    const int q = warp & 3, X = warp >> 2, l16 = lane & 15, up = lane >> 4;
    const int m = q * 16 + l16;                 // gate row (hi copy in lanes 0..15, lo copy in 16..31)
    const int g = lane & 3;                     // gate of this row: 0 i, 1 j, 2 f, 3 o
    const int ul = m >> 2;                      // unit inside the slice
    const int unit = j * TSU + ul;
    const int ctid = (int)threadIdx.x & 127;    // thread inside the chain
    const int rowsX = min(max(B - CHB * X, 0), CHB);
    const uint32_t tmemD = tmem + colD + (uint32_t)(2 * CHB * X) + ((uint32_t)(q * 32) << 16);
    const bool leader = ctid == 0;
    const CUtensorMap* tmD = X ? &tmD1 : &tmD0;
    unsigned* ctr = a.barrier + 32 * X;
    unsigned char* hbase = reinterpret_cast<unsigned char*>(a.h_hi);
    const size_t row_bytes = (size_t)4 * H;                          // [hi H | lo H] bf16
    // publish side: thread ctid < 64 stores the 16-byte half `phalf` of (row pbl, plane ppl) of the staged tile
    const int pbl = ctid >> 2, ppl = (ctid >> 1) & 1, phalf = ctid & 1;
    const bool publisher = ctid < 64 && pbl < rowsX;
    const int ncol = min(max(rowsX - 8 * up, 0), 8);                 // batch columns of this lane that exist
    float c[2], hl[2];
    int lenr[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int b = X * CHB + 8 * up + 2 * g + k;
      const bool ok = b < B;
      c[k] = ok ? a.c0[(size_t)b * H + unit] : 0.f;
      hl[k] = ok ? a.h0[(size_t)b * H + unit] : 0.f;
      lenr[k] = ok ? a.len[b] : 0;
    }
    const float fbias = (g == 2) ? 1.0f : 0.0f;          // forget_bias
    const float pre = (g == 1) ? 2.0f : 1.0f;            // tanh(x) = 2*sigmoid(2x) - 1
    const float post_m = (g == 1) ? 2.0f : 1.0f, post_a = (g == 1) ? -1.0f : 0.0f;
    float2* blob = reinterpret_cast<float2*>(a.gates);
    uint32_t n = 0;                                      // MMA batches consumed

    for (int ti = 0; ti < T; ++ti) {
      const int t = a.t0 + ti;
      // hoisted input projection for this step (independent of the recurrence: issue early); column group X of Bpad = 16 or 32
      const float4* gp = reinterpret_cast<const float4*>(a.gx + (((size_t)t * nslice + j) * 64 + m) * p.Bpad + X * 16 + 8 * up);
      const float4 gx0 = __ldg(gp), gx1 = __ldg(gp + 1);
      float ig[2], jg[2], fg[2], og[2], c_new[2], h_new[2], hout[2];
      for (;;) {
        tc::mbar_wait(&tfull_bar[X], n & 1);
        ++n;
        tc::tc_fence_after();
        if (lane == 0) *reinterpret_cast<volatile uint32_t*>(&steps_done[X]) = (uint32_t)(ti + 1);     // (every warp: whichever wakes first)
        if (leader) RS_STAMP3(ti, 3, gtime());
        float v[16], w[16];
        tc::tmem_ld16(tmemD, v);                               // hi rows: W_hi h_hi ; lo rows: W_lo h_hi
        tc::tmem_ld16(tmemD + (uint32_t)CHB, w);               // hi rows: W_hi h_lo
        tc::tmem_ld_wait();
        float act[8];
        const float xs[8] = {gx0.x, gx0.y, gx0.z, gx0.w, gx1.x, gx1.y, gx1.z, gx1.w};
        bool nan = false;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float s_a = v[i] + w[i], s_b = v[8 + i] + w[8 + i];
          const float got = __shfl_xor_sync(0xffffffffu, up ? v[i] : s_b, 16);
          const float z = ((up ? (got + v[8 + i]) : (s_a + got)) + xs[i] + fbias) * pre;
          nan = nan || (i < ncol && z != z);                   // fill pattern anywhere in tile row 8 up + i (either plane)
          act[i] = fmaf(fast_sigmoid(z), post_m, post_a);
        }
        float own[2], rcv[3][2];
#pragma unroll
        for (int k = 0; k < 2; ++k) own[k] = pick4(g, act[k], act[2 + k], act[4 + k], act[6 + k]);
#pragma unroll
        for (int d = 1; d < 4; ++d)
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const float snd = pick4(g ^ d, act[k], act[2 + k], act[4 + k], act[6 + k]);
            rcv[d - 1][k] = __shfl_xor_sync(0xffffffffu, snd, d);
          }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          ig[k] = pick4(g ^ 0, own[k], rcv[0][k], rcv[1][k], rcv[2][k]);
          jg[k] = pick4(g ^ 1, own[k], rcv[0][k], rcv[1][k], rcv[2][k]);
          fg[k] = pick4(g ^ 2, own[k], rcv[0][k], rcv[1][k], rcv[2][k]);
          og[k] = pick4(g ^ 3, own[k], rcv[0][k], rcv[1][k], rcv[2][k]);
          c_new[k] = c[k] * fg[k] + ig[k] * jg[k];
          h_new[k] = fast_tanh(c_new[k]) * og[k];
          const int bl = 8 * up + 2 * g + k;                   // row inside the chain
          hout[k] = (t < lenr[k]) ? h_new[k] : 0.f;
          __nv_bfloat16 hh, hlo;
          tc::split_bf16(hout[k], hh, hlo);
          sH[X][bl][0][ul] = hh;                               // staged, not yet published
          sH[X][bl][1][ul] = hlo;
        }
        tc::tc_fence_before();
        if (leader) RS_STAMP3(ti, 4, gtime());
        // the barrier before the publish doubles as the vote on the tile (NaN = fill pattern, or the model's own NaN)
        if (!bar_any(1 + X, 128, nan)) break;
        const bool stale = bar_any(1 + X, 128, ti > 0 && tile_has_fill(tc::smem_u32(sRing + (size_t)X * slot_bytes), slot_bytes, CHB, rowsX, ctid, 128));
        if (!stale) break;
        if (leader) *reinterpret_cast<volatile uint32_t*>(&retry_req[X]) = *reinterpret_cast<volatile uint32_t*>(&retry_req[X]) + 1u;
      }
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (t < lenr[k]) { c[k] = c_new[k]; hl[k] = h_new[k]; }
      // publish h_t: slot t + 1, this CTA's 16 units of every row of the chain -- one 32-byte sector per (row, plane) --
      // and the hint
      if (q < 2) {
        void* dst = hbase + (size_t)((t + 1) * B + CHB * X + pbl) * row_bytes + (size_t)ppl * (2 * H) + (size_t)(j * TSU + phalf * 8) * 2;
        const uint4 hv = *reinterpret_cast<const uint4*>(&sH[X][pbl][ppl][phalf * 8]);
        bool faulty = false;
        if constexpr (STAMP) faulty = p.fault != 0;
        if (!faulty) {
          if (publisher) st_relaxed_v4(dst, hv);
          __syncwarp();
          if (lane == 0 && ti + 1 < T) red_relaxed_add(ctr, 1u);
        } else {
          // test of the retry path: the hint overtakes half of the data by 3 us
          if (publisher && !phalf) st_relaxed_v4(dst, hv);
          __syncwarp();
          if (lane == 0 && ti + 1 < T) red_relaxed_add(ctr, 1u);
          const unsigned long long until = gtime() + 3000ull;
          while (gtime() < until) {}
          __syncwarp();
          if (publisher && phalf) st_relaxed_v4(dst, hv);
        }
      }
      if (leader) { RS_STAMP3(ti, 5, gtime()); RS_STAMP3(ti, 7, (unsigned long long)n); }
      // reserve for backward (not on the critical path of the recurrence)
      if (blob) {
        blob[blob_idx(t, nslice, j, 1, 0, 0, warp, lane)] = make_float2(ig[0], ig[1]);
        blob[blob_idx(t, nslice, j, 1, 0, 1, warp, lane)] = make_float2(jg[0], jg[1]);
        blob[blob_idx(t, nslice, j, 1, 0, 2, warp, lane)] = make_float2(fg[0], fg[1]);
        blob[blob_idx(t, nslice, j, 1, 0, 3, warp, lane)] = make_float2(og[0], og[1]);
        blob[blob_idx(t, nslice, j, 1, 0, 4, warp, lane)] = make_float2(c_new[0], c_new[1]);
      }
      // fused hop: dropout(out_t) planes for the next layer / the output dense
      if (a.drop_hi) {
        if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");       // the previous hop store has read sD
        chain_bar_sync(X);                                     // (also: every publisher has read sH)
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int bl = 8 * up + 2 * g + k;
          float dv = hout[k];
          const unsigned long long idx = ((unsigned long long)t * B + (X * CHB + bl)) * H + unit;
          if (a.drop_thr_a != 0xffffffffu) dv = dropout_keep(a.drop_key, a.drop_sa, idx, a.drop_thr_a) ? dv * a.drop_inv_a : 0.f;
          if (a.drop_thr_b != 0xffffffffu) dv = dropout_keep(a.drop_key, a.drop_sb, idx, a.drop_thr_b) ? dv * a.drop_inv_b : 0.f;
          __nv_bfloat16 dh_, dl_;
          tc::split_bf16(dv, dh_, dl_);
          sD[X][bl][0][ul] = dh_;
          sD[X][bl][1][ul] = dl_;
        }
        tc::fence_proxy_async_smem();
        chain_bar_sync(X);
        if (leader) {
          tc::tma_store_3d(tmD, &sD[X][0][0][0], j * TSU, 0, t * B + CHB * X);
          tc::bulk_commit();
        }
      }
      // (sH is staged again only after the next fetch, i.e. after the grid-wide count that includes this CTA's two adds,
      //  each of which follows its warp's reads of sH -- except with the injected fault, where the add comes first)
      if (p.endbar) chain_bar_sync(X);
      if (leader) RS_STAMP3(ti, 6, gtime());
    }
    if (leader) {
      // release the chain's producer and MMA warps
      *reinterpret_cast<volatile uint32_t*>(&quit[X]) = 1u;
      tc::mbar_arrive(&full_bar[X]);
      if (a.drop_hi) tc::bulk_wait_all();
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int b = X * CHB + 8 * up + 2 * g + k;
      if (b < B) {
        if (a.cT) a.cT[(size_t)b * H + unit] = c[k];
        if (a.hT) a.hT[(size_t)b * H + unit] = hl[k];
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------
// Backward
// ------------------------------------------------------------------------------------
struct KBwd {
  RecTcBwdArgs a;
  int H, B, Bpad, nslice, nkbs, ngl;       // nkbs = K-blocks of this CTA's K segment (H/2 / 64)
  int x3;                                  // 1: bf16x3 product (Wh_lo resident, dgates lo plane streamed); 0: one product
  int nlo_t;                               // K-blocks of Wh_lo resident in tensor memory (the other nkbs - nlo_t in smem)
  int variant;
  uint32_t tmem_cols;
  int fault;                               // STAMP instantiations of rec_ts_bwd3/4_kernel: delay half of every publish behind its hint
  int turns;                               // rec_ts_bwd4_kernel: the chains take turns at the TMA port
  int bulk;                                // rec_ts_bwd4_kernel: the reduce-scatter pushes 1 KB bulk copies (one per half-warp and owner)
                                           // out of a staging buffer instead of 16-byte st.async (RS_TS_PUSH_BULK)
};

// 128 registers (no spills; 168 unconstrained) x 320 threads and ~50 KB of shared memory: with RS_TC_CORES=1 a 256-thread
// GEMM CTA (gemm_tc_kernel<128, 2>) fits on the same SM
// X3: 1 / 0 = the product's term count fixed at compile time (1 with HH > 0: every Wh_lo block in tensor memory), -1 = from the arguments
template <bool STAMP, int VARIANT, int BPAD, int HH, int X3>
__global__ void __maxnreg__(128)
rec_ts_bwd_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmS_hi,
                  const __grid_constant__ CUtensorMap tmS_lo, KBwd p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar, tfull_bar, recv_bar;
  __shared__ uint32_t tmem_slot;
  const RecTcBwdArgs& a = p.a;
  const int variant = VARIANT >= 0 ? VARIANT : p.variant;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = HH > 0 ? HH : p.H, B = p.B, Bpad = BPAD > 0 ? BPAD : p.Bpad, T = a.T, nkbs = HH > 0 ? HH / 128 : p.nkbs, G = 4 * H;
  const int nslice = HH > 0 ? HH / TSU : p.nslice;
  const uint32_t rank = cluster_ctarank();            // K segment / owned unit slice inside the 128-unit block
  const int blk = blockIdx.x / CL;
  const int j = blk * CL + (int)rank;                   // 16-unit slice (same numbering as forward)
  const int kseg0 = (int)rank * (H / 2);                // first dgates column of this CTA's K segment
  const bool x3 = X3 >= 0 ? (X3 != 0) : (p.x3 != 0);
  const int nlo_t = !x3 ? 0 : ((X3 == 1 && HH > 0) ? HH / 128 : p.nlo_t);
  const int nlo_s = x3 ? nkbs - nlo_t : 0;
  const uint32_t kb_bytes = (uint32_t)Bpad * 128 * (x3 ? 2u : 1u);          // one streamed K-block: hi rows (; lo rows)
  unsigned char* sG = smem;                                                  // [nkbs][planes][Bpad x 128 B] dgates_t K segment
  unsigned char* sAlo = smem + (size_t)nkbs * kb_bytes;                      // [nlo_s][128 rows x 128 B] Wh_lo blocks outside TMEM
  float* sR = reinterpret_cast<float*>(sAlo + (size_t)nlo_s * 16384);        // [CL src][16 units][Bpad] partial dh
  __nv_bfloat16* sDG = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<unsigned char*>(sR) + (size_t)CL * TSU * Bpad * 4);
                                                                             // [2 planes][Bpad][4 gates][16 units]
  const uint32_t colLo = (uint32_t)(H / 4);              // Wh_hi: columns [0, H/4); Wh_lo: nlo_t blocks of 32 columns behind
  const uint32_t colD = colLo + (uint32_t)nlo_t * 32;
  const int t1 = a.t0 + T;                                // this launch: steps t1 - 1 down to t0
  const bool primed = t1 < a.Ttot;                        // dh_{t1-1} comes from dgates_{t1} of the previous launch
  const int ts_first = primed ? t1 : t1 - 1;              // first dgates step streamed through the tensor core

  if (threadIdx.x == 0) {
    tc::mbar_init(&full_bar, 1);
    tc::mbar_init(&tfull_bar, 1);
    tc::mbar_init(&recv_bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 8) tc::tmem_alloc(&tmem_slot, p.tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp < 4) {
    // resident Wh[128 rows of the block][K segment]: TMEM lane = threadIdx.x = hidden unit k
    const __nv_bfloat16* src = a.wh_hi + (size_t)(blk * 128 + (int)threadIdx.x) * G + kseg0;
    for (int c0 = 0; c0 < H / 4; c0 += 8) {
      const uint4 x0 = __ldg(reinterpret_cast<const uint4*>(src + 2 * c0));
      const uint4 x1 = __ldg(reinterpret_cast<const uint4*>(src + 2 * c0 + 8));
      const uint32_t r[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      tc::tmem_st8(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0, r);
    }
    if (x3) {
      const __nv_bfloat16* srcl = a.wh_lo + (size_t)(blk * 128 + (int)threadIdx.x) * G + kseg0;
      for (int c0 = 0; c0 < nlo_t * 32; c0 += 8) {
        const uint4 x0 = __ldg(reinterpret_cast<const uint4*>(srcl + 2 * c0));
        const uint4 x1 = __ldg(reinterpret_cast<const uint4*>(srcl + 2 * c0 + 8));
        const uint32_t r[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        tc::tmem_st8(tmem + ((uint32_t)(32 * warp) << 16) + colLo + (uint32_t)c0, r);
      }
      for (int kb = nlo_t; kb < nkbs; ++kb) {
        unsigned char* tile = sAlo + (size_t)(kb - nlo_t) * 16384 + (size_t)threadIdx.x * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(tile + ((c ^ (threadIdx.x & 7)) << 4)) = __ldg(reinterpret_cast<const uint4*>(srcl + kb * 64 + c * 8));
      }
      if (nlo_s > 0) tc::fence_proxy_async_smem();
    }
    tc::tmem_st_wait();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  cluster_sync_all();                                   // peers' mbarriers are initialised before any st.async

  if (warp == 8) {
    if (lane == 0) tc::tma_prefetch_desc(&tmG);
    __syncwarp();
    const unsigned per_step = gridDim.x * 8u;
    unsigned epoch = 0;
    for (int t = ts_first; t >= a.t0 + 1; --t) {        // dh_{t-1} from dgates_t
      if (t < t1) {                                       // dgates_t comes from this launch: grid barrier
        ++epoch;
        wait_counter(a.barrier, per_step * epoch, variant);
        __syncwarp();
        if (!(variant & 8)) tc::fence_proxy_async_all();
      }
      if (lane == 0) RS_STAMP(a.dbg, t, 0);
      __syncwarp();
      tc::mbar_arrive_expect_tx_warp(&full_bar, (uint32_t)nkbs * kb_bytes);
      if (x3) tc::tma_load_4d_warp(sG, &tmG, 0, t * B, 0, kseg0 / 64, &full_bar);       // one box: [kb][plane][Bpad][64]
      else for (int i = 0; i < nkbs; ++i) tc::tma_load_2d_warp(sG + (size_t)i * kb_bytes, &tmG, kseg0 + i * 64, t * B, &full_bar);
      if (lane == 0) RS_STAMP(a.dbg, t, 1);
      __syncwarp();
    }
  } else if (warp == 9) {
    const uint32_t idesc = tc::instr_desc_bf16(128, Bpad);               // N = one plane
    const uint32_t idesc2 = tc::instr_desc_bf16(128, 2 * Bpad);          // N = both planes stacked
    const uint64_t dg0 = tc::smem_desc_sw128(tc::smem_u32(sG));
    const uint64_t dAlo0 = tc::smem_desc_sw128(tc::smem_u32(sAlo));
    const uint32_t tmemD = tmem + colD;
    uint32_t n = 0;
    for (int t = ts_first; t >= a.t0 + 1; --t, ++n) {
      tc::mbar_wait(&full_bar, n & 1);
      tc::tc_fence_after();
      if (lane == 0) RS_STAMP(a.dbg, t, 2);
      __syncwarp();
      if (x3) {
        // D[:, 0:2Bpad) = Wh_hi [dg_hi | dg_lo] ; D[:, 0:Bpad) += Wh_lo dg_hi
        issue_ts_blocks<6>(nkbs, tmemD, tmem, dg0, (uint64_t)(kb_bytes >> 4), idesc2, 0u);                             // 6: H = 768
        issue_ts_blocks<6>(nlo_t, tmemD, tmem + colLo, dg0, (uint64_t)(kb_bytes >> 4), idesc, 1u);
        for (int i = nlo_t; i < nkbs; ++i)
          tc::mma4_bf16_ss_warp(tmemD, dAlo0 + (uint64_t)(i - nlo_t) * (16384 >> 4), dg0 + (uint64_t)i * (kb_bytes >> 4), idesc, 1u);
      } else
      issue_ts_blocks<6>(nkbs, tmemD, tmem, dg0, (uint64_t)(kb_bytes >> 4), idesc, 0u);
      tc::mma_commit_warp(&tfull_bar);
      if (lane == 0) RS_STAMP(a.dbg, t, 3);
      __syncwarp();
    }
  } else {
    const int q = warp & 3, hf = warp >> 2, l16 = lane & 15, up = lane >> 4;
    const int g = lane & 3;
    const int ul = q * 4 + (l16 >> 2);                   // owned unit inside the slice (cell math)
    const int unit = j * TSU + ul;
    const int ng = Bpad / 16;
    const uint32_t tmemD = tmem + colD + ((uint32_t)(q * 32) << 16);
    // push side: TMEM lane 32q + lane = unit 32q + lane of the block -> owner rank, unit inside its slice
    const uint32_t owner = (uint32_t)(2 * q + up);
    const uint32_t sR_local = tc::smem_u32(sR);
    const uint32_t push_base = mapa_u32(sR_local + (uint32_t)(((int)rank * TSU + l16) * Bpad) * 4u, owner);
    const uint32_t push_bar = mapa_u32(tc::smem_u32(&recv_bar), owner);
    const float2* blob = reinterpret_cast<const float2*>(a.gates);
    int lenr[MAXG][2];
    float dc[MAXG][2];
#pragma unroll
    for (int gl = 0; gl < MAXG; ++gl)
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int b = (hf + 2 * gl) * 16 + 8 * up + 2 * g + k;
        lenr[gl][k] = ((hf + 2 * gl) < ng && b < B) ? a.len[b] : 0;
        dc[gl][k] = 0.f;
      }
    float2* carry = reinterpret_cast<float2*>(a.dc_carry);
    if (primed && carry) {
#pragma unroll
      for (int gl = 0; gl < MAXG; ++gl) {
        const float2 x = carry[(((size_t)j * MAXG + gl) * 8 + warp) * 32 + lane];
        dc[gl][0] = x.x; dc[gl][1] = x.y;
      }
    }
    const unsigned nchunk = (unsigned)(2 * Bpad * 4 * 2);     // 16-byte chunks of the staged dgates tile
    uint32_t n = primed ? 1u : 0u;                            // n - 1 = index of the MMA batch this step consumes
    for (int t = t1 - 1; t >= a.t0; --t, ++n) {
      // operands of the cell backward (independent of the recurrence: issue before waiting)
      float2 it2[MAXG][BLOB_ITEMS], cp2[MAXG];
      float dy[MAXG][2];
#pragma unroll
      for (int gl = 0; gl < MAXG; ++gl) {
        const int gi = hf + 2 * gl;
        if (gi < ng) {
#pragma unroll
          for (int it = 0; it < BLOB_ITEMS; ++it) it2[gl][it] = __ldg(blob + blob_idx(t, nslice, j, p.ngl, gl, it, warp, lane));
          if (t > 0) cp2[gl] = __ldg(blob + blob_idx(t - 1, nslice, j, p.ngl, gl, 4, warp, lane));
          float cpv[2];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int b = gi * 16 + 8 * up + 2 * g + k;
            float d = (b < B) ? __ldg(a.dout + ((size_t)t * B + b) * H + unit) : 0.f;
            // backward of the hop's dropout(s): the same masks, applied as the gradient is read
            const unsigned long long idx = ((unsigned long long)t * B + b) * H + unit;
            if (a.drop_thr_a != 0xffffffffu) d = dropout_keep(a.drop_key, a.drop_sa, idx, a.drop_thr_a) ? d * a.drop_inv_a : 0.f;
            if (a.drop_thr_b != 0xffffffffu) d = dropout_keep(a.drop_key, a.drop_sb, idx, a.drop_thr_b) ? d * a.drop_inv_b : 0.f;
            dy[gl][k] = d;
            cpv[k] = (t == 0 && b < B) ? __ldg(a.c0 + (size_t)b * H + unit) : 0.f;
          }
          if (t == 0) cp2[gl] = make_float2(cpv[0], cpv[1]);
        }
      }
      float dh[MAXG][2];
#pragma unroll
      for (int gl = 0; gl < MAXG; ++gl) dh[gl][0] = dh[gl][1] = 0.f;
      if (n > 0) {
        // partial dh_t^T of this CTA's K segment -> reduce-scatter over the cluster
        if (threadIdx.x == 0) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                       ::"r"(tc::smem_u32(&recv_bar)), "r"((uint32_t)(CL * TSU * Bpad * 4)) : "memory");
        }
        tc::mbar_wait(&tfull_bar, (n - 1) & 1);
        tc::tc_fence_after();
        if (threadIdx.x == 0) RS_STAMP(a.dbg, t, 4);
#pragma unroll
        for (int gl = 0; gl < MAXG; ++gl) {
          const int gi = hf + 2 * gl;
          if (gi < ng) {
            float v[16];
            tc::tmem_ld16(tmemD + (uint32_t)(gi * 16), v);
            if (x3) {
              float w[16];
              tc::tmem_ld16(tmemD + (uint32_t)(Bpad + gi * 16), w);      // Wh_hi dg_lo
              tc::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += w[i];
            } else {
              tc::tmem_ld_wait();
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
              st_async_v4(push_base + (uint32_t)(gi * 16 + 4 * i) * 4u, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]),
                          push_bar);
          }
        }
        tc::tc_fence_before();
        if (threadIdx.x == 0) RS_STAMP(a.dbg, t, 9);
        tc::mbar_wait(&recv_bar, (n - 1) & 1);
        if (threadIdx.x == 0) RS_STAMP(a.dbg, t, 8);
#pragma unroll
        for (int gl = 0; gl < MAXG; ++gl) {
          const int gi = hf + 2 * gl;
          if (gi < ng) {
#pragma unroll
            for (int s = 0; s < CL; ++s) {
              const float2 x = *reinterpret_cast<const float2*>(sR + ((size_t)s * TSU + ul) * Bpad + gi * 16 + 8 * up + 2 * g);
              dh[gl][0] += x.x; dh[gl][1] += x.y;
            }
          }
        }
      }
      if (threadIdx.x == 0) RS_STAMP(a.dbg, t, 13);
#pragma unroll
      for (int gl = 0; gl < MAXG; ++gl) {
        const int gi = hf + 2 * gl;
        if (gi < ng) {
          const float* pi = reinterpret_cast<const float*>(&it2[gl][0]);
          const float* pj = reinterpret_cast<const float*>(&it2[gl][1]);
          const float* pf = reinterpret_cast<const float*>(&it2[gl][2]);
          const float* po = reinterpret_cast<const float*>(&it2[gl][3]);
          const float* pct = reinterpret_cast<const float*>(&it2[gl][4]);
          const float* pcp = reinterpret_cast<const float*>(&cp2[gl]);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int b = gi * 16 + 8 * up + 2 * g + k;
            float di = 0.f, dj = 0.f, df = 0.f, dob = 0.f;
            if (t < lenr[gl][k]) {
              const float ig = pi[k], jg = pj[k], fg = pf[k], og = po[k];
              const float dh_tot = dh[gl][k] + dy[gl][k];
              const float tch = fast_tanh(pct[k]);
              dob = dh_tot * tch * og * (1.f - og);
              const float dc_tot = dc[gl][k] + dh_tot * og * (1.f - tch * tch);
              di = dc_tot * jg * ig * (1.f - ig);
              dj = dc_tot * ig * (1.f - jg * jg);
              df = dc_tot * pcp[k] * fg * (1.f - fg);
              dc[gl][k] = dc_tot * fg;
            }
            const float d4[4] = {di, dj, df, dob};
#pragma unroll
            for (int gg = 0; gg < 4; ++gg) {
              __nv_bfloat16 h0, l0;
              tc::split_bf16(d4[gg], h0, l0);
              sDG[(((size_t)0 * Bpad + b) * 4 + gg) * TSU + ul] = h0;
              sDG[(((size_t)1 * Bpad + b) * 4 + gg) * TSU + ul] = l0;
            }
          }
        }
      }
      if (threadIdx.x == 0) RS_STAMP(a.dbg, t, 5);
      if (variant & 16) tc::fence_proxy_async_smem();
      epi_bar_sync();
      if (threadIdx.x == 0) RS_STAMP(a.dbg, t, 10);
      if (variant & 16) {
        if (threadIdx.x == 0) {
          tc::tma_store_3d(&tmS_hi, sDG, j * TSU, 0, t * B);
          tc::tma_store_3d(&tmS_lo, sDG + (size_t)Bpad * 4 * TSU, j * TSU, 0, t * B);
          tc::bulk_commit();
          RS_STAMP(a.dbg, t, 11);
          tc::bulk_wait_all();
          RS_STAMP(a.dbg, t, 12);
          if (t > a.t0) {
            if ((variant & 32) && !(variant & 64)) { if (variant & 256) tc::fence_proxy_async_global(); else tc::fence_proxy_async_all(); }
            if ((variant & 32) && !(variant & 128)) red_release_add(a.barrier, 8u);
            else red_relaxed_add(a.barrier, 8u);
          }
          RS_STAMP(a.dbg, t, 6);
        }
        continue;
      }
      // publish dgates_t: 16-byte coalesced stores of the staged tile, one release per warp
      for (unsigned ch = threadIdx.x; ch < nchunk; ch += 256) {
        const int half = ch & 1, gg = (ch >> 1) & 3, b = (ch >> 3) % Bpad, pl = (ch >> 3) / Bpad;
        if (b < B) {
          const uint4 x = *reinterpret_cast<const uint4*>(sDG + (((size_t)pl * Bpad + b) * 4 + gg) * TSU + half * 8);
          __nv_bfloat16* dst = (pl ? a.dg_lo : a.dg_hi) + ((size_t)t * B + b) * G + (size_t)gg * H + j * TSU + half * 8;
          *reinterpret_cast<uint4*>(dst) = x;
        }
      }
      if (threadIdx.x == 0) RS_STAMP(a.dbg, t, 11);
      if (!(variant & 1)) tc::fence_proxy_async_all();
      __syncwarp();
      if (threadIdx.x == 0) RS_STAMP(a.dbg, t, 12);
      if (lane == 0 && t > a.t0) red_release_add(a.barrier, 1u);
      if (threadIdx.x == 0) RS_STAMP(a.dbg, t, 6);
      // (the staged tile is rewritten only after the next tfull wait, i.e. after the grid barrier
      //  that needs all eight releases above, each of which follows its warp's reads of the tile)
    }
    if (carry) {
#pragma unroll
      for (int gl = 0; gl < MAXG; ++gl)
        carry[(((size_t)j * MAXG + gl) * 8 + warp) * 32 + lane] = make_float2(dc[gl][0], dc[gl][1]);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 8) tc::tmem_dealloc(tmem, p.tmem_cols);
}

// ------------------------------------------------------------------------------------
// Backward with the validated exchange of rec_ts_fwd3_kernel (bf16x3 only).  The dgates planes of the launch's steps
// carry the fill pattern until their producers overwrite it; a CTA publishes its 4 x 16 columns of dgates_t with 16-byte
// st.relaxed.gpu stores and a relaxed `red` per warp (the hint), the producer warp fetches the K segment with one TMA box
// when the counter says so.  A tile row (one utterance's dgates) that still holds fill pattern makes that batch column of
// the partial dh NaN in all 128 unit rows; unlike forward the result of the MMAs leaves the CTA before the cell math (the
// reduce-scatter through distributed shared memory), so the epilogue warps vote on their accumulator columns in a barrier of
// their own before they push; a NaN vote takes the slow path of rec_ts_fwd3_kernel (scan for the fill pattern, fetch and
// multiply again if it is there).
// ------------------------------------------------------------------------------------
#define RS_STAMPB(step, ev, val) do { if constexpr (STAMP) if (a.dbg && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) \
    a.dbg[((size_t)(blockIdx.x ? a.T : 0) + (size_t)((step) - a.t0)) * 16 + (ev)] = (val); } while (0)
template <bool STAMP, int BPAD, int HH>
__global__ void __maxnreg__(128)
rec_ts_bwd3_kernel(const __grid_constant__ CUtensorMap tmG, KBwd p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar, tfull_bar, recv_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ uint32_t retry_req, quit;                   // fetches asked for again (running count); the last step is done
  const RecTcBwdArgs& a = p.a;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = HH > 0 ? HH : p.H, B = p.B, Bpad = BPAD > 0 ? BPAD : p.Bpad, T = a.T, nkbs = HH > 0 ? HH / 128 : p.nkbs, G = 4 * H;
  const int nslice = HH > 0 ? HH / TSU : p.nslice;
  const uint32_t rank = cluster_ctarank();            // K segment / owned unit slice inside the 128-unit block
  const int blk = blockIdx.x / CL;
  const int j = blk * CL + (int)rank;                   // 16-unit slice (same numbering as forward)
  const int kseg0 = (int)rank * (H / 2);                // first dgates column of this CTA's K segment
  const int nlo_t = HH > 0 ? HH / 128 : p.nlo_t;
  const int nlo_s = nkbs - nlo_t;
  const uint32_t kb_bytes = (uint32_t)Bpad * 128 * 2u;                       // one streamed K-block: hi rows ; lo rows
  unsigned char* sG = smem;                                                  // [nkbs][2 planes][Bpad x 128 B] dgates_t K segment
  unsigned char* sAlo = smem + (size_t)nkbs * kb_bytes;                      // [nlo_s][128 rows x 128 B] Wh_lo blocks outside TMEM
  float* sR = reinterpret_cast<float*>(sAlo + (size_t)nlo_s * 16384);        // [CL src][16 units][Bpad] partial dh
  __nv_bfloat16* sDG = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<unsigned char*>(sR) + (size_t)CL * TSU * Bpad * 4);
                                                                             // [2 planes][Bpad][4 gates][16 units]
  const uint32_t colLo = (uint32_t)(H / 4);              // Wh_hi: columns [0, H/4); Wh_lo: nlo_t blocks of 32 columns behind
  const uint32_t colD = colLo + (uint32_t)nlo_t * 32;
  const int t1 = a.t0 + T;                                // this launch: steps t1 - 1 down to t0
  const bool primed = t1 < a.Ttot;                        // dh_{t1-1} comes from dgates_{t1} of the previous launch
  const int ts_first = primed ? t1 : t1 - 1;              // first dgates step streamed through the tensor core

  if (threadIdx.x == 0) {
    tc::mbar_init(&full_bar, 1);
    tc::mbar_init(&tfull_bar, 1);
    tc::mbar_init(&recv_bar, 2 * CL);                    // two epilogue warps (column halves) of each of the CL sources
    retry_req = 0; quit = 0;
    tc::fence_mbar_init();
  }
  if (warp == 8) tc::tmem_alloc(&tmem_slot, p.tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp < 4) {
    // resident Wh[128 rows of the block][K segment]: TMEM lane = threadIdx.x = hidden unit k
    const __nv_bfloat16* src = a.wh_hi + (size_t)(blk * 128 + (int)threadIdx.x) * G + kseg0;
    for (int c0 = 0; c0 < H / 4; c0 += 8) {
      const uint4 x0 = __ldg(reinterpret_cast<const uint4*>(src + 2 * c0));
      const uint4 x1 = __ldg(reinterpret_cast<const uint4*>(src + 2 * c0 + 8));
      const uint32_t r[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      tc::tmem_st8(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0, r);
    }
    const __nv_bfloat16* srcl = a.wh_lo + (size_t)(blk * 128 + (int)threadIdx.x) * G + kseg0;
    for (int c0 = 0; c0 < nlo_t * 32; c0 += 8) {
      const uint4 x0 = __ldg(reinterpret_cast<const uint4*>(srcl + 2 * c0));
      const uint4 x1 = __ldg(reinterpret_cast<const uint4*>(srcl + 2 * c0 + 8));
      const uint32_t r[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      tc::tmem_st8(tmem + ((uint32_t)(32 * warp) << 16) + colLo + (uint32_t)c0, r);
    }
    for (int kb = nlo_t; kb < nkbs; ++kb) {
      unsigned char* tile = sAlo + (size_t)(kb - nlo_t) * 16384 + (size_t)threadIdx.x * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(tile + ((c ^ (threadIdx.x & 7)) << 4)) = __ldg(reinterpret_cast<const uint4*>(srcl + kb * 64 + c * 8));
    }
    if (nlo_s > 0) tc::fence_proxy_async_smem();
    tc::tmem_st_wait();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  cluster_sync_all();                                   // peers' mbarriers are initialised before any st.async

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) tc::tma_prefetch_desc(&tmG);
    __syncwarp();
    const unsigned per_step = gridDim.x * 8u;
    const uint32_t tile_bytes = (uint32_t)nkbs * kb_bytes;
    unsigned epoch = 0;
    uint32_t served = 0;                                 // retry requests answered
    auto fetch = [&](int t) {
      tc::mbar_arrive_expect_tx_warp(&full_bar, tile_bytes);
      tc::tma_load_4d_warp(sG, &tmG, 0, t * B, 0, kseg0 / 64, &full_bar);         // one box: [kb][plane][Bpad][64]
    };
    auto serve = [&](int t) {
      if (*reinterpret_cast<volatile uint32_t*>(&retry_req) != served) { ++served; __syncwarp(); fetch(t); }
    };
    for (int t = ts_first; t >= a.t0 + 1; --t) {        // dh_{t-1} from dgates_t
      if (t < t1) {                                       // dgates_t comes from this launch: wait for the hint
        ++epoch;
        uint32_t spins = 0;
        while (ld_relaxed_u32(a.barrier) < per_step * epoch) { serve(t + 1); if (++spins > (1u << 26)) __trap(); }
        __syncwarp();
      }
      if (lane == 0) RS_STAMPB(t, 0, gtime());
      __syncwarp();
      fetch(t);
      if (lane == 0) RS_STAMPB(t, 1, gtime());
      __syncwarp();
    }
    while (*reinterpret_cast<volatile uint32_t*>(&quit) == 0) serve(a.t0 + 1);
  } else if (warp == 9) {
    const uint32_t idesc = tc::instr_desc_bf16(128, Bpad);               // N = one plane
    const uint32_t idesc2 = tc::instr_desc_bf16(128, 2 * Bpad);          // N = both planes stacked
    const uint64_t dg0 = tc::smem_desc_sw128(tc::smem_u32(sG));
    const uint64_t dAlo0 = tc::smem_desc_sw128(tc::smem_u32(sAlo));
    const uint32_t tmemD = tmem + colD;
    for (uint32_t n = 0;; ++n) {                       // one batch of MMAs per fetch
      tc::mbar_wait(&full_bar, n & 1);
      if (*reinterpret_cast<volatile uint32_t*>(&quit) != 0) break;
      tc::tc_fence_after();
      // D[:, 0:2Bpad) = Wh_hi [dg_hi | dg_lo] ; D[:, 0:Bpad) += Wh_lo dg_hi
      issue_ts_blocks<6>(nkbs, tmemD, tmem, dg0, (uint64_t)(kb_bytes >> 4), idesc2, 0u);                             // 6: H = 768
      issue_ts_blocks<6>(nlo_t, tmemD, tmem + colLo, dg0, (uint64_t)(kb_bytes >> 4), idesc, 1u);
      for (int i = nlo_t; i < nkbs; ++i)
        tc::mma4_bf16_ss_warp(tmemD, dAlo0 + (uint64_t)(i - nlo_t) * (16384 >> 4), dg0 + (uint64_t)i * (kb_bytes >> 4), idesc, 1u);
      tc::mma_commit_warp(&tfull_bar);
    }
  } else {
    const int q = warp & 3, hf = warp >> 2, l16 = lane & 15, up = lane >> 4;
    const int g = lane & 3;
    const int ul = q * 4 + (l16 >> 2);                   // owned unit inside the slice (cell math)
    const int unit = j * TSU + ul;
    const int ng = Bpad / 16;
    const uint32_t tmemD = tmem + colD + ((uint32_t)(q * 32) << 16);
    // push side: TMEM lane 32q + lane = unit 32q + lane of the block -> owner rank, unit inside its slice
    const uint32_t owner = (uint32_t)(2 * q + up);
    const uint32_t sR_local = tc::smem_u32(sR);
    const uint32_t push_base = mapa_u32(sR_local + (uint32_t)(((int)rank * TSU + l16) * Bpad) * 4u, owner);
    const uint32_t push_bar = mapa_u32(tc::smem_u32(&recv_bar), owner);
    const float2* blob = reinterpret_cast<const float2*>(a.gates);
    int lenr[MAXG][2];
    float dc[MAXG][2];
#pragma unroll
    for (int gl = 0; gl < MAXG; ++gl)
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int b = (hf + 2 * gl) * 16 + 8 * up + 2 * g + k;
        lenr[gl][k] = ((hf + 2 * gl) < ng && b < B) ? a.len[b] : 0;
        dc[gl][k] = 0.f;
      }
    float2* carry = reinterpret_cast<float2*>(a.dc_carry);
    if (primed && carry) {
#pragma unroll
      for (int gl = 0; gl < MAXG; ++gl) {
        const float2 x = carry[(((size_t)j * MAXG + gl) * 8 + warp) * 32 + lane];
        dc[gl][0] = x.x; dc[gl][1] = x.y;
      }
    }
    const unsigned nchunk = (unsigned)(2 * Bpad * 4 * 2);     // 16-byte chunks of the staged dgates tile
    uint32_t n = primed ? 1u : 0u;                            // n - 1 = index of the MMA batch this step consumes
    uint32_t na = 0;                                          // MMA batches consumed
    for (int t = t1 - 1; t >= a.t0; --t, ++n) {
      // operands of the cell backward (independent of the recurrence: issue before waiting)
      float2 it2[MAXG][BLOB_ITEMS], cp2[MAXG];
      float dy[MAXG][2];
#pragma unroll
      for (int gl = 0; gl < MAXG; ++gl) {
        const int gi = hf + 2 * gl;
        if (gi < ng) {
#pragma unroll
          for (int it = 0; it < BLOB_ITEMS; ++it) it2[gl][it] = __ldg(blob + blob_idx(t, nslice, j, p.ngl, gl, it, warp, lane));
          if (t > 0) cp2[gl] = __ldg(blob + blob_idx(t - 1, nslice, j, p.ngl, gl, 4, warp, lane));
          float cpv[2];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int b = gi * 16 + 8 * up + 2 * g + k;
            float d = (b < B) ? __ldg(a.dout + ((size_t)t * B + b) * H + unit) : 0.f;
            // backward of the hop's dropout(s): the same masks, applied as the gradient is read
            const unsigned long long idx = ((unsigned long long)t * B + b) * H + unit;
            if (a.drop_thr_a != 0xffffffffu) d = dropout_keep(a.drop_key, a.drop_sa, idx, a.drop_thr_a) ? d * a.drop_inv_a : 0.f;
            if (a.drop_thr_b != 0xffffffffu) d = dropout_keep(a.drop_key, a.drop_sb, idx, a.drop_thr_b) ? d * a.drop_inv_b : 0.f;
            dy[gl][k] = d;
            cpv[k] = (t == 0 && b < B) ? __ldg(a.c0 + (size_t)b * H + unit) : 0.f;
          }
          if (t == 0) cp2[gl] = make_float2(cpv[0], cpv[1]);
        }
      }
      float dh[MAXG][2];
#pragma unroll
      for (int gl = 0; gl < MAXG; ++gl) dh[gl][0] = dh[gl][1] = 0.f;
      if (n > 0) {
        // partial dh_t^T of this CTA's K segment -> reduce-scatter over the cluster
        float v[MAXG][16];
        const bool fresh = t + 1 < t1;                          // the tile (dgates_{t+1}) was produced by this launch
        for (;;) {
          tc::mbar_wait(&tfull_bar, na & 1);
          ++na;
          tc::tc_fence_after();
          if (threadIdx.x == 0) RS_STAMPB(t, 3, gtime());
          bool nan = false;
#pragma unroll
          for (int gl = 0; gl < MAXG; ++gl) {
            const int gi = hf + 2 * gl;
            if (gi < ng) {
              float w[16];
              tc::tmem_ld16(tmemD + (uint32_t)(gi * 16), v[gl]);
              tc::tmem_ld16(tmemD + (uint32_t)(Bpad + gi * 16), w);      // Wh_hi dg_lo
              tc::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                v[gl][i] += w[i];
                nan = nan || (gi * 16 + i < B && v[gl][i] != v[gl][i]);   // fill pattern in tile row gi*16 + i (either plane)
              }
            }
          }
          tc::tc_fence_before();
          if (threadIdx.x == 0) RS_STAMPB(t, 9, gtime());
          // the accumulator may leave the CTA only when its tile is known to have been complete: vote
          if (!bar_any(1, 256, nan)) break;
          const bool stale = bar_any(1, 256, fresh && tile_has_fill(tc::smem_u32(sG), (uint32_t)nkbs * kb_bytes, Bpad, B, (int)threadIdx.x, 256));
          if (!stale) break;                                      // the model's own NaN
          if (threadIdx.x == 0) *reinterpret_cast<volatile uint32_t*>(&retry_req) = *reinterpret_cast<volatile uint32_t*>(&retry_req) + 1u;
        }
        if (threadIdx.x == 0) RS_STAMPB(t, 10, gtime());
#pragma unroll
        for (int gl = 0; gl < MAXG; ++gl) {
          const int gi = hf + 2 * gl;
          if (gi < ng) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              st_cluster_v4(push_base + (uint32_t)(gi * 16 + 4 * i) * 4u, make_float4(v[gl][4 * i], v[gl][4 * i + 1], v[gl][4 * i + 2], v[gl][4 * i + 3]));
          }
        }
        // one arrival per warp and owner instead of a transaction count per 16 bytes (1024 updates of one mbarrier per step:
        // measured 0.69 us from the vote to the last st.async, 0.36 more until the barrier completed)
        __syncwarp();
        if (l16 == 0) mbar_arrive_remote_release(push_bar);
        if (threadIdx.x == 0) RS_STAMPB(t, 4, gtime());
        mbar_wait_cluster(&recv_bar, (n - 1) & 1);
        if (threadIdx.x == 0) RS_STAMPB(t, 5, gtime());
#pragma unroll
        for (int gl = 0; gl < MAXG; ++gl) {
          const int gi = hf + 2 * gl;
          if (gi < ng) {
#pragma unroll
            for (int s = 0; s < CL; ++s) {
              const float2 x = *reinterpret_cast<const float2*>(sR + ((size_t)s * TSU + ul) * Bpad + gi * 16 + 8 * up + 2 * g);
              dh[gl][0] += x.x; dh[gl][1] += x.y;
            }
          }
        }
      }
      if (threadIdx.x == 0) RS_STAMPB(t, 11, gtime());
#pragma unroll
      for (int gl = 0; gl < MAXG; ++gl) {
        const int gi = hf + 2 * gl;
        if (gi < ng) {
          const float* pi = reinterpret_cast<const float*>(&it2[gl][0]);
          const float* pj = reinterpret_cast<const float*>(&it2[gl][1]);
          const float* pf = reinterpret_cast<const float*>(&it2[gl][2]);
          const float* po = reinterpret_cast<const float*>(&it2[gl][3]);
          const float* pct = reinterpret_cast<const float*>(&it2[gl][4]);
          const float* pcp = reinterpret_cast<const float*>(&cp2[gl]);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int b = gi * 16 + 8 * up + 2 * g + k;
            float di = 0.f, dj = 0.f, df = 0.f, dob = 0.f;
            if (t < lenr[gl][k]) {
              const float ig = pi[k], jg = pj[k], fg = pf[k], og = po[k];
              const float dh_tot = dh[gl][k] + dy[gl][k];
              const float tch = fast_tanh(pct[k]);
              dob = dh_tot * tch * og * (1.f - og);
              const float dc_tot = dc[gl][k] + dh_tot * og * (1.f - tch * tch);
              di = dc_tot * jg * ig * (1.f - ig);
              dj = dc_tot * ig * (1.f - jg * jg);
              df = dc_tot * pcp[k] * fg * (1.f - fg);
              dc[gl][k] = dc_tot * fg;
            }
            const float d4[4] = {di, dj, df, dob};
#pragma unroll
            for (int gg = 0; gg < 4; ++gg) {
              __nv_bfloat16 h0, l0;
              tc::split_bf16(d4[gg], h0, l0);
              sDG[(((size_t)0 * Bpad + b) * 4 + gg) * TSU + ul] = h0;
              sDG[(((size_t)1 * Bpad + b) * 4 + gg) * TSU + ul] = l0;
            }
          }
        }
      }
      if (threadIdx.x == 0) RS_STAMPB(t, 6, gtime());
      epi_bar_sync();
      if (threadIdx.x == 0) RS_STAMPB(t, 12, gtime());
      // publish dgates_t: 16-byte stores of the staged tile (one 32-byte sector per (row, gate, plane)), then the hint
      bool faulty = false;
      if constexpr (STAMP) faulty = p.fault != 0;
      if (!faulty) {
        for (unsigned ch = threadIdx.x; ch < nchunk; ch += 256) {
          const int half = ch & 1, gg = (ch >> 1) & 3, b = (ch >> 3) % Bpad, pl = (ch >> 3) / Bpad;
          if (b < B) {
            const uint4 x = *reinterpret_cast<const uint4*>(sDG + (((size_t)pl * Bpad + b) * 4 + gg) * TSU + half * 8);
            st_relaxed_v4((pl ? a.dg_lo : a.dg_hi) + ((size_t)t * B + b) * G + (size_t)gg * H + j * TSU + half * 8, x);
          }
        }
        __syncwarp();
        if (lane == 0 && t > a.t0) red_relaxed_add(a.barrier, 1u);
      } else {
        // test of the retry path: the hint overtakes the odd chunks by 3 us (Bpad <= 32: at most two chunks per thread)
        uint4 x[2];
        __nv_bfloat16* dst[2];
        int nx = 0;
        for (unsigned ch = threadIdx.x; ch < nchunk && nx < 2; ch += 256) {
          const int half = ch & 1, gg = (ch >> 1) & 3, b = (ch >> 3) % Bpad, pl = (ch >> 3) / Bpad;
          x[nx] = *reinterpret_cast<const uint4*>(sDG + (((size_t)pl * Bpad + b) * 4 + gg) * TSU + half * 8);
          dst[nx] = b < B ? (pl ? a.dg_lo : a.dg_hi) + ((size_t)t * B + b) * G + (size_t)gg * H + j * TSU + half * 8 : nullptr;
          ++nx;
        }
        const bool late = threadIdx.x & 1;
        if (!late) for (int i = 0; i < nx; ++i) if (dst[i]) st_relaxed_v4(dst[i], x[i]);
        __syncwarp();
        if (lane == 0 && t > a.t0) red_relaxed_add(a.barrier, 1u);
        const unsigned long long until = gtime() + 3000ull;
        while (gtime() < until) {}
        __syncwarp();
        if (late) for (int i = 0; i < nx; ++i) if (dst[i]) st_relaxed_v4(dst[i], x[i]);
      }
      if (threadIdx.x == 0) { RS_STAMPB(t, 7, gtime()); RS_STAMPB(t, 2, (unsigned long long)na); }
      // (the staged tile is rewritten only after the next tfull wait, i.e. after a fetch that follows the grid-wide
      //  count, which needs all eight adds above, each of which follows its warp's reads of the tile)
    }
    if (threadIdx.x == 0) {
      // release the producer and MMA warps
      *reinterpret_cast<volatile uint32_t*>(&quit) = 1u;
      tc::mbar_arrive(&full_bar);
    }
    if (carry) {
#pragma unroll
      for (int gl = 0; gl < MAXG; ++gl)
        carry[(((size_t)j * MAXG + gl) * 8 + warp) * 32 + lane] = make_float2(dc[gl][0], dc[gl][1]);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 8) tc::tmem_dealloc(tmem, p.tmem_cols);
}

// ------------------------------------------------------------------------------------
// Backward, two chains, validated exchange.  The reduce-scatter of the partial dh over the cluster moves 16 KB out of and
// into every CTA per step through distributed shared memory (~20 B/clk per SM: measured 0.7 us from the first st.async to
// the last, 0.35 more until the transaction count is complete), the cell backward and the publish of dgates_t take another
// 1.4 us, and during all of it the TMA port and the tensor pipe of the SM are idle.  As in forward (rec_ts_fwd3_kernel) the
// mini-batch therefore runs as two independent 16-row recurrences per CTA -- own producer / MMA / four epilogue warps, own
// tile, accumulator columns, receive buffer, mbarriers, named barrier and grid counter -- that take turns at the TMA port,
// so that one chain fetches and multiplies while the other one reduces, computes its cells and publishes.
// Exchange and validation as in rec_ts_bwd3_kernel (fill pattern -> NaN accumulator columns -> vote before the push).
// Requires Bpad == 32, B > 16, bf16x3.
// ------------------------------------------------------------------------------------
template <bool STAMP, int HH>
__global__ void __maxnreg__(128)
rec_ts_bwd4_kernel(const __grid_constant__ CUtensorMap tmG, KBwd p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[2], tfull_bar[2], recv_bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ uint32_t retry_req[2], quit[2], tiles_landed[2];
  const RecTcBwdArgs& a = p.a;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = HH > 0 ? HH : p.H, B = p.B, T = a.T, nkbs = HH > 0 ? HH / 128 : p.nkbs, G = 4 * H;
  const int nslice = HH > 0 ? HH / TSU : p.nslice;
  const uint32_t rank = cluster_ctarank();            // K segment / owned unit slice inside the 128-unit block
  const int blk = blockIdx.x / CL;
  const int j = blk * CL + (int)rank;                   // 16-unit slice (same numbering as forward)
  const int kseg0 = (int)rank * (H / 2);                // first dgates column of this CTA's K segment
  const int nlo_t = HH > 0 ? HH / 128 : p.nlo_t;
  const int nlo_s = nkbs - nlo_t;
  constexpr uint32_t kb_bytes = 2u * CHB * 128u;                             // one streamed K-block of a chain: hi rows ; lo rows
  const uint32_t tile_bytes = (uint32_t)nkbs * kb_bytes;
  unsigned char* sG = smem;                                                  // [2 chains][nkbs][2 planes][16 x 128 B] dgates_t K segment
  unsigned char* sAlo = smem + (size_t)2 * tile_bytes;                       // [nlo_s][128 rows x 128 B] Wh_lo blocks outside TMEM
  float* sR = reinterpret_cast<float*>(sAlo + (size_t)nlo_s * 16384);        // [2 chains][CL src][16 units][16] partial dh
  __nv_bfloat16* sDG = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<unsigned char*>(sR) + (size_t)2 * CL * TSU * CHB * 4);
                                                                             // [2 chains][2 planes][16][4 gates][16 units]
  float* sPush = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(sDG) + (size_t)2 * 2 * CHB * 4 * TSU * 2);
                                                                             // [2 chains][CL owners][16 units][16] staged pushes (p.bulk)
  const uint32_t colLo = (uint32_t)(H / 4);              // Wh_hi: columns [0, H/4); Wh_lo: nlo_t blocks of 32 columns behind
  const uint32_t colD = colLo + (uint32_t)nlo_t * 32;    // accumulator of chain X: 32 columns at colD + 32 X
  const int t1 = a.t0 + T;                                // this launch: steps t1 - 1 down to t0
  const bool primed = t1 < a.Ttot;                        // dh_{t1-1} comes from dgates_{t1} of the previous launch
  const int ts_first = primed ? t1 : t1 - 1;              // first dgates step streamed through the tensor core

  if (threadIdx.x == 0) {
    for (int x = 0; x < 2; ++x) {
      tc::mbar_init(&full_bar[x], 1); tc::mbar_init(&tfull_bar[x], 1); tc::mbar_init(&recv_bar[x], 1);
      retry_req[x] = 0; quit[x] = 0; tiles_landed[x] = 0;
    }
    tc::fence_mbar_init();
  }
  if (warp == 8) tc::tmem_alloc(&tmem_slot, p.tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp < 4) {
    // resident Wh[128 rows of the block][K segment]: TMEM lane = threadIdx.x = hidden unit k
    const __nv_bfloat16* src = a.wh_hi + (size_t)(blk * 128 + (int)threadIdx.x) * G + kseg0;
    for (int c0 = 0; c0 < H / 4; c0 += 8) {
      const uint4 x0 = __ldg(reinterpret_cast<const uint4*>(src + 2 * c0));
      const uint4 x1 = __ldg(reinterpret_cast<const uint4*>(src + 2 * c0 + 8));
      const uint32_t r[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      tc::tmem_st8(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0, r);
    }
    const __nv_bfloat16* srcl = a.wh_lo + (size_t)(blk * 128 + (int)threadIdx.x) * G + kseg0;
    for (int c0 = 0; c0 < nlo_t * 32; c0 += 8) {
      const uint4 x0 = __ldg(reinterpret_cast<const uint4*>(srcl + 2 * c0));
      const uint4 x1 = __ldg(reinterpret_cast<const uint4*>(srcl + 2 * c0 + 8));
      const uint32_t r[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      tc::tmem_st8(tmem + ((uint32_t)(32 * warp) << 16) + colLo + (uint32_t)c0, r);
    }
    for (int kb = nlo_t; kb < nkbs; ++kb) {
      unsigned char* tile = sAlo + (size_t)(kb - nlo_t) * 16384 + (size_t)threadIdx.x * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(tile + ((c ^ (threadIdx.x & 7)) << 4)) = __ldg(reinterpret_cast<const uint4*>(srcl + kb * 64 + c * 8));
    }
    if (nlo_s > 0) tc::fence_proxy_async_smem();
    tc::tmem_st_wait();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  cluster_sync_all();                                   // peers' mbarriers are initialised before any st.async

  if (warp == 8 || warp == 10) {
    // ------------------------------------------------------------------ TMA producer of chain X
    const int X = (warp - 8) >> 1;
    if (lane == 0) tc::tma_prefetch_desc(&tmG);
    __syncwarp();
    const unsigned per_step = gridDim.x * 4u;           // every CTA adds 4 per chain and step
    const unsigned* ctr = a.barrier + 32 * X;
    unsigned epoch = 0;
    uint32_t served = 0, nf = 0;                         // retry requests answered; fetches of the schedule issued
    auto fetch = [&](int t) {
      tc::mbar_arrive_expect_tx_warp(&full_bar[X], tile_bytes);
      tc::tma_load_4d_warp(sG + (size_t)X * tile_bytes, &tmG, 0, t * B + CHB * X, 0, kseg0 / 64, &full_bar[X]);   // [kb][plane][16][64]
    };
    auto serve = [&](int t) {
      if (*reinterpret_cast<volatile uint32_t*>(&retry_req[X]) != served) { ++served; __syncwarp(); fetch(t); }
    };
    for (int t = ts_first; t >= a.t0 + 1; --t, ++nf) {  // dh_{t-1} from dgates_t
      if (t < t1) {                                       // dgates_t comes from this launch: wait for the hint
        ++epoch;
        uint32_t spins = 0;
        while (ld_relaxed_u32(ctr) < per_step * epoch) { serve(t + 1); if (++spins > (1u << 26)) __trap(); }
        __syncwarp();
      }
      if (lane == 0) RS_STAMPB(t, 8 * X + 0, gtime());
      __syncwarp();
      if (p.turns) {
        // the chains take turns at the TMA port (see rec_ts_fwd3_kernel): chain 1's fetch i follows the landing of chain 0's
        // fetch i, chain 0's fetch i + 1 that of chain 1's fetch i
        const uint32_t need = nf + (uint32_t)X;
        while (*reinterpret_cast<volatile uint32_t*>(&tiles_landed[X ^ 1]) < need) {}
        __syncwarp();
      }
      fetch(t);
      if (lane == 0) RS_STAMPB(t, 8 * X + 1, gtime());
      __syncwarp();
    }
    while (*reinterpret_cast<volatile uint32_t*>(&quit[X]) == 0) serve(a.t0 + 1);
  } else if (warp == 9 || warp == 11) {
    // ------------------------------------------------------------------ MMA issuer of chain X: one batch per fetch
    const int X = (warp - 9) >> 1;
    const uint32_t idesc = tc::instr_desc_bf16(128, CHB);                // N = one plane
    const uint32_t idesc2 = tc::instr_desc_bf16(128, 2 * CHB);           // N = both planes stacked
    const uint64_t dg0 = tc::smem_desc_sw128(tc::smem_u32(sG + (size_t)X * tile_bytes));
    const uint64_t dAlo0 = tc::smem_desc_sw128(tc::smem_u32(sAlo));
    const uint32_t tmemD = tmem + colD + (uint32_t)(2 * CHB * X);
    for (uint32_t n = 0;; ++n) {
      tc::mbar_wait(&full_bar[X], n & 1);
      if (*reinterpret_cast<volatile uint32_t*>(&quit[X]) != 0) break;
      if (lane == 0) *reinterpret_cast<volatile uint32_t*>(&tiles_landed[X]) = n + 1;
      tc::tc_fence_after();
      // D[:, 0:32) = Wh_hi [dg_hi | dg_lo] ; D[:, 0:16) += Wh_lo dg_hi
      issue_ts_blocks<6>(nkbs, tmemD, tmem, dg0, (uint64_t)(kb_bytes >> 4), idesc2, 0u);                             // 6: H = 768
      issue_ts_blocks<6>(nlo_t, tmemD, tmem + colLo, dg0, (uint64_t)(kb_bytes >> 4), idesc, 1u);
      for (int i = nlo_t; i < nkbs; ++i)
        tc::mma4_bf16_ss_warp(tmemD, dAlo0 + (uint64_t)(i - nlo_t) * (16384 >> 4), dg0 + (uint64_t)i * (kb_bytes >> 4), idesc, 1u);
      tc::mma_commit_warp(&tfull_bar[X]);
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps of chain X = warp >> 2
    const int q = warp & 3, X = warp >> 2, l16 = lane & 15, up = lane >> 4;
    const int g = lane & 3;
    const int ul = q * 4 + (l16 >> 2);                   // owned unit inside the slice (cell math)
    const int unit = j * TSU + ul;
    const int ctid = (int)threadIdx.x & 127;
    const int rowsX = min(max(B - CHB * X, 0), CHB);
    const uint32_t tmemD = tmem + colD + (uint32_t)(2 * CHB * X) + ((uint32_t)(q * 32) << 16);
    // push side: TMEM lane 32q + lane = unit 32q + lane of the block -> owner rank, unit inside its slice
    const uint32_t owner = (uint32_t)(2 * q + up);
    float* sRX = sR + (size_t)X * CL * TSU * CHB;
    __nv_bfloat16* sDGX = sDG + (size_t)X * 2 * CHB * 4 * TSU;
    const uint32_t push_base = mapa_u32(tc::smem_u32(sRX) + (uint32_t)(((int)rank * TSU + l16) * CHB) * 4u, owner);
    const uint32_t push_bar = mapa_u32(tc::smem_u32(&recv_bar[X]), owner);
    unsigned* ctr = a.barrier + 32 * X;
    const float2* blob = reinterpret_cast<const float2*>(a.gates);
    int lenr[2];
    float dc[2] = {0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int b = X * CHB + 8 * up + 2 * g + k;
      lenr[k] = b < B ? a.len[b] : 0;
    }
    float2* carry = reinterpret_cast<float2*>(a.dc_carry);
    if (primed && carry) {
      const float2 x = carry[(((size_t)j * MAXG + 0) * 8 + warp) * 32 + lane];
      dc[0] = x.x; dc[1] = x.y;
    }
    uint32_t n = primed ? 1u : 0u;                            // n - 1 = index of the receive this step consumes
    uint32_t na = 0;                                          // MMA batches consumed
    for (int t = t1 - 1; t >= a.t0; --t, ++n) {
      // operands of the cell backward (independent of the recurrence: issue before waiting)
      float2 it2[BLOB_ITEMS], cp2;
      float dy[2];
#pragma unroll
      for (int it = 0; it < BLOB_ITEMS; ++it) it2[it] = __ldg(blob + blob_idx(t, nslice, j, 1, 0, it, warp, lane));
      if (t > 0) cp2 = __ldg(blob + blob_idx(t - 1, nslice, j, 1, 0, 4, warp, lane));
      {
        float cpv[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int b = X * CHB + 8 * up + 2 * g + k;
          float d = (b < B) ? __ldg(a.dout + ((size_t)t * B + b) * H + unit) : 0.f;
          // backward of the hop's dropout(s): the same masks, applied as the gradient is read
          const unsigned long long idx = ((unsigned long long)t * B + b) * H + unit;
          if (a.drop_thr_a != 0xffffffffu) d = dropout_keep(a.drop_key, a.drop_sa, idx, a.drop_thr_a) ? d * a.drop_inv_a : 0.f;
          if (a.drop_thr_b != 0xffffffffu) d = dropout_keep(a.drop_key, a.drop_sb, idx, a.drop_thr_b) ? d * a.drop_inv_b : 0.f;
          dy[k] = d;
          cpv[k] = (t == 0 && b < B) ? __ldg(a.c0 + (size_t)b * H + unit) : 0.f;
        }
        if (t == 0) cp2 = make_float2(cpv[0], cpv[1]);
      }
      float dh[2] = {0.f, 0.f};
      if (n > 0) {
        // partial dh_t^T of this CTA's K segment -> reduce-scatter over the cluster
        if (ctid == 0) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                       ::"r"(tc::smem_u32(&recv_bar[X])), "r"((uint32_t)(CL * TSU * CHB * 4)) : "memory");
        }
        float v[16];
        const bool fresh = t + 1 < t1;                          // the tile (dgates_{t+1}) was produced by this launch
        for (;;) {
          tc::mbar_wait(&tfull_bar[X], na & 1);
          ++na;
          tc::tc_fence_after();
          if (ctid == 0) RS_STAMPB(t, 8 * X + 3, gtime());
          float w[16];
          tc::tmem_ld16(tmemD, v);
          tc::tmem_ld16(tmemD + (uint32_t)CHB, w);                // Wh_hi dg_lo
          tc::tmem_ld_wait();
          bool nan = false;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            v[i] += w[i];
            nan = nan || (i < rowsX && v[i] != v[i]);               // fill pattern in tile row i (either plane)
          }
          tc::tc_fence_before();
          // the accumulator may leave the CTA only when its tile is known to have been complete: vote
          if (!bar_any(1 + X, 128, nan)) break;
          const bool stale = bar_any(1 + X, 128, fresh && tile_has_fill(tc::smem_u32(sG + (size_t)X * tile_bytes), tile_bytes, CHB, rowsX, ctid, 128));
          if (!stale) break;                                      // the model's own NaN
          if (ctid == 0) *reinterpret_cast<volatile uint32_t*>(&retry_req[X]) = *reinterpret_cast<volatile uint32_t*>(&retry_req[X]) + 1u;
        }
        if (p.bulk) {
          // Stage the half-warp's 16 units x 16 rows (1 KB, the layout of the owner's receive block) and send it as ONE
          // bulk copy: 8 transactions of 1 KB per chain and step instead of 512 of 16 bytes.  (The staging block is
          // rewritten a step later, after a tile that every CTA's publish precedes, each of which follows the completion
          // of its receive -- this CTA's copies included.)
          float* stg = sPush + (((size_t)X * CL + owner) * TSU + l16) * CHB;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(stg + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          tc::fence_proxy_async_smem();
          __syncwarp();
          if (l16 == 0) bulk_copy_to_peer(push_base, tc::smem_u32(stg), (uint32_t)(TSU * CHB * 4), push_bar);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            st_async_v4(push_base + (uint32_t)(4 * i) * 4u, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]), push_bar);
        }
        if (ctid == 0) RS_STAMPB(t, 8 * X + 4, gtime());
        tc::mbar_wait(&recv_bar[X], (n - 1) & 1);
        if (ctid == 0) RS_STAMPB(t, 8 * X + 5, gtime());
#pragma unroll
        for (int s = 0; s < CL; ++s) {
          const float2 x = *reinterpret_cast<const float2*>(sRX + ((size_t)s * TSU + ul) * CHB + 8 * up + 2 * g);
          dh[0] += x.x; dh[1] += x.y;
        }
      }
      {
        const float* pi = reinterpret_cast<const float*>(&it2[0]);
        const float* pj = reinterpret_cast<const float*>(&it2[1]);
        const float* pf = reinterpret_cast<const float*>(&it2[2]);
        const float* po = reinterpret_cast<const float*>(&it2[3]);
        const float* pct = reinterpret_cast<const float*>(&it2[4]);
        const float* pcp = reinterpret_cast<const float*>(&cp2);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int bl = 8 * up + 2 * g + k;
          float di = 0.f, dj = 0.f, df = 0.f, dob = 0.f;
          if (t < lenr[k]) {
            const float ig = pi[k], jg = pj[k], fg = pf[k], og = po[k];
            const float dh_tot = dh[k] + dy[k];
            const float tch = fast_tanh(pct[k]);
            dob = dh_tot * tch * og * (1.f - og);
            const float dc_tot = dc[k] + dh_tot * og * (1.f - tch * tch);
            di = dc_tot * jg * ig * (1.f - ig);
            dj = dc_tot * ig * (1.f - jg * jg);
            df = dc_tot * pcp[k] * fg * (1.f - fg);
            dc[k] = dc_tot * fg;
          }
          const float d4[4] = {di, dj, df, dob};
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) {
            __nv_bfloat16 h0, l0;
            tc::split_bf16(d4[gg], h0, l0);
            sDGX[(((size_t)0 * CHB + bl) * 4 + gg) * TSU + ul] = h0;
            sDGX[(((size_t)1 * CHB + bl) * 4 + gg) * TSU + ul] = l0;
          }
        }
      }
      if (ctid == 0) RS_STAMPB(t, 8 * X + 6, gtime());
      chain_bar_sync(X);
      // publish dgates_t: 16-byte stores of the staged tile (one 32-byte sector per (row, gate, plane)), then the hint
      bool faulty = false;
      if constexpr (STAMP) faulty = p.fault != 0;
      {
        uint4 x[2];
        __nv_bfloat16* dst[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int ch = ctid + 128 * i;                          // 256 chunks: [plane][row][gate][half]
          const int half = ch & 1, gg = (ch >> 1) & 3, bl = (ch >> 3) & 15, pl = ch >> 7;
          x[i] = *reinterpret_cast<const uint4*>(sDGX + (((size_t)pl * CHB + bl) * 4 + gg) * TSU + half * 8);
          dst[i] = bl < rowsX ? (pl ? a.dg_lo : a.dg_hi) + ((size_t)t * B + CHB * X + bl) * G + (size_t)gg * H + j * TSU + half * 8 : nullptr;
        }
        if (!faulty) {
#pragma unroll
          for (int i = 0; i < 2; ++i) if (dst[i]) st_relaxed_v4(dst[i], x[i]);
          __syncwarp();
          if (lane == 0 && t > a.t0) red_relaxed_add(ctr, 1u);
        } else {
          // test of the retry path: the hint overtakes the odd chunks by 3 us
          const bool late = ctid & 1;
          if (!late) for (int i = 0; i < 2; ++i) if (dst[i]) st_relaxed_v4(dst[i], x[i]);
          __syncwarp();
          if (lane == 0 && t > a.t0) red_relaxed_add(ctr, 1u);
          const unsigned long long until = gtime() + 3000ull;
          while (gtime() < until) {}
          __syncwarp();
          if (late) for (int i = 0; i < 2; ++i) if (dst[i]) st_relaxed_v4(dst[i], x[i]);
        }
      }
      if (ctid == 0) { RS_STAMPB(t, 8 * X + 7, gtime()); if (X == 0) RS_STAMPB(t, 2, (unsigned long long)na); }
      // (the staged tile is rewritten only after the next tfull wait, i.e. after a fetch that follows the grid-wide
      //  count, which needs all four adds above, each of which follows its warp's reads of the tile)
    }
    if (ctid == 0) {
      // release the chain's producer and MMA warps
      *reinterpret_cast<volatile uint32_t*>(&quit[X]) = 1u;
      tc::mbar_arrive(&full_bar[X]);
    }
    if (carry) carry[(((size_t)j * MAXG + 0) * 8 + warp) * 32 + lane] = make_float2(dc[0], dc[1]);
  }
  tc::tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 8) tc::tmem_dealloc(tmem, p.tmem_cols);
}

}  // namespace

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
// RS_TS_VARIANT switches (default 317; 0 = the conservative protocol with generic stores, red.release /
// ld.acquire and proxy fences on both sides):
//   1  no writer-side fence.proxy.async (generic-store publish)   4  relaxed polling without an acquire fence
//   8  no reader-side fence.proxy.async                          16  publish with TMA stores + bulk-group wait
//  32  after the bulk-group wait: fence.proxy.async + red.RELEASE.gpu (64: no fence, 128: relaxed red, 256: the
//      fence is fence.proxy.async.global)
// The writer side needs 32.  Completion of a bulk group makes the stored tile visible to the issuing thread only;
// with a relaxed signal (the round's earlier default, 29) other CTAs' TMA loads occasionally fetched the previous
// contents of the tile -- invisible to a bitwise comparison of two runs on the same input (stale == fresh), caught
// by tests/gpu_diag.py stress (alternating inputs): 36 of 38 backward passes differed.  With the release the
// stress run is clean with or without the reader-side fences (tests/test_gpu_model.py::test_stale_tile_stress).
constexpr int kDefaultVariant = 317;
static int ts_variant() {
  const char* v = getenv("RS_TS_VARIANT");
  return v ? atoi(v) : kDefaultVariant;
}
static bool ts_enabled() {
  const char* v = getenv("RS_REC_TS");
  return !(v && v[0] == '0');
}

bool rec_ts_geometry(int H, int B, RecTcGeom* g) {
  if (!ts_enabled()) return false;
  if (H % 128 != 0 || H < 128 || B < 1 || B > 64) return false;
  if (H / TSU > sm_count()) return false;
  const int Bpad = (B + 15) / 16 * 16;
  const int nkb = H / 64;
  int nkb_t = (512 - 2 * Bpad) / 32;
  if (nkb_t > nkb) nkb_t = nkb;
  // K-blocks per TMA box / mbarrier.  Measured (cfg-2): ONE box carrying all of h_{t-1} lands sooner than the
  // same bytes split over several TMA instructions (they are served one after the other), and the MMAs of an
  // early group slow the arrival of the later ones; so take the whole of h when it fits.
  const size_t budget = 227 * 1024 - 8192;
  const size_t a_bytes = (size_t)(nkb - nkb_t) * 16384;
  int gkb = nkb;
  if (a_bytes + (size_t)gkb * 2 * Bpad * 128 > budget || gkb > 256) gkb = (nkb % 4 == 0) ? 4 : (nkb % 3 == 0) ? 3 : 2;
  { const char* v = getenv("RS_TS_GKB"); if (v && atoi(v) > 0 && nkb % atoi(v) == 0) gkb = atoi(v); }
  const int ngroups = nkb / gkb;
  const size_t slot = (size_t)gkb * 2 * Bpad * 128;
  if (a_bytes + slot > budget) return false;
  int slots = (int)((budget - a_bytes) / slot);
  if (slots > MAXSLOTS) slots = MAXSLOTS;
  if (slots > ngroups) slots = ngroups;
  g->H = H; g->B = B; g->Bpad = Bpad; g->U = TSU; g->nslice = H / TSU; g->stages = slots;
  g->smem_bytes = a_bytes + (size_t)slots * slot + 1024;
  if (g->smem_bytes < 120 * 1024) g->smem_bytes = 120 * 1024;      // one CTA per SM (each allocates all of TMEM)
  g->ts = 1;
  g->gkb = gkb;
  g->nkb_t = nkb_t;
  g->ngl = (Bpad / 16 + 1) / 2;
  // backward: TMEM columns = H/4 (Wh_hi segment) + 2 Bpad (accumulator of the stacked planes) must fit; Wh_lo takes
  // what is left and shared memory beyond that
  if (H / 4 + 2 * Bpad > 512) return false;
  return true;
}

size_t rec_ts_blob_floats(const RecTcGeom& g, int T) {
  return (size_t)T * g.nslice * g.ngl * BLOB_ITEMS * 8 * 32 * 2;
}

size_t rec_ts_dc_carry_floats(const RecTcGeom& g) { return (size_t)g.nslice * MAXG * 8 * 32 * 2; }

int lstm_rec_ts_forward(const RecTcGeom& g, const RecTcFwdArgs& a_in, cudaStream_t st) {
  RecTcFwdArgs a = a_in;
  if (a.Ttot <= 0) { a.Ttot = a.T; a.t0 = 0; }
  RS_REQUIRE(a.T > 0 && a.t0 >= 0 && a.t0 + a.T <= a.Ttot, RS_ERR_INVALID, "lstm_rec_ts_forward: T=%d t0=%d Ttot=%d", a.T, a.t0, a.Ttot);
  RS_REQUIRE(a.h_ld == 2 * g.H && a.h_lo == a.h_hi + g.H, RS_ERR_INVALID,
             "lstm_rec_ts_forward: h planes must be row-interleaved ([rows][hi H | lo H])");
  CUtensorMap th, ts;
  int rc;
  const int hrows = (a.Ttot + 1) * g.B;
  // two independent 16-row chains per CTA (rec_ts_fwd2_kernel) when the batch allows it; RS_TS_CHAINS=0: one chain
  static const bool chains_env = [] { const char* v = getenv("RS_TS_CHAINS"); return !(v && v[0] == '0'); }();
  const bool shape2 = chains_env && ts_variant() == kDefaultVariant && g.Bpad == 32 && g.B > CHB && g.gkb == g.H / 64 && g.stages == 1;
  // (rec_ts_fwd3_kernel also takes batches of at most 16 rows, as one chain; RS_TS_XCHG16=0: the round-1 kernel for those)
  static const bool xchg16_env = [] { const char* v = getenv("RS_TS_XCHG16"); return !(v && v[0] == '0'); }();
  const bool shape1 = xchg16_env && chains_env && ts_variant() == kDefaultVariant && g.Bpad == 16 && g.gkb == g.H / 64 && g.stages == 1;
  const bool two_chains = shape2 && !a.dbg;
  // self-validating exchange (rec_ts_fwd3_kernel; RS_TS_XCHG=0: the counter + TMA exchange of rec_ts_fwd2_kernel)
  static const bool xchg_env = [] { const char* v = getenv("RS_TS_XCHG"); return !(v && v[0] == '0'); }();
  if ((shape2 || shape1) && xchg_env) {
    CUtensorMap td0, td1;
    if ((rc = tmap_stacked_bf16(&th, a.h_hi, hrows, g.H, CHB, g.gkb)) != RS_OK) return rc;
    td0 = th; td1 = th;
    if (a.drop_hi) {
      RS_REQUIRE(a.drop_lo > a.drop_hi, RS_ERR_INVALID, "lstm_rec_ts_forward: the low hop plane must follow the high one");
      const size_t pstride = (size_t)((const char*)a.drop_lo - (const char*)a.drop_hi);
      if ((rc = tmap_store3_bf16(&td0, a.drop_hi, g.H, 2, a.Ttot * g.B, pstride, (size_t)g.H * 2, TSU, 2, g.B < CHB ? g.B : CHB)) != RS_OK) return rc;
      td1 = td0;
      if (g.B > CHB && (rc = tmap_store3_bf16(&td1, a.drop_hi, g.H, 2, a.Ttot * g.B, pstride, (size_t)g.H * 2, TSU, 2, g.B - CHB)) != RS_OK) return rc;
    }
    // slots 1..Ttot of the planes carry the fill pattern until their producers overwrite it (first launch of a pass)
    if (a.t0 == 0)
      RS_CHECK_CUDA(cudaMemsetAsync(a.h_hi + (size_t)g.B * a.h_ld, 0xff, (size_t)a.Ttot * g.B * a.h_ld * sizeof(__nv_bfloat16), st));
    KFwd3 p3;
    p3.a = a;
    p3.H = g.H; p3.B = g.B; p3.Bpad = g.Bpad; p3.nslice = g.nslice; p3.nkb = g.H / 64; p3.nkb_t = g.nkb_t;
    static const int turns_env = [] { const char* v = getenv("RS_TS_TURNS"); return v ? atoi(v) : 1; }();
    p3.turns = turns_env;
    { const char* v = getenv("RS_TS_FAULT"); p3.fault = (a.dbg && v && v[0] == '1') ? 1 : 0; }
    static const int endbar_env = [] { const char* v = getenv("RS_TS_ENDBAR"); return v ? atoi(v) : 0; }();
    p3.endbar = (endbar_env || p3.fault) ? 1 : 0;
    RS_CHECK_CUDA(cudaMemsetAsync(a.barrier, 0, 256 * sizeof(unsigned), st));
    const bool fast3 = g.H == 768 && g.nkb_t == 12;
    auto kern3 = a.dbg ? (fast3 ? rec_ts_fwd3_kernel<768, true> : rec_ts_fwd3_kernel<0, true>) : fast3 ? rec_ts_fwd3_kernel<768, false> : rec_ts_fwd3_kernel<0, false>;
    static size_t checked3_smem[kMaxDevices * 4] = {};
    static int checked3_cap[kMaxDevices * 4] = {};
    const int s3 = device_slot() * 4 + (a.dbg ? (fast3 ? 3 : 2) : fast3 ? 1 : 0);
    if (checked3_smem[s3] != g.smem_bytes) {
      int per_sm = 0;
      RS_CHECK_CUDA(cudaFuncSetAttribute(kern3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes));
      RS_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern3, NTHREADS2, g.smem_bytes));
      checked3_cap[s3] = per_sm * sm_count();
      checked3_smem[s3] = g.smem_bytes;
    }
    RS_REQUIRE(checked3_cap[s3] >= g.nslice, RS_ERR_UNSUPPORTED, "lstm_rec_ts_forward: %d CTAs cannot be co-resident", g.nslice);
    kern3<<<dim3(g.nslice), dim3(NTHREADS2), g.smem_bytes, st>>>(th, td0, td1, p3);
    RS_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return RS_OK;
  }
  if (two_chains) {
    CUtensorMap ts0, ts1, td0, td1;
    if ((rc = tmap_stacked_bf16(&th, a.h_hi, hrows, g.H, CHB, g.gkb)) != RS_OK) return rc;
    if ((rc = tmap_store3_bf16(&ts0, a.h_hi, g.H, 2, hrows, (size_t)g.H * 2, (size_t)2 * g.H * 2, TSU, 2, CHB)) != RS_OK) return rc;
    if ((rc = tmap_store3_bf16(&ts1, a.h_hi, g.H, 2, hrows, (size_t)g.H * 2, (size_t)2 * g.H * 2, TSU, 2, g.B - CHB)) != RS_OK) return rc;
    td0 = ts0; td1 = ts1;
    if (a.drop_hi) {
      RS_REQUIRE(a.drop_lo > a.drop_hi, RS_ERR_INVALID, "lstm_rec_ts_forward: the low hop plane must follow the high one");
      const size_t pstride = (size_t)((const char*)a.drop_lo - (const char*)a.drop_hi);
      if ((rc = tmap_store3_bf16(&td0, a.drop_hi, g.H, 2, a.Ttot * g.B, pstride, (size_t)g.H * 2, TSU, 2, CHB)) != RS_OK) return rc;
      if ((rc = tmap_store3_bf16(&td1, a.drop_hi, g.H, 2, a.Ttot * g.B, pstride, (size_t)g.H * 2, TSU, 2, g.B - CHB)) != RS_OK) return rc;
    }
    KFwd2 p2;
    p2.a = a;
    p2.H = g.H; p2.B = g.B; p2.nslice = g.nslice; p2.nkb = g.H / 64; p2.nkb_t = g.nkb_t; p2.rows1 = g.B - CHB;
    RS_CHECK_CUDA(cudaMemsetAsync(a.barrier, 0, 256 * sizeof(unsigned), st));
    const bool fast2 = g.H == 768 && g.nkb_t == 12;
    auto kern2 = fast2 ? rec_ts_fwd2_kernel<768> : rec_ts_fwd2_kernel<0>;
    static size_t checked2_smem[kMaxDevices * 2] = {};
    static int checked2_cap[kMaxDevices * 2] = {};
    const int s2 = device_slot() * 2 + (fast2 ? 1 : 0);
    if (checked2_smem[s2] != g.smem_bytes) {
      int per_sm = 0;
      RS_CHECK_CUDA(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes));
      RS_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern2, NTHREADS2, g.smem_bytes));
      checked2_cap[s2] = per_sm * sm_count();
      checked2_smem[s2] = g.smem_bytes;
    }
    RS_REQUIRE(checked2_cap[s2] >= g.nslice, RS_ERR_UNSUPPORTED, "lstm_rec_ts_forward: %d CTAs cannot be co-resident", g.nslice);
    kern2<<<dim3(g.nslice), dim3(NTHREADS2), g.smem_bytes, st>>>(th, ts0, ts1, td0, td1, p2);
    RS_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return RS_OK;
  }
  if ((rc = tmap_stacked_bf16(&th, a.h_hi, hrows, g.H, g.Bpad, g.gkb)) != RS_OK) return rc;
  if ((rc = tmap_store3_bf16(&ts, a.h_hi, g.H, 2, hrows, (size_t)g.H * 2, (size_t)2 * g.H * 2, TSU, 2, g.B)) != RS_OK) return rc;
  CUtensorMap td = ts;
  if (a.drop_hi) {
    RS_REQUIRE(a.drop_lo > a.drop_hi, RS_ERR_INVALID, "lstm_rec_ts_forward: the low hop plane must follow the high one");
    if ((rc = tmap_store3_bf16(&td, a.drop_hi, g.H, 2, a.Ttot * g.B, (size_t)((const char*)a.drop_lo - (const char*)a.drop_hi),
                               (size_t)g.H * 2, TSU, 2, g.B)) != RS_OK) return rc;
  }
  KFwd p;
  p.a = a;
  p.H = g.H; p.B = g.B; p.Bpad = g.Bpad; p.nslice = g.nslice; p.slots = g.stages; p.nkb = g.H / 64; p.nkb_t = g.nkb_t;
  p.gkb = g.gkb;
  p.ngroups = p.nkb / g.gkb;
  p.ngl = g.ngl;
  p.kb_bytes = 2u * (uint32_t)g.Bpad * 128;
  p.variant = ts_variant();
  RS_CHECK_CUDA(cudaMemsetAsync(a.barrier, 0, 256 * sizeof(unsigned), st));
  // the benchmark shape's instantiation: H = 768 (all twelve K-blocks in tensor memory, one TMA group), Bpad = 32
  const bool fast = !a.dbg && p.variant == kDefaultVariant && g.Bpad == 32 && g.H == 768 && g.gkb == 12 && g.nkb_t == 12 && g.stages == 1;
  auto kern = a.dbg ? rec_ts_fwd_kernel<true, -1, 0, 0> : (fast ? rec_ts_fwd_kernel<false, kDefaultVariant, 32, 768> : rec_ts_fwd_kernel<false, -1, 0, 0>);
  const int si = device_slot() * 3 + (a.dbg ? 1 : (fast ? 2 : 0));
  static size_t checked_smem[kMaxDevices * 3] = {};  // attribute + co-residency check once per device and shared-memory size
  static int checked_cap[kMaxDevices * 3] = {};
  if (checked_smem[si] != g.smem_bytes) {
    int per_sm = 0;
    RS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes));
    RS_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NTHREADS, g.smem_bytes));
    checked_cap[si] = per_sm * sm_count();
    checked_smem[si] = g.smem_bytes;
  }
  RS_REQUIRE(checked_cap[si] >= g.nslice, RS_ERR_UNSUPPORTED, "lstm_rec_ts_forward: %d CTAs cannot be co-resident", g.nslice);
  // A plain launch: the grid barrier is the kernel's own counter, co-residency was checked above (one CTA per SM,
  // nslice <= SM count), and plain launches of different layers run side by side (RS_TS_COOP=1: cooperative launch).
  static const bool coop = [] { const char* v = getenv("RS_TS_COOP"); return v && v[0] == '1'; }();
  if (coop) {
    void* kargs[] = {(void*)&th, (void*)&ts, (void*)&td, (void*)&p};
    RS_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(g.nslice), dim3(NTHREADS), kargs, g.smem_bytes, st));
  } else {
    kern<<<dim3(g.nslice), dim3(NTHREADS), g.smem_bytes, st>>>(th, ts, td, p);
    RS_CHECK_CUDA(cudaGetLastError());
  }
  count_launch();
  return RS_OK;
}

int lstm_rec_ts_backward(const RecTcGeom& g, const RecTcBwdArgs& a_in, cudaStream_t st) {
  RecTcBwdArgs a = a_in;
  if (a.Ttot <= 0) { a.Ttot = a.T; a.t0 = 0; }
  RS_REQUIRE(a.T > 0 && a.t0 >= 0 && a.t0 + a.T <= a.Ttot, RS_ERR_INVALID, "lstm_rec_ts_backward: T=%d t0=%d Ttot=%d", a.T, a.t0, a.Ttot);
  RS_REQUIRE(a.t0 + a.T == a.Ttot || a.dc_carry, RS_ERR_INVALID, "lstm_rec_ts_backward: a continued launch needs dc_carry");
  // bf16x3 (default) needs both weight planes; RS_BWD_X3=0 or a missing low plane: one bf16 product
  static const bool x3_env = [] { const char* v = getenv("RS_BWD_X3"); return !(v && v[0] == '0'); }();
  const bool x3 = x3_env && a.wh_lo != nullptr;
  KBwd p;
  p.a = a;
  p.H = g.H; p.B = g.B; p.Bpad = g.Bpad; p.nslice = g.nslice; p.nkbs = g.H / 2 / 64; p.ngl = g.ngl;
  p.x3 = x3 ? 1 : 0;
  // tensor-memory columns: Wh_hi (H/4) + as many 32-column blocks of Wh_lo as fit + the accumulator (N = planes x Bpad)
  const int dcols = (x3 ? 2 : 1) * g.Bpad;
  int nlo_t = 0;
  if (x3) {
    nlo_t = (512 - g.H / 4 - dcols) / 32;
    if (nlo_t > p.nkbs) nlo_t = p.nkbs;
    RS_REQUIRE(nlo_t >= 0, RS_ERR_UNSUPPORTED, "lstm_rec_ts_backward: H=%d does not fit tensor memory", g.H);
  }
  p.nlo_t = nlo_t;
  const int nlo_s = x3 ? p.nkbs - nlo_t : 0;
  uint32_t cols = 32;
  while ((int)cols < g.H / 4 + nlo_t * 32 + dcols) cols <<= 1;
  p.tmem_cols = cols;
  p.variant = ts_variant();
  p.fault = 0;
  p.turns = 0;
  CUtensorMap tg;
  int rc;
  if (x3) {
    RS_REQUIRE(a.dg_lo > a.dg_hi, RS_ERR_INVALID, "lstm_rec_ts_backward: the low dgates plane must follow the high one");
    if ((rc = tmap_stacked2_bf16(&tg, a.dg_hi, (size_t)((const char*)a.dg_lo - (const char*)a.dg_hi), a.Ttot * g.B, 4 * g.H,
                                 4 * g.H, g.Bpad, p.nkbs)) != RS_OK) return rc;
  } else {
    if ((rc = tmap_2d_bf16(&tg, a.dg_hi, a.Ttot * g.B, 4 * g.H, 4 * g.H, g.Bpad)) != RS_OK) return rc;
  }
  size_t smem = (size_t)p.nkbs * g.Bpad * 128 * (x3 ? 2 : 1) + (size_t)nlo_s * 16384 + (size_t)CL * TSU * g.Bpad * 4 +
                (size_t)2 * g.Bpad * 4 * TSU * 2 + 1024;
  RS_REQUIRE(smem <= 227 * 1024 - 2048, RS_ERR_UNSUPPORTED, "lstm_rec_ts_backward: H=%d B=%d needs %zu bytes of shared memory", g.H, g.B, smem);
  // (two of these cannot share an SM: 2 x 320 x 128 registers exceed the register file)
  static const bool share = [] { const char* v = getenv("RS_TC_CORES"); return v && v[0] == '1'; }();
  // keep the SM to this CTA unless sharing is asked for (RS_TC_CORES=1) and a GEMM CTA can really share it (256 free
  // tensor-memory columns); sharing measured slower at cfg-2, see lstm_tc.cu
  if ((!share || cols > 256) && smem < 120 * 1024) smem = 120 * 1024;
  RS_CHECK_CUDA(cudaMemsetAsync(a.barrier, 0, 256 * sizeof(unsigned), st));
  // validated exchange (rec_ts_bwd3_kernel): RS_TS_XCHG_BWD=1.  Measured slower than TMA stores + release + counter so far
  // (5.7 vs 5.3 us per step at cfg-2: the vote before the push is one more barrier in the longest phase), so off by default.
  // two chains per CTA (rec_ts_bwd4_kernel) when the batch allows it; RS_TS_CHAINS_BWD=0: one chain
  static const bool chains_bwd_env = [] { const char* v = getenv("RS_TS_CHAINS_BWD"); return !(v && v[0] == '0'); }();
  if (x3 && chains_bwd_env && p.variant == kDefaultVariant && g.Bpad == 32 && g.B > CHB) {
    CUtensorMap tg2;
    if ((rc = tmap_stacked2_bf16(&tg2, a.dg_hi, (size_t)((const char*)a.dg_lo - (const char*)a.dg_hi), a.Ttot * g.B, 4 * g.H,
                                 4 * g.H, CHB, p.nkbs)) != RS_OK) return rc;
    // the launch's own rows of the dgates planes carry the fill pattern until their producers overwrite it
    if (!a.prefilled) {
      RS_CHECK_CUDA(cudaMemsetAsync(a.dg_hi + (size_t)a.t0 * g.B * 4 * g.H, 0xff, (size_t)a.T * g.B * 4 * g.H * sizeof(__nv_bfloat16), st));
      RS_CHECK_CUDA(cudaMemsetAsync(a.dg_lo + (size_t)a.t0 * g.B * 4 * g.H, 0xff, (size_t)a.T * g.B * 4 * g.H * sizeof(__nv_bfloat16), st));
    }
    { const char* v = getenv("RS_TS_FAULT"); p.fault = (a.dbg && v && v[0] == '1') ? 1 : 0; }
    static const int turns_env = [] { const char* v = getenv("RS_TS_TURNS_BWD"); return v ? atoi(v) : 1; }();
    p.turns = turns_env;
    // shared memory: two tiles, the Wh_lo blocks outside tensor memory, two receive buffers, two staged dgates tiles
    static const int bulk_env = [] { const char* v = getenv("RS_TS_PUSH_BULK"); return v ? atoi(v) : 0; }();
    p.bulk = bulk_env;
    size_t smem4 = (size_t)2 * p.nkbs * 2 * CHB * 128 + (size_t)nlo_s * 16384 + (size_t)2 * CL * TSU * CHB * 4 +
                   (size_t)2 * 2 * CHB * 4 * TSU * 2 + (size_t)2 * CL * TSU * CHB * 4 + 1024;
    if (smem4 < 120 * 1024) smem4 = 120 * 1024;          // one CTA per SM
    const bool fast4 = g.H == 768 && nlo_t == p.nkbs;
    const int s4 = device_slot() * 4 + (a.dbg ? (fast4 ? 3 : 2) : fast4 ? 1 : 0);
    auto kern4 = a.dbg ? (fast4 ? rec_ts_bwd4_kernel<true, 768> : rec_ts_bwd4_kernel<true, 0>)
                       : fast4 ? rec_ts_bwd4_kernel<false, 768> : rec_ts_bwd4_kernel<false, 0>;
    static size_t attr4_smem[kMaxDevices * 4] = {};
    static int nclusters4[kMaxDevices * 4] = {};
    cudaLaunchConfig_t cfg4 = {};
    cfg4.gridDim = dim3(g.nslice);
    cfg4.blockDim = dim3(NTHREADS2);
    cfg4.dynamicSmemBytes = smem4;
    cfg4.stream = st;
    cudaLaunchAttribute attr4[1];
    attr4[0].id = cudaLaunchAttributeClusterDimension;
    attr4[0].val.clusterDim.x = CL; attr4[0].val.clusterDim.y = 1; attr4[0].val.clusterDim.z = 1;
    cfg4.attrs = attr4;
    cfg4.numAttrs = 1;
    if (attr4_smem[s4] != smem4) {
      RS_CHECK_CUDA(cudaFuncSetAttribute(kern4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
      RS_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&nclusters4[s4], kern4, &cfg4));
      attr4_smem[s4] = smem4;
    }
    RS_REQUIRE(nclusters4[s4] * CL >= g.nslice, RS_ERR_UNSUPPORTED, "lstm_rec_ts_backward: %d CTAs cannot be co-resident (%d clusters)",
               g.nslice, nclusters4[s4]);
    RS_CHECK_CUDA(cudaLaunchKernelEx(&cfg4, kern4, tg2, p));
    count_launch();
    return RS_OK;
  }
  static const bool xchg_env = [] { const char* v = getenv("RS_TS_XCHG_BWD"); return v && v[0] == '1'; }();
  const char* fault_env = getenv("RS_TS_FAULT");
  if (x3 && (xchg_env || (a.dbg && fault_env && fault_env[0] == '1')) && p.variant == kDefaultVariant) {
    // the launch's own rows of the dgates planes carry the fill pattern until their producers overwrite it
    if (!a.prefilled) {
      RS_CHECK_CUDA(cudaMemsetAsync(a.dg_hi + (size_t)a.t0 * g.B * 4 * g.H, 0xff, (size_t)a.T * g.B * 4 * g.H * sizeof(__nv_bfloat16), st));
      RS_CHECK_CUDA(cudaMemsetAsync(a.dg_lo + (size_t)a.t0 * g.B * 4 * g.H, 0xff, (size_t)a.T * g.B * 4 * g.H * sizeof(__nv_bfloat16), st));
    }
    { const char* v = getenv("RS_TS_FAULT"); p.fault = (a.dbg && v && v[0] == '1' && g.Bpad <= 32) ? 1 : 0; }
    const bool fast3 = g.Bpad == 32 && g.H == 768 && nlo_t == p.nkbs;
    const int s3 = device_slot() * 4 + (a.dbg ? (fast3 ? 3 : 2) : fast3 ? 1 : 0);
    auto kern3 = a.dbg ? (fast3 ? rec_ts_bwd3_kernel<true, 32, 768> : rec_ts_bwd3_kernel<true, 0, 0>)
                       : fast3 ? rec_ts_bwd3_kernel<false, 32, 768> : rec_ts_bwd3_kernel<false, 0, 0>;
    static size_t attr3_smem[kMaxDevices * 4] = {};
    static int nclusters3[kMaxDevices * 4] = {};
    cudaLaunchConfig_t cfg3 = {};
    cfg3.gridDim = dim3(g.nslice);
    cfg3.blockDim = dim3(NTHREADS);
    cfg3.dynamicSmemBytes = smem;
    cfg3.stream = st;
    cudaLaunchAttribute attr3[1];
    attr3[0].id = cudaLaunchAttributeClusterDimension;
    attr3[0].val.clusterDim.x = CL; attr3[0].val.clusterDim.y = 1; attr3[0].val.clusterDim.z = 1;
    cfg3.attrs = attr3;
    cfg3.numAttrs = 1;
    if (attr3_smem[s3] != smem) {
      RS_CHECK_CUDA(cudaFuncSetAttribute(kern3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      RS_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&nclusters3[s3], kern3, &cfg3));
      attr3_smem[s3] = smem;
    }
    RS_REQUIRE(nclusters3[s3] * CL >= g.nslice, RS_ERR_UNSUPPORTED, "lstm_rec_ts_backward: %d CTAs cannot be co-resident (%d clusters)",
               g.nslice, nclusters3[s3]);
    RS_CHECK_CUDA(cudaLaunchKernelEx(&cfg3, kern3, tg, p));
    count_launch();
    return RS_OK;
  }
  const bool fast = !a.dbg && p.variant == kDefaultVariant && g.Bpad == 32 && g.H == 768 && (!x3 || nlo_t == p.nkbs);
  const int si = device_slot() * 4 + (a.dbg ? 1 : (fast ? (x3 ? 2 : 3) : 0));
  auto kern = a.dbg ? rec_ts_bwd_kernel<true, -1, 0, 0, -1>
                    : (fast ? (x3 ? rec_ts_bwd_kernel<false, kDefaultVariant, 32, 768, 1> : rec_ts_bwd_kernel<false, kDefaultVariant, 32, 768, 0>)
                            : rec_ts_bwd_kernel<false, -1, 0, 0, -1>);
  static size_t attr_smem[kMaxDevices * 4] = {};
  if (attr_smem[si] != smem) {
    RS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(g.nslice);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static int nclusters_s[kMaxDevices * 4] = {};
  if (attr_smem[si] != smem) {
    RS_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&nclusters_s[si], kern, &cfg));
    attr_smem[si] = smem;
  }
  const int nclusters = nclusters_s[si];
  RS_REQUIRE(nclusters * CL >= g.nslice, RS_ERR_UNSUPPORTED, "lstm_rec_ts_backward: %d CTAs cannot be co-resident (%d clusters)",
             g.nslice, nclusters);
  CUtensorMap ts_hi, ts_lo;
  if ((rc = tmap_store3_bf16(&ts_hi, a.dg_hi, g.H, 4, a.Ttot * g.B, (size_t)g.H * 2, (size_t)4 * g.H * 2, TSU, 4, g.B)) != RS_OK) return rc;
  if ((rc = tmap_store3_bf16(&ts_lo, a.dg_lo, g.H, 4, a.Ttot * g.B, (size_t)g.H * 2, (size_t)4 * g.H * 2, TSU, 4, g.B)) != RS_OK) return rc;
  RS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tg, ts_hi, ts_lo, p));
  count_launch();
  return RS_OK;
}

}  // namespace rs
