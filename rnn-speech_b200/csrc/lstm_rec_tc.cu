// Tensor-core persistent recurrent LSTM kernels (forward and backward), sm_100a.
//
// Replaces the tf.nn.dynamic_rnn loop over BasicLSTMCell
// (/root/reference/models/AcousticModel.py:227-237, :277-278) and its BPTT for one layer;
// same semantics as lstm_rec.cu / oracle/model.py.
//
// Forward.  CTA j owns hidden units [jU, jU+U) (U = 8 or 16).  Its 4U gate rows of Wh^T
// (bf16 hi + lo planes, K-major, SWIZZLE_128B) are loaded ONCE by TMA and stay in shared
// memory for all T steps; the cell state stays in registers.  Per step:
//   producer thread : waits on the grid barrier (all CTAs published h_{t-1}), then streams
//                     h_{t-1} (hi + lo planes, [Bpad x H]) with TMA, 64 columns a box, four
//                     boxes (one "group") per mbarrier -- the whole of h_{t-1} is in flight
//   MMA thread      : D[64 x Bpad] (TMEM, fp32) = W_hi h_hi + W_hi h_lo + W_lo h_hi with
//                     tcgen05.mma M = 64 (rows >= 4U are don't-care), N = Bpad, K = 16
//   epilogue warps  : tcgen05.ld the accumulator, add the hoisted input projection gx, apply
//                     the gate non-linearities (row m = 4*unit + gate, so a lane quad holds
//                     i,j,f,o of one unit), exchange inside the quad by shuffles, update c,
//                     publish h_t as bf16 hi/lo planes, arrive on the grid barrier.
// Backward.  CTA j owns hidden units [16j, 16j+16): its 16 rows of Wh ([16][4H] bf16) are
// the resident B operand (N = 16); dgates_t ([Bpad x 4H] bf16) streams through a TMA ring as
// the A operand (M = 64), D[b][u] = dh_{t-1}.
//
// M = 64 accumulator layout (cta_group::1): row m lives in TMEM lane 32*(m/16) + m%16, i.e.
// 16 lanes of each of the four 32-lane sub-partitions; epilogue warp w reads sub-partition
// w%4 and only its lanes 0..15 carry rows.
//
// Measured on B200 (tests/gpu_diag.py timeline): one tcgen05.mma with a 128-row SS operand
// costs ~44 cycles (shared-memory operand bandwidth), one mbarrier wait + commit round ~400
// cycles: hence M = 64 and four K-blocks per barrier.
#include "lstm_rec_tc.cuh"
#include "tc_common.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace rs {
namespace {

constexpr int NTHREADS = 320;   // warps 0-7: epilogue (sub-partition w%4, column half w/4); 8: TMA; 9: MMA
constexpr int NEPI = 256;
constexpr int MAXG = 2;         // 16-column groups per epilogue thread (Bpad <= 64, two column halves)
constexpr int MAXSLOTS = 16;    // ring slots (groups of GKB K-blocks)
constexpr int GKB = 4;          // K-blocks (64 columns each) per mbarrier

struct KArgs {
  RecTcFwdArgs a;
  int H, B, Bpad, U, nslice, slots, nkb, ngroups;
  int variant;                  // bit0: per-CTA flags (else one atomic counter); bit1: reserve stores before the signal
  uint32_t a_kb_bytes;          // one K-block of the resident operand: 4U rows x 128 B
  uint32_t kb_bytes;            // one streamed K-block: 2 planes x Bpad rows x 128 B
};

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
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Grid barrier without atomics: CTA j publishes its step count in flags[j]; a whole warp polls
// all flags with coalesced acquire loads (one L2 round trip per poll, no same-address traffic).
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void wait_flags_warp(const unsigned* flags, unsigned nctas, unsigned target, int lane,
                                                unsigned long long* dbg = nullptr, int step = 0) {
  // relaxed polling (all flag loads of one poll in flight together), one acquire fence at the end
  for (;;) {
    unsigned v0 = target, v1 = target, v2 = target, v3 = target, v4 = target;
    if ((unsigned)lane < nctas) v0 = ld_relaxed_u32(flags + lane);
    if ((unsigned)lane + 32 < nctas) v1 = ld_relaxed_u32(flags + lane + 32);
    if ((unsigned)lane + 64 < nctas) v2 = ld_relaxed_u32(flags + lane + 64);
    if ((unsigned)lane + 96 < nctas) v3 = ld_relaxed_u32(flags + lane + 96);
    if ((unsigned)lane + 128 < nctas) v4 = ld_relaxed_u32(flags + lane + 128);
    const bool ok = v0 >= target && v1 >= target && v2 >= target && v3 >= target && v4 >= target;
    if (__all_sync(0xffffffffu, ok)) break;
  }
  if (dbg && lane == 0 && blockIdx.x == 0) dbg[(size_t)step * 16 + 15] = gtime();
  __syncwarp();
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
#define RS_STAMP(dbgp, step, ev) do { if ((dbgp) && blockIdx.x == 0) (dbgp)[(size_t)(step) * 16 + (ev)] = gtime(); } while (0)
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ float pick4(int sel, float a0, float a1, float a2, float a3) {
  const float lo = (sel & 1) ? a1 : a0;
  const float hi = (sel & 1) ? a3 : a2;
  return (sel & 2) ? hi : lo;
}

__global__ void __launch_bounds__(NTHREADS, 1)
rec_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                  const __grid_constant__ CUtensorMap tmH_hi, const __grid_constant__ CUtensorMap tmH_lo, KArgs p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t a_full, full_bar[MAXSLOTS], empty_bar[MAXSLOTS], tfull_bar;
  __shared__ uint32_t tmem_slot;
  const RecTcFwdArgs& a = p.a;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x;
  const int H = p.H, B = p.B, Bpad = p.Bpad, U = p.U, T = a.T, nkb = p.nkb, ngroups = p.ngroups;
  unsigned char* sA = smem;                                         // [2 planes][nkb][a_kb_bytes]
  unsigned char* sRing = smem + 2 * (size_t)nkb * p.a_kb_bytes;     // [slots][GKB][2 planes][Bpad*128]
  const uint32_t plane_b_bytes = (uint32_t)Bpad * 128;
  const uint32_t slot_bytes = GKB * p.kb_bytes;
  const bool full_flight = p.slots >= ngroups;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < Bpad) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    tc::mbar_init(&a_full, 1);
    for (int s = 0; s < p.slots; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    tc::mbar_init(&tfull_bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 8) tc::tmem_alloc(&tmem_slot, tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    // (the whole warp runs this converged; one elected lane issues each TMA / arrive)
    if (lane == 0) {
      tc::tma_prefetch_desc(&tmW_hi); tc::tma_prefetch_desc(&tmW_lo);
      tc::tma_prefetch_desc(&tmH_hi); tc::tma_prefetch_desc(&tmH_lo);
    }
    __syncwarp();
    tc::mbar_arrive_expect_tx_warp(&a_full, 2u * (uint32_t)nkb * p.a_kb_bytes);
    for (int kb = 0; kb < nkb; ++kb) {
      tc::tma_load_2d_warp(sA + (size_t)kb * p.a_kb_bytes, &tmW_hi, kb * 64, j * 4 * U, &a_full);
      tc::tma_load_2d_warp(sA + (size_t)(nkb + kb) * p.a_kb_bytes, &tmW_lo, kb * 64, j * 4 * U, &a_full);
    }
    const unsigned nctas = gridDim.x;
    uint32_t git = 0;
    for (int t = 0; t < T; ++t) {
      if (t > 0) {
        if (p.variant & 1) wait_flags_warp(a.barrier, nctas, (unsigned)t, lane, a.dbg, t);
        else { while (ld_acquire_u32(a.barrier) < nctas * (unsigned)t) {} __syncwarp(); }
        tc::fence_proxy_async_all();     // other CTAs' generic-proxy stores of h_{t-1} -> TMA reads
      }
      if (lane == 0) RS_STAMP(a.dbg, t, 0);
      __syncwarp();
      for (int grp = 0; grp < ngroups; ++grp, ++git) {
        const int s = git % p.slots;
        const uint32_t ph = (git / p.slots) & 1;
        // with the whole of h in flight a slot is reused one step later, after the grid barrier,
        // i.e. after this CTA's epilogue consumed the accumulator: no empty barrier needed
        if (!full_flight) tc::mbar_wait(&empty_bar[s], ph ^ 1);
        const int kb0 = grp * GKB, kbn = min(GKB, nkb - kb0);
        unsigned char* st = sRing + (size_t)s * slot_bytes;
        tc::mbar_arrive_expect_tx_warp(&full_bar[s], (uint32_t)kbn * p.kb_bytes);
        for (int i = 0; i < kbn; ++i) {
          unsigned char* dst = st + (size_t)i * p.kb_bytes;
          tc::tma_load_2d_warp(dst, &tmH_hi, (kb0 + i) * 64, t * B, &full_bar[s]);               // slot t = h_{t-1}
          tc::tma_load_2d_warp(dst + plane_b_bytes, &tmH_lo, (kb0 + i) * 64, t * B, &full_bar[s]);
        }
      }
      if (lane == 0) RS_STAMP(a.dbg, t, 1);
      __syncwarp();
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer
    // (converged warp, elect.sync inside each issue: see tc_common.cuh)
    {
      const uint32_t idesc = tc::instr_desc_bf16(64, Bpad);
      tc::mbar_wait(&a_full, 0);
      tc::tc_fence_after();
      const uint64_t dw_hi0 = tc::smem_desc_sw128(tc::smem_u32(sA));
      const uint64_t dw_lo0 = tc::smem_desc_sw128(tc::smem_u32(sA + (size_t)nkb * p.a_kb_bytes));
      const uint64_t dring0 = tc::smem_desc_sw128(tc::smem_u32(sRing));
      const uint64_t a_kb_u = p.a_kb_bytes >> 4, kb_u = p.kb_bytes >> 4, pl_u = plane_b_bytes >> 4;
      uint32_t git = 0;
      for (int t = 0; t < T; ++t) {
        for (int grp = 0; grp < ngroups; ++grp, ++git) {
          const int s = git % p.slots;
          const uint32_t ph = (git / p.slots) & 1;
          tc::mbar_wait(&full_bar[s], ph);
          tc::tc_fence_after();
          if (lane == 0 && grp == 0) RS_STAMP(a.dbg, t, 2);
          if (lane == 0 && grp < 4) RS_STAMP(a.dbg, t, 8 + grp);
          __syncwarp();
          const int kb0 = grp * GKB, kbn = min(GKB, nkb - kb0);
          for (int i = 0; i < kbn; ++i) {
            const uint64_t dwh = dw_hi0 + (uint64_t)(kb0 + i) * a_kb_u;
            const uint64_t dwl = dw_lo0 + (uint64_t)(kb0 + i) * a_kb_u;
            const uint64_t dhh = dring0 + (uint64_t)s * (slot_bytes >> 4) + (uint64_t)i * kb_u;
            const uint64_t dhl = dhh + pl_u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              tc::mma_bf16_ss_warp(tmem, dwh + 2 * k, dhh + 2 * k, idesc, (uint32_t)((kb0 | i | k) != 0));
              tc::mma_bf16_ss_warp(tmem, dwh + 2 * k, dhl + 2 * k, idesc, 1u);
              tc::mma_bf16_ss_warp(tmem, dwl + 2 * k, dhh + 2 * k, idesc, 1u);
            }
          }
          if (lane == 0 && grp < 4) RS_STAMP(a.dbg, t, 12 + grp);
          __syncwarp();
          if (!full_flight) tc::mma_commit_warp(&empty_bar[s]);
        }
        tc::mma_commit_warp(&tfull_bar);
        if (lane == 0) RS_STAMP(a.dbg, t, 3);
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;                     // TMEM sub-partition: accumulator rows 16q .. 16q+15 in lanes 0..15
    const int hf = warp >> 2;                   // which 16-column groups: hf, hf + 2
    const int m = q * 16 + (lane & 15);         // accumulator row
    const bool row_ok = lane < 16 && m < 4 * U;
    const int g = lane & 3;                     // gate of this row: 0 i, 1 j, 2 f, 3 o
    const int unit = j * U + (m >> 2);
    const int ng = Bpad / 16;
    float c[MAXG][4], hl[MAXG][4];
    float act_keep[MAXG][16], cn_keep[MAXG][4];
    int lenr[MAXG][4];
#pragma unroll
    for (int gl = 0; gl < MAXG; ++gl)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int b = (hf + 2 * gl) * 16 + 4 * g + k;
        const bool ok = row_ok && (hf + 2 * gl) < ng && b < B;
        c[gl][k] = ok ? a.c0[(size_t)b * H + unit] : 0.f;
        hl[gl][k] = ok ? a.h0[(size_t)b * H + unit] : 0.f;
        lenr[gl][k] = ok ? a.len[b] : 0;
      }
    const float fbias = (g == 2) ? 1.0f : 0.0f;          // forget_bias
    const float pre = (g == 1) ? 2.0f : 1.0f;            // tanh(x) = 2*sigmoid(2x) - 1
    const float post_m = (g == 1) ? 2.0f : 1.0f, post_a = (g == 1) ? -1.0f : 0.0f;

    for (int t = 0; t < T; ++t) {
      // hoisted input projection for this step (independent of the recurrence: issue early)
      float4 gxv[MAXG][4];
#pragma unroll
      for (int gl = 0; gl < MAXG; ++gl) {
        const int gi = hf + 2 * gl;
        if (row_ok && gi < ng) {
          const float4* gp = reinterpret_cast<const float4*>(
              a.gx + (((size_t)t * p.nslice + j) * (4 * U) + m) * Bpad + gi * 16);
#pragma unroll
          for (int i = 0; i < 4; ++i) gxv[gl][i] = __ldg(gp + i);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) gxv[gl][i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      tc::mbar_wait(&tfull_bar, (uint32_t)(t & 1));
      tc::tc_fence_after();
      if (threadIdx.x == 0) RS_STAMP(a.dbg, t, 4);
#pragma unroll
      for (int gl = 0; gl < MAXG; ++gl) {
        const int gi = hf + 2 * gl;
        if (gi < ng) {                           // warp-uniform
          float v[16];
          tc::tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(gi * 16), v);
          tc::tmem_ld_wait();
          float act[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 x = gxv[gl][i];
            const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float z = (v[4 * i + e] + xs[e] + fbias) * pre;
              act[4 * i + e] = fmaf(fast_sigmoid(z), post_m, post_a);
            }
          }
          // (the activated gates are kept for backward; they are stored AFTER h_t is published,
          //  below, so that they do not sit in front of the barrier's store drain)
#pragma unroll
          for (int i = 0; i < 16; ++i) act_keep[gl][i] = act[i];
          // quad exchange: this lane owns cells b = gi*16 + 4g + k; gate tau comes from lane g ^ (g^tau)
          float own[4], rcv[3][4];
#pragma unroll
          for (int k = 0; k < 4; ++k) own[k] = pick4(g, act[k], act[4 + k], act[8 + k], act[12 + k]);
#pragma unroll
          for (int d = 1; d < 4; ++d)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float snd = pick4(g ^ d, act[k], act[4 + k], act[8 + k], act[12 + k]);
              rcv[d - 1][k] = __shfl_xor_sync(0xffffffffu, snd, d);
            }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // value of gate tau arrives from xor-distance d = g ^ tau (d = 0: own)
            const float ig = pick4(g ^ 0, own[k], rcv[0][k], rcv[1][k], rcv[2][k]);
            const float jg = pick4(g ^ 1, own[k], rcv[0][k], rcv[1][k], rcv[2][k]);
            const float fg = pick4(g ^ 2, own[k], rcv[0][k], rcv[1][k], rcv[2][k]);
            const float og = pick4(g ^ 3, own[k], rcv[0][k], rcv[1][k], rcv[2][k]);
            const int b = gi * 16 + 4 * g + k;
            const float c_new = c[gl][k] * fg + ig * jg;
            const float h_new = fast_tanh(c_new) * og;
            const bool valid = t < lenr[gl][k];
            if (valid) { c[gl][k] = c_new; hl[gl][k] = h_new; }
            if (row_ok && b < B) {
              __nv_bfloat16 hh, hlo;
              tc::split_bf16(valid ? h_new : 0.f, hh, hlo);
              const size_t o = ((size_t)(t + 1) * B + b) * H + unit;
              a.h_hi[o] = hh;
              a.h_lo[o] = hlo;
            }
            cn_keep[gl][k] = c_new;
          }
        }
      }
      if (p.variant & 2) {
      if (a.gates && row_ok) {
#pragma unroll
        for (int gl = 0; gl < MAXG; ++gl) {
          const int gi = hf + 2 * gl;
          if (gi < ng) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int b = gi * 16 + i;
              if (b < B) a.gates[((size_t)t * B + b) * 4 * H + (size_t)g * H + unit] = act_keep[gl][i];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int b = gi * 16 + 4 * g + k;
              if (b < B) a.cs[((size_t)t * B + b) * H + unit] = cn_keep[gl][k];
            }
          }
        }
      }
      }
      // publish: generic stores -> async proxy (per-thread proxy fence); gpu-scope visibility:
      // the CTA barrier followed by ONE release-add is cumulative over every thread's stores
      if (threadIdx.x == 0) RS_STAMP(a.dbg, t, 5);
      tc::tc_fence_before();
      tc::fence_proxy_async_all();
      epi_bar_sync();
      if (threadIdx.x == 0) RS_STAMP(a.dbg, t, 6);
      if (threadIdx.x == 0 && t + 1 < T) {
        if (p.variant & 1) st_release_u32(a.barrier + j, (unsigned)(t + 1));
        else red_release_add(a.barrier, 1u);
      }
      if (threadIdx.x == 0) RS_STAMP(a.dbg, t, 7);
      // reserve for backward (not on the critical path of the recurrence)
      if (!(p.variant & 2)) {
      if (a.gates && row_ok) {
#pragma unroll
        for (int gl = 0; gl < MAXG; ++gl) {
          const int gi = hf + 2 * gl;
          if (gi < ng) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int b = gi * 16 + i;
              if (b < B) a.gates[((size_t)t * B + b) * 4 * H + (size_t)g * H + unit] = act_keep[gl][i];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int b = gi * 16 + 4 * g + k;
              if (b < B) a.cs[((size_t)t * B + b) * H + unit] = cn_keep[gl][k];
            }
          }
        }
      }
      }
    }
    if (row_ok) {
#pragma unroll
      for (int gl = 0; gl < MAXG; ++gl)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int b = (hf + 2 * gl) * 16 + 4 * g + k;
          if ((hf + 2 * gl) < ng && b < B) {
            if (a.cT) a.cT[(size_t)b * H + unit] = c[gl][k];
            if (a.hT) a.hT[(size_t)b * H + unit] = hl[gl][k];
          }
        }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem, tmem_cols);
}

// ------------------------------------------------------------------------------------
// Backward recurrent kernel (see the file header).
// ------------------------------------------------------------------------------------
constexpr int BW_U = 16;
constexpr int BW_GKB = 8;        // K-blocks per mbarrier in the backward ring

struct KBwdArgs {
  RecTcBwdArgs a;
  int H, B, Bpad, slots, nkb, ngroups;
  int variant;
  uint32_t kb_bytes;             // Bpad rows x 128 B
};

__global__ void __launch_bounds__(NTHREADS, 1)
rec_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmG, KBwdArgs p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t w_full, full_bar[MAXSLOTS], empty_bar[MAXSLOTS], tfull_bar;
  __shared__ uint32_t tmem_slot;
  const RecTcBwdArgs& a = p.a;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x;
  const int H = p.H, B = p.B, T = a.T, nkb = p.nkb, ngroups = p.ngroups, G = 4 * p.H;
  constexpr uint32_t W_KB_BYTES = BW_U * 128;                       // 2 KB per K-block
  unsigned char* sW = smem;                                          // [nkb][16 rows x 128 B]
  unsigned char* sRing = smem + (size_t)nkb * W_KB_BYTES;            // [slots][GKB][Bpad x 128 B] (+ overhang pad)
  const uint32_t slot_bytes = BW_GKB * p.kb_bytes;

  if (threadIdx.x == 0) {
    tc::mbar_init(&w_full, 1);
    for (int s = 0; s < p.slots; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    tc::mbar_init(&tfull_bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 8) tc::tmem_alloc(&tmem_slot, 32);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 8) {
    if (lane == 0) { tc::tma_prefetch_desc(&tmW); tc::tma_prefetch_desc(&tmG); }
    __syncwarp();
    tc::mbar_arrive_expect_tx_warp(&w_full, (uint32_t)nkb * W_KB_BYTES);
    for (int kb = 0; kb < nkb; ++kb) tc::tma_load_2d_warp(sW + (size_t)kb * W_KB_BYTES, &tmW, kb * 64, j * BW_U, &w_full);
    const unsigned nctas = gridDim.x;
    uint32_t git = 0;
    unsigned epoch = 0;
    for (int t = T - 1; t >= 1; --t) {        // dh_{t-1} from dgates_t
      ++epoch;
      if (p.variant & 1) wait_flags_warp(a.barrier, nctas, epoch, lane, a.dbg, t);
      else { while (ld_acquire_u32(a.barrier) < nctas * epoch) {} __syncwarp(); }
      tc::fence_proxy_async_all();
      if (lane == 0) RS_STAMP(a.dbg, t, 0);
      __syncwarp();
      for (int grp = 0; grp < ngroups; ++grp, ++git) {
        const int s = git % p.slots;
        const uint32_t ph = (git / p.slots) & 1;
        tc::mbar_wait(&empty_bar[s], ph ^ 1);
        const int kb0 = grp * BW_GKB, kbn = min(BW_GKB, nkb - kb0);
        tc::mbar_arrive_expect_tx_warp(&full_bar[s], (uint32_t)kbn * p.kb_bytes);
        for (int i = 0; i < kbn; ++i)
          tc::tma_load_2d_warp(sRing + (size_t)s * slot_bytes + (size_t)i * p.kb_bytes, &tmG, (kb0 + i) * 64, t * B, &full_bar[s]);
      }
      if (lane == 0) RS_STAMP(a.dbg, t, 1);
      __syncwarp();
    }
  } else if (warp == 9) {
    {
      const uint32_t idesc = tc::instr_desc_bf16(64, BW_U);
      tc::mbar_wait(&w_full, 0);
      tc::tc_fence_after();
      const uint64_t dw0 = tc::smem_desc_sw128(tc::smem_u32(sW));
      const uint64_t dring0 = tc::smem_desc_sw128(tc::smem_u32(sRing));
      const uint64_t kb_u = p.kb_bytes >> 4;
      uint32_t git = 0;
      for (int t = T - 1; t >= 1; --t) {
        for (int grp = 0; grp < ngroups; ++grp, ++git) {
          const int s = git % p.slots;
          const uint32_t ph = (git / p.slots) & 1;
          tc::mbar_wait(&full_bar[s], ph);
          tc::tc_fence_after();
          if (lane == 0 && grp == 0) RS_STAMP(a.dbg, t, 2);
          __syncwarp();
          const int kb0 = grp * BW_GKB, kbn = min(BW_GKB, nkb - kb0);
          for (int i = 0; i < kbn; ++i) {
            const uint64_t dg = dring0 + (uint64_t)s * (slot_bytes >> 4) + (uint64_t)i * kb_u;
            const uint64_t dw = dw0 + (uint64_t)(kb0 + i) * (W_KB_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) tc::mma_bf16_ss_warp(tmem, dg + 2 * k, dw + 2 * k, idesc, (uint32_t)((kb0 | i | k) != 0));
          }
          tc::mma_commit_warp(&empty_bar[s]);
        }
        tc::mma_commit_warp(&tfull_bar);
        if (lane == 0) RS_STAMP(a.dbg, t, 3);
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;                     // TMEM sub-partition: batch rows 16q .. 16q+15 in lanes 0..15
    const int hf = warp >> 2;                   // units 8hf .. 8hf+7 of the slice
    const int b = q * 16 + (lane & 15);
    const bool ok = lane < 16 && b < B;
    const int unit0 = j * BW_U + hf * 8;
    const int len_b = ok ? a.len[b] : 0;
    float dc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) dc[u] = 0.f;
    uint32_t n = 0;
    for (int t = T - 1; t >= 0; --t, ++n) {
      // operands of the cell backward (independent of the recurrence: issue before waiting)
      float4 gi[2], gj[2], gf[2], go[2], ct[2], cp[2], dy[2];
      if (ok) {
        const size_t row = (size_t)t * B + b;
        const float4* gp = reinterpret_cast<const float4*>(a.gates + row * G + unit0);
        const float4* cpp = reinterpret_cast<const float4*>(a.cs + row * H + unit0);
        const float4* cpv = (t > 0) ? reinterpret_cast<const float4*>(a.cs + (row - B) * H + unit0)
                                    : reinterpret_cast<const float4*>(a.c0 + (size_t)b * H + unit0);
        const float4* dp = reinterpret_cast<const float4*>(a.dout + row * H + unit0);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          gi[i] = __ldg(gp + i); gj[i] = __ldg(gp + H / 4 + i);
          gf[i] = __ldg(gp + 2 * (H / 4) + i); go[i] = __ldg(gp + 3 * (H / 4) + i);
          ct[i] = __ldg(cpp + i); cp[i] = __ldg(cpv + i); dy[i] = __ldg(dp + i);
        }
      }
      float dh[8];
      if (n > 0) {
        tc::mbar_wait(&tfull_bar, (n - 1) & 1);
        tc::tc_fence_after();
        tc::tmem_ld8(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * 8), dh);
        tc::tmem_ld_wait();
        if (threadIdx.x == 0) RS_STAMP(a.dbg, t, 4);
      } else {
#pragma unroll
        for (int u = 0; u < 8; ++u) dh[u] = 0.f;
      }
      if (ok) {
        const bool valid = t < len_b;
        const float* pi = reinterpret_cast<const float*>(gi);
        const float* pj = reinterpret_cast<const float*>(gj);
        const float* pf = reinterpret_cast<const float*>(gf);
        const float* po = reinterpret_cast<const float*>(go);
        const float* pct = reinterpret_cast<const float*>(ct);
        const float* pcp = reinterpret_cast<const float*>(cp);
        const float* pdy = reinterpret_cast<const float*>(dy);
        float d4[4][8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float di = 0.f, dj = 0.f, df = 0.f, dob = 0.f;
          if (valid) {
            const float ig = pi[u], jg = pj[u], fg = pf[u], og = po[u];
            const float dh_tot = dh[u] + pdy[u];
            const float tch = fast_tanh(pct[u]);
            dob = dh_tot * tch * og * (1.f - og);
            const float dc_tot = dc[u] + dh_tot * og * (1.f - tch * tch);
            di = dc_tot * jg * ig * (1.f - ig);
            dj = dc_tot * ig * (1.f - jg * jg);
            df = dc_tot * pcp[u] * fg * (1.f - fg);
            dc[u] = dc_tot * fg;
          }
          d4[0][u] = di; d4[1][u] = dj; d4[2][u] = df; d4[3][u] = dob;
        }
        const size_t row = (size_t)t * B + b;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t h[4], l[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            __nv_bfloat16 h0, l0, h1, l1;
            tc::split_bf16(d4[g][2 * i], h0, l0);
            tc::split_bf16(d4[g][2 * i + 1], h1, l1);
            h[i] = tc::pack_bf16(h0, h1);
            l[i] = tc::pack_bf16(l0, l1);
          }
          const size_t o = row * G + (size_t)g * H + unit0;
          *reinterpret_cast<uint4*>(a.dg_hi + o) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(a.dg_lo + o) = make_uint4(l[0], l[1], l[2], l[3]);
        }
      }
      if (threadIdx.x == 0) RS_STAMP(a.dbg, t, 5);
      tc::tc_fence_before();
      tc::fence_proxy_async_all();
      epi_bar_sync();
      if (threadIdx.x == 0) RS_STAMP(a.dbg, t, 6);
      if (threadIdx.x == 0 && t > 0) {
        if (p.variant & 1) st_release_u32(a.barrier + j, n + 1);
        else red_release_add(a.barrier, 1u);
      }
      if (threadIdx.x == 0) RS_STAMP(a.dbg, t, 7);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem, 32);
}

// wrec[(unit/U)*4U + (unit%U)*4 + g][k] = W[k][g*H + unit]; W = one half of the TF kernel (row-major [H,4H])
__global__ void pack_wrec_kernel(const float* __restrict__ Wh, int H, int U, __nv_bfloat16* __restrict__ hi,
                                 __nv_bfloat16* __restrict__ lo) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;      // n = column of Wh in [0, 4H)
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < H && n < 4 * H) ? Wh[(size_t)k * 4 * H + n] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    if (n < 4 * H && k < H) {
      const int g = n / H, unit = n - g * H;
      const size_t row = (size_t)(unit / U) * 4 * U + (size_t)(unit % U) * 4 + g;
      __nv_bfloat16 h, l;
      tc::split_bf16(tile[threadIdx.x][i], h, l);
      hi[row * H + k] = h;
      lo[row * H + k] = l;
    }
  }
}

}  // namespace

bool rec_tc_geometry(int H, int B, RecTcGeom* g) {
  if (H % 64 != 0 || H < 64 || B < 1 || B > 64) return false;
  const int Bpad = (B + 15) / 16 * 16;
  const size_t budget = 227 * 1024 - 2048;            // dynamic smem minus alignment slack / static barriers
  const int nkb = H / 64, ngroups = (nkb + GKB - 1) / GKB;
  const int nsm = sm_count();
  // The step time is bounded by how much of h_{t-1} can be in flight at once, so take the
  // slice width U whose resident weights leave room for the deepest ring; ties go to the
  // wider slice (fewer CTAs re-reading h from L2).
  bool found = false;
  double best = -1.0;
  for (int U = 16; U >= 8; U /= 2) {
    if (H % U != 0 || H / U > nsm) continue;
    const size_t a_bytes = 2 * (size_t)nkb * (4 * U) * 128;
    const size_t slot = (size_t)GKB * 2 * Bpad * 128;
    if (a_bytes + slot > budget) continue;
    int slots = (int)((budget - a_bytes) / slot);
    if (slots > MAXSLOTS) slots = MAXSLOTS;
    if (slots > ngroups) slots = ngroups;
    if (slots < 1) continue;
    // the M = 64 instruction reads 64 rows from each K-block base: what lies behind the resident
    // operand (the ring) must cover that overhang
    const size_t overhang = 64 * 128 - (size_t)(4 * U) * 128;
    if ((size_t)slots * slot < overhang) continue;
    const double score = (double)slots / ngroups;
    if (score > best + 1e-9) {
      best = score; found = true;
      g->H = H; g->B = B; g->Bpad = Bpad; g->U = U; g->nslice = H / U; g->stages = slots;
      g->ts = 0; g->nkb_t = 0; g->ngl = 0; g->gkb = 0;
      g->smem_bytes = a_bytes + (size_t)slots * slot + 1024;
    }
  }
  return found;
}

int pack_wrec(const float* W, int H, int U, __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t st) {
  pack_wrec_kernel<<<dim3(cdiv(4 * H, 32), cdiv(H, 32)), dim3(32, 8), 0, st>>>(W, H, U, hi, lo);
  RS_CHECK_LAUNCH();
  return RS_OK;
}

template <typename K>
static int coop_launch(K kernel, int grid, size_t smem, void** kargs, cudaStream_t st, const char* name) {
  int dev = 0, per_sm = 0, nsm = 0;
  RS_CHECK_CUDA(cudaGetDevice(&dev));
  RS_CHECK_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  RS_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RS_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, NTHREADS, smem));
  RS_REQUIRE(per_sm * nsm >= grid, RS_ERR_UNSUPPORTED, "%s: %d CTAs cannot be co-resident", name, grid);
  RS_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)kernel, dim3(grid), dim3(NTHREADS), kargs, smem, st));
  count_launch();
  return RS_OK;
}

int lstm_rec_tc_forward(const RecTcGeom& g, const RecTcFwdArgs& a, cudaStream_t st) {
  RS_REQUIRE(a.T > 0, RS_ERR_INVALID, "lstm_rec_tc_forward: T=%d", a.T);
  CUtensorMap tw_hi, tw_lo, th_hi, th_lo;
  int rc;
  if ((rc = tmap_2d_bf16(&tw_hi, a.wrec_hi, 4 * g.H, g.H, g.H, 4 * g.U)) != RS_OK) return rc;
  if ((rc = tmap_2d_bf16(&tw_lo, a.wrec_lo, 4 * g.H, g.H, g.H, 4 * g.U)) != RS_OK) return rc;
  const int hrows = (a.T + 1) * g.B;
  if ((rc = tmap_2d_bf16(&th_hi, a.h_hi, hrows, g.H, g.H, g.Bpad)) != RS_OK) return rc;
  if ((rc = tmap_2d_bf16(&th_lo, a.h_lo, hrows, g.H, g.H, g.Bpad)) != RS_OK) return rc;
  KArgs p;
  p.a = a;
  p.H = g.H; p.B = g.B; p.Bpad = g.Bpad; p.U = g.U; p.nslice = g.nslice; p.slots = g.stages; p.nkb = g.H / 64;
  p.ngroups = (p.nkb + GKB - 1) / GKB;
  p.a_kb_bytes = (uint32_t)(4 * g.U) * 128;
  p.kb_bytes = 2u * (uint32_t)g.Bpad * 128;
  { const char* v = getenv("RS_REC_VARIANT"); p.variant = v ? atoi(v) : 0; }
  RS_CHECK_CUDA(cudaMemsetAsync(a.barrier, 0, 256 * sizeof(unsigned), st));
  void* kargs[] = {(void*)&tw_hi, (void*)&tw_lo, (void*)&th_hi, (void*)&th_lo, (void*)&p};
  return coop_launch(rec_tc_fwd_kernel, g.nslice, g.smem_bytes, kargs, st, "lstm_rec_tc_forward");
}

bool rec_tc_bwd_geometry(int H, int B, RecTcBwdGeom* g) {
  if (H % 64 != 0 || H < 64 || B < 1 || B > 64) return false;
  if (H / BW_U > sm_count()) return false;
  const int Bpad = (B + 15) / 16 * 16;
  const size_t budget = 227 * 1024 - 2048;
  const int nkb = 4 * H / 64, ngroups = (nkb + BW_GKB - 1) / BW_GKB;
  const size_t w_bytes = (size_t)nkb * BW_U * 128;
  const size_t slot = (size_t)BW_GKB * Bpad * 128;
  const size_t overhang = 64 * 128 - (size_t)Bpad * 128;      // M = 64 reads 64 rows from each K-block base
  if (w_bytes + slot + overhang > budget) return false;
  int slots = (int)((budget - w_bytes - overhang) / slot);
  if (slots > MAXSLOTS) slots = MAXSLOTS;
  if (slots > ngroups) slots = ngroups;
  g->H = H; g->B = B; g->Bpad = Bpad; g->nslice = H / BW_U; g->stages = slots;
  g->smem_bytes = w_bytes + (size_t)slots * slot + overhang + 1024;
  return true;
}

int lstm_rec_tc_backward(const RecTcBwdGeom& g, const RecTcBwdArgs& a, cudaStream_t st) {
  RS_REQUIRE(a.T > 0, RS_ERR_INVALID, "lstm_rec_tc_backward: T=%d", a.T);
  CUtensorMap tw, tg;
  int rc;
  if ((rc = tmap_2d_bf16(&tw, a.wh_hi, g.H, 4 * g.H, 4 * g.H, BW_U)) != RS_OK) return rc;
  if ((rc = tmap_2d_bf16(&tg, a.dg_hi, a.T * g.B, 4 * g.H, 4 * g.H, g.Bpad)) != RS_OK) return rc;
  KBwdArgs p;
  p.a = a;
  p.H = g.H; p.B = g.B; p.Bpad = g.Bpad; p.slots = g.stages; p.nkb = 4 * g.H / 64;
  p.ngroups = (p.nkb + BW_GKB - 1) / BW_GKB;
  p.kb_bytes = (uint32_t)g.Bpad * 128;
  { const char* v = getenv("RS_REC_VARIANT"); p.variant = v ? atoi(v) : 0; }
  RS_CHECK_CUDA(cudaMemsetAsync(a.barrier, 0, 256 * sizeof(unsigned), st));
  void* kargs[] = {(void*)&tw, (void*)&tg, (void*)&p};
  return coop_launch(rec_tc_bwd_kernel, g.nslice, g.smem_bytes, kargs, st, "lstm_rec_tc_backward");
}

}  // namespace rs
