// Tensor-core persistent recurrent LSTM kernel (forward), sm_100a.
//
// Replaces the tf.nn.dynamic_rnn loop over BasicLSTMCell
// (/root/reference/models/AcousticModel.py:227-237, :277-278) for one layer; same
// semantics as lstm_rec.cu / oracle/model.py.
//
// Ownership: CTA j owns hidden units [jU, jU+U) (U = 16 or 8).  Its 4U gate rows of
// Wh^T (bf16 hi + lo planes, K-major, SWIZZLE_128B) are loaded ONCE by TMA and stay in
// shared memory for all T steps; the cell state stays in registers.  Per step:
//   producer thread : waits on the grid barrier (all CTAs published h_{t-1}), then streams
//                     h_{t-1} (hi + lo planes, [Bpad x H]) through a TMA ring, 64 columns a stage
//   MMA thread      : D[128 x Bpad] (TMEM, fp32) = W_hi h_hi + W_hi h_lo + W_lo h_hi, tcgen05.mma
//                     M = 128 (rows >= 4U are don't-care), N = Bpad, K = 16 per instruction
//   epilogue warps  : tcgen05.ld the accumulator, add the hoisted input projection gx, apply
//                     the gate non-linearities (row r = 4*unit + gate, so a lane quad holds
//                     i,j,f,o of one unit), exchange inside the quad by shuffles, update c,
//                     publish h_t as bf16 hi/lo planes, arrive on the grid barrier.
#include "lstm_rec_tc.cuh"
#include "tc_common.cuh"
#include <cuda.h>

namespace rs {
namespace {

constexpr int NTHREADS = 192;   // warps 0,1,4,5: epilogue; warp 2: TMA producer; warp 3: MMA issuer
constexpr int MAXG = 2;         // 16-column groups per epilogue thread (Bpad <= 64, two column halves)
constexpr int MAXSTAGES = 8;

struct KArgs {
  RecTcFwdArgs a;
  int H, B, Bpad, U, nslice, stages, nkb;
  uint32_t a_kb_bytes;          // one K-block of the resident operand: 4U rows x 128 B
  uint32_t stage_bytes;         // one ring stage: 2 planes x Bpad rows x 128 B
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
#define RS_STAMP(dbgp, step, ev) do { if ((dbgp) && blockIdx.x == 0) (dbgp)[(size_t)(step) * 8 + (ev)] = gtime(); } while (0)
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

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
  __shared__ uint64_t a_full, full_bar[MAXSTAGES], empty_bar[MAXSTAGES], tfull_bar;
  __shared__ uint32_t tmem_slot;
  const RecTcFwdArgs& a = p.a;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x;
  const int H = p.H, B = p.B, Bpad = p.Bpad, U = p.U, T = a.T, nkb = p.nkb;
  unsigned char* sA = smem;                                         // [2 planes][nkb][a_kb_bytes]
  unsigned char* sRing = smem + 2 * (size_t)nkb * p.a_kb_bytes;     // [stages][2 planes][Bpad*128]
  const uint32_t plane_b_bytes = (uint32_t)Bpad * 128;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < Bpad) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    tc::mbar_init(&a_full, 1);
    for (int s = 0; s < p.stages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    tc::mbar_init(&tfull_bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 2) tc::tmem_alloc(&tmem_slot, tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 2) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      tc::tma_prefetch_desc(&tmW_hi); tc::tma_prefetch_desc(&tmW_lo);
      tc::tma_prefetch_desc(&tmH_hi); tc::tma_prefetch_desc(&tmH_lo);
      tc::mbar_arrive_expect_tx(&a_full, 2u * (uint32_t)nkb * p.a_kb_bytes);
      for (int kb = 0; kb < nkb; ++kb) {
        tc::tma_load_2d(sA + (size_t)kb * p.a_kb_bytes, &tmW_hi, kb * 64, j * 4 * U, &a_full);
        tc::tma_load_2d(sA + (size_t)(nkb + kb) * p.a_kb_bytes, &tmW_lo, kb * 64, j * 4 * U, &a_full);
      }
      const unsigned nctas = gridDim.x;
      uint32_t it = 0;
      for (int t = 0; t < T; ++t) {
        if (t > 0) {
          while (ld_acquire_u32(a.barrier) < nctas * (unsigned)t) {}
          tc::fence_proxy_async_all();     // other CTAs' generic-proxy stores of h_{t-1} -> TMA reads
        }
        RS_STAMP(a.dbg, t, 0);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          tc::mbar_wait(&empty_bar[s], ph ^ 1);
          unsigned char* st = sRing + (size_t)s * p.stage_bytes;
          tc::mbar_arrive_expect_tx(&full_bar[s], 2 * plane_b_bytes);
          tc::tma_load_2d(st, &tmH_hi, kb * 64, t * B, &full_bar[s]);                  // slot t = h_{t-1}
          tc::tma_load_2d(st + plane_b_bytes, &tmH_lo, kb * 64, t * B, &full_bar[s]);
        }
        RS_STAMP(a.dbg, t, 1);
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = tc::instr_desc_bf16(128, Bpad);
      tc::mbar_wait(&a_full, 0);
      tc::tc_fence_after();
      const uint32_t sa = tc::smem_u32(sA);
      uint32_t it = 0;
      for (int t = 0; t < T; ++t) {
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          tc::mbar_wait(&full_bar[s], ph);
          tc::tc_fence_after();
          if (kb == 0) RS_STAMP(a.dbg, t, 2);
          const uint64_t dwh = tc::smem_desc_sw128(sa + (uint32_t)kb * p.a_kb_bytes);
          const uint64_t dwl = tc::smem_desc_sw128(sa + (uint32_t)(nkb + kb) * p.a_kb_bytes);
          const uint32_t sb = tc::smem_u32(sRing + (size_t)s * p.stage_bytes);
          const uint64_t dhh = tc::smem_desc_sw128(sb), dhl = tc::smem_desc_sw128(sb + plane_b_bytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            tc::mma_bf16_ss(tmem, dwh + 2 * k, dhh + 2 * k, idesc, (kb | k) != 0);
            tc::mma_bf16_ss(tmem, dwh + 2 * k, dhl + 2 * k, idesc, true);
            tc::mma_bf16_ss(tmem, dwl + 2 * k, dhh + 2 * k, idesc, true);
          }
          tc::mma_commit(&empty_bar[s]);
        }
        tc::mma_commit(&tfull_bar);
        RS_STAMP(a.dbg, t, 3);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;                     // TMEM lane quadrant (rows 32q .. 32q+31)
    const int hf = warp >> 2;                   // which 16-column groups: hf, hf + 2
    const int r = q * 32 + lane;                // accumulator row
    const bool row_ok = r < 4 * U;
    const int g = lane & 3;                     // gate of this row: 0 i, 1 j, 2 f, 3 o
    const int unit = j * U + (r >> 2);
    const int ng = Bpad / 16;
    float c[MAXG][4], hl[MAXG][4];
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
              a.gx + (((size_t)t * p.nslice + j) * (4 * U) + r) * Bpad + gi * 16);
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
          if (a.gates && row_ok) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int b = gi * 16 + i;
              if (b < B) a.gates[((size_t)t * B + b) * 4 * H + (size_t)g * H + unit] = act[i];
            }
          }
          // quad exchange: this lane owns cells b = gi*16 + 4g + k; gate tau comes from lane g^... = tau
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
              if (a.cs) a.cs[((size_t)t * B + b) * H + unit] = c_new;
            }
          }
        }
      }
      // publish: every epilogue thread's stores -> gpu scope -> async proxy of other SMs
      if (threadIdx.x == 0) RS_STAMP(a.dbg, t, 5);
      tc::tc_fence_before();
      __threadfence();
      tc::fence_proxy_async_all();
      epi_bar_sync();
      if (threadIdx.x == 0) RS_STAMP(a.dbg, t, 6);
      if (threadIdx.x == 0 && t + 1 < T) red_release_add(a.barrier, 1u);
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
  if (warp == 2) tc::tmem_dealloc(tmem, tmem_cols);
}


// ------------------------------------------------------------------------------------
// Backward recurrent kernel.  CTA j owns hidden units [16j, 16j+16): its 16 rows of Wh
// ([16][4H] bf16, K-major) are resident in shared memory (the B operand, N = 16); per
// step the freshly published dgates_t ([Bpad x 4H] bf16) stream through a TMA ring as
// the A operand (M = 128, rows >= Bpad are don't-care) and D[b][u] = dh_{t-1} lands in
// TMEM lanes 0..B-1.  Epilogue thread (b, 8 units) does the cell backward for its units.
// ------------------------------------------------------------------------------------
constexpr int BW_U = 16;

struct KBwdArgs {
  RecTcBwdArgs a;
  int H, B, Bpad, stages, nkb;
  uint32_t stage_bytes;          // Bpad rows x 128 B
};

__global__ void __launch_bounds__(NTHREADS, 1)
rec_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmG, KBwdArgs p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t w_full, full_bar[MAXSTAGES], empty_bar[MAXSTAGES], tfull_bar;
  __shared__ uint32_t tmem_slot;
  const RecTcBwdArgs& a = p.a;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x;
  const int H = p.H, B = p.B, T = a.T, nkb = p.nkb, G = 4 * p.H;
  constexpr uint32_t W_KB_BYTES = BW_U * 128;                       // 2 KB per K-block
  unsigned char* sW = smem;                                          // [nkb][16 rows x 128 B]
  unsigned char* sRing = smem + (size_t)nkb * W_KB_BYTES;            // [stages][Bpad x 128 B] (+ overhang pad)

  if (threadIdx.x == 0) {
    tc::mbar_init(&w_full, 1);
    for (int s = 0; s < p.stages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    tc::mbar_init(&tfull_bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 2) tc::tmem_alloc(&tmem_slot, 32);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 2) {
    if (lane == 0) {
      tc::tma_prefetch_desc(&tmW); tc::tma_prefetch_desc(&tmG);
      tc::mbar_arrive_expect_tx(&w_full, (uint32_t)nkb * W_KB_BYTES);
      for (int kb = 0; kb < nkb; ++kb) tc::tma_load_2d(sW + (size_t)kb * W_KB_BYTES, &tmW, kb * 64, j * BW_U, &w_full);
      const unsigned nctas = gridDim.x;
      uint32_t it = 0;
      unsigned epoch = 0;
      for (int t = T - 1; t >= 1; --t) {        // dh_{t-1} from dgates_t
        ++epoch;
        while (ld_acquire_u32(a.barrier) < nctas * epoch) {}
        tc::fence_proxy_async_all();
        RS_STAMP(a.dbg, t, 0);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          tc::mbar_wait(&empty_bar[s], ph ^ 1);
          tc::mbar_arrive_expect_tx(&full_bar[s], p.stage_bytes);
          tc::tma_load_2d(sRing + (size_t)s * p.stage_bytes, &tmG, kb * 64, t * B, &full_bar[s]);
        }
        RS_STAMP(a.dbg, t, 1);
      }
    }
  } else if (warp == 3) {
    if (lane == 0) {
      const uint32_t idesc = tc::instr_desc_bf16(128, BW_U);
      tc::mbar_wait(&w_full, 0);
      tc::tc_fence_after();
      const uint32_t sw = tc::smem_u32(sW);
      uint32_t it = 0;
      for (int t = T - 1; t >= 1; --t) {
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          tc::mbar_wait(&full_bar[s], ph);
          tc::tc_fence_after();
          if (kb == 0) RS_STAMP(a.dbg, t, 2);
          const uint64_t dg = tc::smem_desc_sw128(tc::smem_u32(sRing + (size_t)s * p.stage_bytes));
          const uint64_t dw = tc::smem_desc_sw128(sw + (uint32_t)kb * W_KB_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) tc::mma_bf16_ss(tmem, dg + 2 * k, dw + 2 * k, idesc, (kb | k) != 0);
          tc::mma_commit(&empty_bar[s]);
        }
        tc::mma_commit(&tfull_bar);
        RS_STAMP(a.dbg, t, 3);
      }
    }
  } else {
    const int q = warp & 3;                     // TMEM lane quadrant: batch rows 32q .. 32q+31
    const int hf = warp >> 2;                   // units 8hf .. 8hf+7 of the slice
    const int b = q * 32 + lane;
    const bool ok = b < B;
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
      __threadfence();
      tc::fence_proxy_async_all();
      epi_bar_sync();
      if (threadIdx.x == 0) RS_STAMP(a.dbg, t, 6);
      if (threadIdx.x == 0 && t > 0) red_release_add(a.barrier, 1u);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem, 32);
}

}  // namespace

bool rec_tc_bwd_geometry(int H, int B, RecTcBwdGeom* g) {
  if (H % 64 != 0 || H < 64 || B < 1 || B > 64) return false;
  if (H / BW_U > sm_count()) return false;
  const int Bpad = (B + 15) / 16 * 16;
  const size_t budget = 227 * 1024 - 2048;
  const int nkb = 4 * H / 64;
  const size_t w_bytes = (size_t)nkb * BW_U * 128;
  const size_t stage = (size_t)Bpad * 128;
  const size_t overhang = 128 * 128 - stage;          // M = 128 reads 128 rows from each stage base
  if (w_bytes + 2 * stage + overhang > budget) return false;
  int stages = (int)((budget - w_bytes - overhang) / stage);
  if (stages > MAXSTAGES) stages = MAXSTAGES;
  g->H = H; g->B = B; g->Bpad = Bpad; g->nslice = H / BW_U; g->stages = stages;
  g->smem_bytes = w_bytes + (size_t)stages * stage + overhang + 1024;
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
  p.H = g.H; p.B = g.B; p.Bpad = g.Bpad; p.stages = g.stages; p.nkb = 4 * g.H / 64;
  p.stage_bytes = (uint32_t)g.Bpad * 128;
  int dev = 0, per_sm = 0, nsm = 0;
  RS_CHECK_CUDA(cudaGetDevice(&dev));
  RS_CHECK_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  RS_CHECK_CUDA(cudaFuncSetAttribute(rec_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes));
  RS_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rec_tc_bwd_kernel, NTHREADS, g.smem_bytes));
  RS_REQUIRE(per_sm * nsm >= g.nslice, RS_ERR_UNSUPPORTED, "lstm_rec_tc_backward: %d CTAs cannot be co-resident", g.nslice);
  RS_CHECK_CUDA(cudaMemsetAsync(a.barrier, 0, sizeof(unsigned), st));
  void* kargs[] = {(void*)&tw, (void*)&tg, (void*)&p};
  RS_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)rec_tc_bwd_kernel, dim3(g.nslice), dim3(NTHREADS), kargs,
                                            g.smem_bytes, st));
  count_launch();
  return RS_OK;
}

namespace {

// wrec[(unit/U)*4U + (unit%U)*4 + g][k] = Wh[k][g*H + unit]; Wh = kernel + H*4H (row-major [H,4H])
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
  const int nkb = H / 64;
  const int nsm = sm_count();
  for (int U = 16; U >= 8; U /= 2) {
    if (H % U != 0 || H / U > nsm) continue;
    const size_t a_bytes = 2 * (size_t)nkb * (4 * U) * 128;
    const size_t stage = 2 * (size_t)Bpad * 128;
    if (a_bytes + 2 * stage > budget) continue;
    // the M = 128 instruction reads 128 rows from each K-block base: the ring behind the
    // resident operand must cover that overhang
    int stages = (int)((budget - a_bytes) / stage);
    if (stages > MAXSTAGES) stages = MAXSTAGES;
    if (stages > nkb) stages = nkb;
    if (stages < 2) continue;
    const size_t overhang = 128 * 128 - (size_t)(4 * U) * 128;
    if ((size_t)stages * stage < overhang) continue;
    g->H = H; g->B = B; g->Bpad = Bpad; g->U = U; g->nslice = H / U; g->stages = stages;
    g->smem_bytes = a_bytes + (size_t)stages * stage + 1024;
    return true;
  }
  return false;
}

int pack_wrec(const float* kernel, int H, int U, __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t st) {
  pack_wrec_kernel<<<dim3(cdiv(4 * H, 32), cdiv(H, 32)), dim3(32, 8), 0, st>>>(kernel + (size_t)H * 4 * H, H, U, hi, lo);
  RS_CHECK_LAUNCH();
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
  p.H = g.H; p.B = g.B; p.Bpad = g.Bpad; p.U = g.U; p.nslice = g.nslice; p.stages = g.stages; p.nkb = g.H / 64;
  p.a_kb_bytes = (uint32_t)(4 * g.U) * 128;
  p.stage_bytes = 2u * (uint32_t)g.Bpad * 128;
  int dev = 0, per_sm = 0, nsm = 0;
  RS_CHECK_CUDA(cudaGetDevice(&dev));
  RS_CHECK_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  RS_CHECK_CUDA(cudaFuncSetAttribute(rec_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes));
  RS_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rec_tc_fwd_kernel, NTHREADS, g.smem_bytes));
  RS_REQUIRE(per_sm * nsm >= g.nslice, RS_ERR_UNSUPPORTED, "lstm_rec_tc_forward: %d CTAs cannot be co-resident", g.nslice);
  RS_CHECK_CUDA(cudaMemsetAsync(a.barrier, 0, sizeof(unsigned), st));
  void* kargs[] = {(void*)&tw_hi, (void*)&tw_lo, (void*)&th_hi, (void*)&th_lo, (void*)&p};
  RS_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)rec_tc_fwd_kernel, dim3(g.nslice), dim3(NTHREADS), kargs,
                                            g.smem_bytes, st));
  count_launch();
  return RS_OK;
}

}  // namespace rs
