// CTC loss + gradient and greedy decode for sm_100a.
//
// Replaces tf.nn.ctc_loss(...) / prediction at
// /root/reference/models/AcousticModel.py:357 and :312-314.  The lattice rules
// are TF's (tensorflow/core/util/ctc/ctc_loss_calculator.{h,cc}), including the
// asymmetric alpha/beta skip test that is live because the reference's EOS label
// id equals the blank id; see oracle/ctc.py for the restated rules.
//
// Kernels
//   ctc_lse_kernel        one warp per (t,b) row: log-sum-exp over the C logits
//   ctc_lattice_kernel    2 CTAs per item (blockIdx.y = 0: alpha forward in time,
//                         1: beta backward in time), lattice row double-buffered
//                         in shared memory, rows streamed to the workspace
//   ctc_grad_kernel       one warp per (t,b) row: y - sum_u exp(alpha+beta-logp)
//   ctc_greedy_kernel     one CTA per item: per-frame argmax, collapse, compact
#include "common.cuh"

namespace rs {

static constexpr float kNegInf = -INFINITY;

__device__ __forceinline__ float lse2(float a, float b) {
  float m = fmaxf(a, b);
  if (m == kNegInf) return kNegInf;
  return m + logf(expf(a - m) + expf(b - m));
}
__device__ __forceinline__ float lse3(float a, float b, float c) {
  float m = fmaxf(a, fmaxf(b, c));
  if (m == kNegInf) return kNegInf;
  return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

// lse[t*B+b] = logsumexp_k logits[t,b,k]   (rows with t >= len[b] are skipped)
__global__ void ctc_lse_kernel(const float* __restrict__ logits, const int* __restrict__ len,
                               int T, int B, int C, float* __restrict__ lse) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= T * B) return;
  int t = row / B, b = row - t * B;
  if (t >= len[b]) return;
  const float* p = logits + (size_t)row * C;
  float m = kNegInf;
  for (int k = lane; k < C; k += 32) m = fmaxf(m, p[k]);
  m = warp_max(m);
  float s = 0.f;
  for (int k = lane; k < C; k += 32) s += expf(p[k] - m);
  s = warp_sum(s);
  if (lane == 0) lse[row] = m + logf(s);
}

// status per item: 0 = normal, 1 = skipped (len == 0 or labels longer than input)
struct CtcItem {
  int L;       // frames
  int N;       // labels
  int U;       // 2N+1
  int skip;
};

__device__ __forceinline__ CtcItem ctc_item(const int* len, const int* label_offsets, int b) {
  CtcItem it;
  it.L = len[b];
  it.N = label_offsets[b + 1] - label_offsets[b];
  it.U = 2 * it.N + 1;
  it.skip = (it.L == 0 || it.N > it.L) ? 1 : 0;
  return it;
}

// Lattice rows are stored SHIFTED: alpha~[t,u] = alpha[t,u] - Oa[t], beta~[t,u] = beta[t,u] - Ob[t],
// with the per-(item, t) offsets Oa / Ob accumulated in double.  The shift applied at step t is
// the maximum of the previous row, so the fp32 lattice values stay O(|log y|) instead of growing
// to |log p(z|x)| (thousands for 10 s utterances, where an fp32 ulp is 2.4e-4): the gradient then
// agrees with a float64 evaluation to ~1e-5 instead of ~1e-2 (TF's own fp32 lattice has the
// latter error).  The shift is exact bookkeeping, not an approximation.
//
// dynamic smem: int lp[Upad]; float row[2][Upad + 2]; float wmax[2][32]
__global__ void __launch_bounds__(1024)
ctc_lattice_kernel(const float* __restrict__ logits, const float* __restrict__ lse,
                   const int* __restrict__ labels, const int* __restrict__ label_offsets,
                   const int* __restrict__ len, int T, int B, int C, int Upad, int blank,
                   int beta_skip_dest, float* __restrict__ alpha, float* __restrict__ beta,
                   double* __restrict__ offs_a, double* __restrict__ offs_b, double* __restrict__ logp_out,
                   float* __restrict__ loss) {
  extern __shared__ unsigned char smem_raw[];
  int* lp = reinterpret_cast<int*>(smem_raw);
  float* rowbuf = reinterpret_cast<float*>(lp + Upad);
  float* wmax = rowbuf + 2 * (Upad + 2);
  const int b = blockIdx.x;
  const bool is_beta = blockIdx.y == 1;
  const CtcItem it = ctc_item(len, label_offsets, b);
  if (it.skip) {
    if (is_beta && threadIdx.x == 0) { loss[b] = 0.f; logp_out[b] = 0.0; }
    return;
  }
  const int U = it.U, L = it.L;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int* lab = labels + label_offsets[b];
  for (int u = threadIdx.x; u < U; u += blockDim.x) lp[u] = (u & 1) ? lab[u >> 1] : blank;
  const int stride = Upad + 2;
  // row r lives at rowbuf[r*stride + 2 + u] (alpha: two -inf guards in front)
  //                 rowbuf[r*stride + u]     (beta: two -inf guards behind)
  for (int i = threadIdx.x; i < 2 * stride; i += blockDim.x) rowbuf[i] = kNegInf;
  __syncthreads();
  float* out = (is_beta ? beta : alpha) + (size_t)b * T * Upad;
  double* offs = (is_beta ? offs_b : offs_a) + (size_t)b * T;
  double O = 0.0;
  auto row_max = [&](int slot) -> float {   // max over the per-warp maxima written last step
    float m = kNegInf;
    for (int w = 0; w < nw; ++w) m = fmaxf(m, wmax[slot * 32 + w]);
    return (m == kNegInf) ? 0.f : m;
  };

  if (!is_beta) {
    float* prev = rowbuf + 2;
    float* cur = rowbuf + stride + 2;
    const float* lg = logits + (size_t)b * C;
    const float l0 = lse[b];
    float mloc = kNegInf;
    for (int u = threadIdx.x; u < U; u += blockDim.x) {
      float v = kNegInf;
      if (u == 0) v = lg[blank] - l0;
      else if (u == 1) v = lg[lp[1]] - l0;
      prev[u] = v;
      out[u] = v;
      mloc = fmaxf(mloc, v);
    }
    mloc = warp_max(mloc);
    if (lane == 0) wmax[warp] = mloc;
    if (threadIdx.x == 0) offs[0] = 0.0;
    __syncthreads();
    // The emission log-probability of this thread's lattice state does not depend on the recurrence, but a gather
    // issued where it is used puts an L2 round trip on the chain of T dependent steps (ncu: half of the kernel's
    // stall samples).  It is loaded two steps ahead instead -- the loop is unrolled by two so that the two values in
    // flight live in fixed registers (moving a register whose load is pending would wait for it).
    const int u0 = threadIdx.x;
    const int l_u0 = (u0 < U) ? lp[u0] : blank;
    // raw (logit, log-sum-exp) pair: the subtraction happens where the value is used, after the step's barrier
    auto emission = [&](int t) -> float2 {
      return (t < L) ? make_float2(logits[((size_t)t * B + b) * C + l_u0], lse[t * B + b]) : make_float2(0.f, 0.f);
    };
    auto alpha_step = [&](int t, float2 em) {
      const float M = row_max((t - 1) & 1);
      O += (double)M;
      if (threadIdx.x == 0) offs[t] = O;
      const float* lgt = logits + ((size_t)t * B + b) * C;
      const int lo = max(0, U - 2 * (L - t)), hi = min(U, 2 * (t + 1));
      mloc = kNegInf;
      for (int u = threadIdx.x; u < U; u += blockDim.x) {
        float v = kNegInf;
        if (u >= lo && u < hi) {
          const int l = lp[u];
          const bool skip = (u > 1) && (l != blank) && (l != lp[u - 2]);
          const float a0 = prev[u], a1 = prev[u - 1], a2 = skip ? prev[u - 2] : kNegInf;
          const float e = (u == u0) ? (em.x - em.y) : (lgt[l] - lse[t * B + b]);
          v = (lse3(a0, a1, a2) - M) + e;
        }
        cur[u] = v;
        out[(size_t)t * Upad + u] = v;
        mloc = fmaxf(mloc, v);
      }
      mloc = warp_max(mloc);
      if (lane == 0) wmax[(t & 1) * 32 + warp] = mloc;
      __syncthreads();
      float* tmp = prev; prev = cur; cur = tmp;
    };
    float2 em_a = emission(1), em_b = emission(2);
    int t = 1;
    for (; t + 1 < L; t += 2) {
      { const float2 e = em_a; em_a = emission(t + 2); alpha_step(t, e); }
      { const float2 e = em_b; em_b = emission(t + 3); alpha_step(t + 1, e); }
    }
    if (t < L) alpha_step(t, em_a);
  } else {
    float* prev = rowbuf;            // beta~[t+1] + logp[t+1]  ("nxt" in the oracle), offset Ob[t+1]
    float* cur = rowbuf + stride;
    float mloc = kNegInf;
    {
      const float* lgt = logits + ((size_t)(L - 1) * B + b) * C;
      const float lt = lse[(L - 1) * B + b];
      for (int u = threadIdx.x; u < U; u += blockDim.x) {
        float v = (u >= U - 2) ? 0.f : kNegInf;
        out[(size_t)(L - 1) * Upad + u] = v;
        const float nx = v + (lgt[lp[u]] - lt);
        prev[u] = nx;
        mloc = fmaxf(mloc, nx);
      }
    }
    mloc = warp_max(mloc);
    if (lane == 0) wmax[((L - 1) & 1) * 32 + warp] = mloc;
    if (threadIdx.x == 0) offs[L - 1] = 0.0;
    __syncthreads();
    const int u0 = threadIdx.x;                         // emission gathered two steps ahead: see the alpha loop
    const int l_u0 = (u0 < U) ? lp[u0] : blank;
    auto emission = [&](int t) -> float2 {
      return (t >= 0) ? make_float2(logits[((size_t)t * B + b) * C + l_u0], lse[t * B + b]) : make_float2(0.f, 0.f);
    };
    auto beta_step = [&](int t, float2 em) {
      const float M = row_max((t + 1) & 1);
      O += (double)M;
      if (threadIdx.x == 0) offs[t] = O;
      const float* lgt = logits + ((size_t)t * B + b) * C;
      const int lo = max(0, U - 2 * (L - t)), hi = min(U, 2 * (t + 1));
      mloc = kNegInf;
      for (int u = threadIdx.x; u < U; u += blockDim.x) {
        float v = kNegInf;
        const int l = lp[u];
        if (u >= lo && u < hi) {
          bool skip = false;
          if (u + 2 < U) {
            const int l2 = lp[u + 2];
            skip = beta_skip_dest ? (l2 != blank && l2 != l) : (l != blank && l != l2);
          }
          const float b0 = prev[u], b1 = prev[u + 1], b2 = skip ? prev[u + 2] : kNegInf;
          v = lse3(b0, b1, b2) - M;
        }
        out[(size_t)t * Upad + u] = v;
        const float nx = v + ((u == u0) ? (em.x - em.y) : (lgt[l] - lse[t * B + b]));   // becomes "nxt" for row t-1
        cur[u] = nx;
        mloc = fmaxf(mloc, nx);
      }
      mloc = warp_max(mloc);
      if (lane == 0) wmax[(t & 1) * 32 + warp] = mloc;
      __syncthreads();
      float* tmp = prev; prev = cur; cur = tmp;
    };
    float2 em_a = emission(L - 2), em_b = emission(L - 3);
    int t = L - 2;
    for (; t - 1 >= 0; t -= 2) {
      { const float2 e = em_a; em_a = emission(t - 2); beta_step(t, e); }
      { const float2 e = em_b; em_b = emission(t - 3); beta_step(t - 1, e); }
    }
    if (t >= 0) beta_step(t, em_a);
    // log p(z|x) = LSE_u(alpha[0,u] + beta[0,u]); alpha[0,u] = logp[0,l'u] for u in {0,1}
    // and prev[u] now holds beta~[0,u] + logp[0,l'u] (offset O = Ob[0]).
    if (threadIdx.x == 0) {
      const float lpz = (U > 1) ? lse2(prev[0], prev[1]) : prev[0];
      const double logp = (lpz == kNegInf) ? -INFINITY : (double)lpz + O;
      logp_out[b] = logp;
      loss[b] = (float)(-logp);   // +inf when no valid path
    }
  }
}

// ------------------------------------------------------------------------------------
// ctc_lattice1_kernel: the same lattice with ONE state per thread (2 N + 1 <= blockDim) and a short dependent chain
// per time step.  The lattice is a chain of T dependent steps per item, so the kernel's time is T x (latency of one
// step); ncu of the general kernel above and of this kernel's first version shows no memory or pipe limit, only
// instructions that wait for each other (~6 cycles each).  What is on the chain here: three shared-memory loads, two
// maxima, three ex2, two adds, one lg2, two adds, one store, the barrier.  Everything else is kept off it:
//   * base-2 logarithms throughout (ex2.approx / lg2.approx are single SFU instructions; absolute error ~2e-7 per
//     step on values of order one, against 1e-7 for libm; tests/test_gpu_ctc.py holds the gradient to 1e-4 / 1e-3).
//     The rows in the workspace and the offsets are in base-2 units, log p(z|x) is converted at the end, and the
//     gradient kernel is told (log2_domain);
//   * the shift of row t is derived from the maximum of row t-3, not t-1: O[t] = O[t-3] + max(row t-3), i.e.
//     M[t] = max(row t-3) - M[t-1] - M[t-2].  Still exact bookkeeping (the offsets are summed in double and the
//     gradient kernel adds them back), the stored values stay within three steps' emissions of zero, and the per-warp
//     maxima of row t-1 (one REDUX over order-preserving integer keys) are produced beside step t's log-sum-exp,
//     fetched before its barrier and combined in the shadow of step t+1's loads;
//   * label, skip flag and window of the thread's state live in registers, pointers advance by constants, emissions
//     are gathered 8 steps ahead, the window is a -inf added to the emission instead of a branch.
// dynamic smem: float row[2][Upad + 2]; uint32 wkey[2][32]
// ------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// log2(2^a + 2^b + 2^c); all three -inf: the maximum is clamped to a finite value, 2^(-inf) = 0, lg2(0) = -inf
__device__ __forceinline__ float l2se3(float a, float b, float c) {
  const float m = fmaxf(fmaxf(a, b), fmaxf(c, -3.0e38f));
  return m + lg2_approx(ex2_approx(a - m) + ex2_approx(b - m) + ex2_approx(c - m));
}
// order-preserving image of a float in the unsigned integers, and back
__device__ __forceinline__ uint32_t fkey(float v) {
  const uint32_t bits = __float_as_uint(v);
  return (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

template <int MAXT>
__global__ void __launch_bounds__(MAXT)
ctc_lattice1_kernel(const float* __restrict__ logits, const float* __restrict__ lse,
                    const int* __restrict__ labels, const int* __restrict__ label_offsets,
                    const int* __restrict__ len, int T, int B, int C, int Upad, int blank,
                    int beta_skip_dest, float* __restrict__ alpha, float* __restrict__ beta,
                    double* __restrict__ offs_a, double* __restrict__ offs_b, double* __restrict__ logp_out,
                    float* __restrict__ loss) {
  extern __shared__ unsigned char smem_raw[];
  float* rowbuf = reinterpret_cast<float*>(smem_raw);
  const int stride = Upad + 2;
  uint32_t* wkey = reinterpret_cast<uint32_t*>(rowbuf + 2 * stride);
  const int b = blockIdx.x;
  const bool is_beta = blockIdx.y == 1;
  const CtcItem it = ctc_item(len, label_offsets, b);
  if (it.skip) {
    if (is_beta && threadIdx.x == 0) { loss[b] = 0.f; logp_out[b] = 0.0; }
    return;
  }
  constexpr float kLog2e = 1.4426950408889634f;
  constexpr double kLn2 = 0.6931471805599453;
  const int U = it.U, L = it.L;
  const int u = threadIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NW4 = MAXT / 128;                        // uint4 loads that cover one per-warp maximum per warp
  const int* lab = labels + label_offsets[b];
  auto lab_at = [&](int v) -> int { return (v >= 0 && v < U) ? ((v & 1) ? lab[v >> 1] : blank) : -1; };
  const int l_u = (u < U) ? lab_at(u) : blank;
  // row r of the lattice lives at rowbuf[(r & 1) * stride + 2 + v] (alpha: two -inf guards in front)
  //                               rowbuf[(r & 1) * stride + v]     (beta: two -inf guards behind)
  for (int i = threadIdx.x; i < 2 * stride; i += blockDim.x) rowbuf[i] = kNegInf;
  for (int i = threadIdx.x; i < 64; i += blockDim.x) wkey[i] = fkey(kNegInf);      // (a block may be a single warp)
  __syncthreads();
  const int dir = is_beta ? -1 : 1;                      // direction of time
  const int tfirst = is_beta ? L - 1 : 0;                // the row that is written directly
  double O = 0.0;                                        // thread 0: offset of the current row (base-2 units)
  float M1 = 0.f, M2 = 0.f, vlast;                       // the shifts of the two previous steps
  constexpr int D = MAXT <= 256 ? 8 : 2;                 // emissions in flight (what the register budget of 1024 threads allows)
  float2 em[D];
  // running pointers (the chain of T dependent steps has no room for address arithmetic): emission of the next step to
  // be gathered, lattice row / offset of the next step to be written
  const ptrdiff_t BC = (ptrdiff_t)B * C;
  const float* pe = logits + ((ptrdiff_t)(tfirst + dir) * B + b) * C + l_u;
  const float* pl = lse + (ptrdiff_t)(tfirst + dir) * B + b;
  int te = tfirst + dir;                                 // the step pe / pl point at
  auto emission = [&]() -> float2 {
    const bool ok = is_beta ? te >= 0 : te < L;
    const float2 e = ok ? make_float2(__ldg(pe), __ldg(pl)) : make_float2(0.f, 0.f);
    pe += dir * BC; pl += dir * B; te += dir;
    return e;
  };
  float* po = (is_beta ? beta : alpha) + ((size_t)b * T + tfirst) * Upad + u;
  double* poffs = (is_beta ? offs_b : offs_a) + (size_t)b * T + tfirst;
  uint4 wk[NW4];                                         // per-warp maxima (keys) of the row three steps back, fetched before the barrier
#pragma unroll
  for (int i = 0; i < NW4; ++i) wk[i] = make_uint4(0u, 0u, 0u, 0u);
  // the shift of this step from the keys fetched during the previous one (uniform over the block)
  auto shift = [&](bool have) -> float {
    float M = 0.f;
    if (have) {
      uint32_t k = 0u;
#pragma unroll
      for (int i = 0; i < NW4; ++i) k = max(max(k, max(wk[i].x, wk[i].y)), max(wk[i].z, wk[i].w));
      const float m = fkey_inv(k);
      M = ((m == kNegInf) ? 0.f : m) - M1 - M2;
    }
    M2 = M1; M1 = M;
    return M;
  };
  // end of a step: the maxima of the row before the one just written go to slot `wslot`, the maxima written during the
  // previous step (slot `rslot`) are fetched for the next step's shift
  auto row_end = [&](int wslot, int rslot) {
    const uint32_t k = __reduce_max_sync(0xffffffffu, fkey(vlast));
    if (lane == 0) wkey[wslot * 32 + warp] = k;
    const uint4* w4 = reinterpret_cast<const uint4*>(wkey + rslot * 32);
#pragma unroll
    for (int i = 0; i < NW4; ++i) wk[i] = w4[i];
  };

  if (!is_beta) {
    const bool skip = (u > 1) && (u < U) && (l_u != blank) && (l_u != lab_at(u - 2));
    {
      const float* lg = logits + (size_t)b * C;
      float v = kNegInf;
      if (u == 0) v = (lg[blank] - lse[b]) * kLog2e;
      else if (u == 1 && U > 1) v = (lg[l_u] - lse[b]) * kLog2e;
      rowbuf[2 + u] = v;
      if (u < U) *po = v;
      vlast = v;
      if (threadIdx.x == 0) *poffs = 0.0;
    }
#pragma unroll
    for (int i = 0; i < D; ++i) em[i] = emission();
    __syncthreads();
    // step t writes row t from row t-1; its shift comes from row t-3 (t >= 3)
    auto step = [&](int t, float2 e) {
      const float* prev = rowbuf + ((t - 1) & 1) * stride + 2;
      float* cur = rowbuf + (t & 1) * stride + 2;
      const float a0 = prev[u], a1 = prev[u - 1], a2 = skip ? prev[u - 2] : kNegInf;
      const float M = shift(t >= 3);
      po += Upad; poffs += 1;
      if (threadIdx.x == 0) { O += (double)M; *poffs = O; }
      const bool inwin = (u >= U - 2 * (L - t)) && (u < U) && (u < 2 * (t + 1));
      const float add = fmaf(e.x - e.y, kLog2e, -M) + (inwin ? 0.f : kNegInf);
      const float v = l2se3(a0, a1, a2) + add;
      cur[u] = v;
      if (u < U) *po = v;
      row_end(t & 1, (t - 1) & 1);                       // keys of row t-1 out, keys of row t-2 in (for step t+1)
      vlast = v;
      __syncthreads();
    };
    int t = 1;
    for (; t + D <= L; t += D) {
#pragma unroll
      for (int i = 0; i < D; ++i) { const float2 e = em[i]; em[i] = emission(); step(t + i, e); }
    }
#pragma unroll
    for (int i = 0; i < D; ++i) if (t + i < L) step(t + i, em[i]);
  } else {
    bool skip = false;
    if (u + 2 < U) {
      const int l2 = lab_at(u + 2);
      skip = beta_skip_dest ? (l2 != blank && l2 != l_u) : (l_u != blank && l_u != l2);
    }
    {
      const float* lgt = logits + ((size_t)(L - 1) * B + b) * C;
      const float v = (u < U && u >= U - 2) ? 0.f : kNegInf;
      if (u < U) *po = v;
      const float nx = (u < U) ? v + (lgt[l_u] - lse[(L - 1) * B + b]) * kLog2e : kNegInf;
      rowbuf[((L - 1) & 1) * stride + u] = nx;          // beta~[t] + logp[t]: what row t-1 sums over
      vlast = nx;
      if (threadIdx.x == 0) *poffs = 0.0;
    }
#pragma unroll
    for (int i = 0; i < D; ++i) em[i] = emission();
    __syncthreads();
    // step t writes row t from row t+1; its shift comes from row t+3 (t <= L-4)
    auto step = [&](int t, float2 e) {
      const float* prev = rowbuf + ((t + 1) & 1) * stride;
      float* cur = rowbuf + (t & 1) * stride;
      const float b0 = prev[u], b1 = prev[u + 1], b2 = skip ? prev[u + 2] : kNegInf;
      const float M = shift(t <= L - 4);
      po -= Upad; poffs -= 1;
      if (threadIdx.x == 0) { O += (double)M; *poffs = O; }
      const bool inwin = (u >= U - 2 * (L - t)) && (u < U) && (u < 2 * (t + 1));
      const float win = inwin ? 0.f : kNegInf;
      const float s = l2se3(b0, b1, b2);
      const float nx = s + (fmaf(e.x - e.y, kLog2e, -M) + win);      // becomes the summand of row t-1
      cur[u] = nx;
      if (u < U) *po = s + (win - M);                                // beta~[t]
      row_end(t & 1, (t + 1) & 1);                       // keys of row t+1 out, keys of row t+2 in (for step t-1)
      vlast = nx;
      __syncthreads();
    };
    int t = L - 2;
    for (; t - D + 1 >= 0; t -= D) {
#pragma unroll
      for (int i = 0; i < D; ++i) { const float2 e = em[i]; em[i] = emission(); step(t - i, e); }
    }
#pragma unroll
    for (int i = 0; i < D; ++i) if (t - i >= 0) step(t - i, em[i]);
    // log p(z|x) = LSE_u(alpha[0,u] + beta[0,u]); alpha[0,u] = logp[0,l'u] for u in {0,1}; row 0 of this pass holds
    // beta~[0,u] + logp[0,l'u] with offset O = Ob[0], all in base-2 units
    if (threadIdx.x == 0) {
      const float* r0 = rowbuf;
      const float m = (U > 1) ? fmaxf(r0[0], r0[1]) : r0[0];
      double logp = -INFINITY;
      if (m != kNegInf) {
        const float lpz = (U > 1) ? m + log2f(exp2f(r0[0] - m) + exp2f(r0[1] - m)) : m;
        logp = ((double)lpz + O) * kLn2;
      }
      logp_out[b] = logp;
      loss[b] = (float)(-logp);   // +inf when no valid path
    }
  }
}

// one warp per (t,b) row.  dynamic smem: float acc[warps][C]
__global__ void ctc_grad_kernel(const float* __restrict__ logits, const float* __restrict__ lse,
                                const int* __restrict__ labels, const int* __restrict__ label_offsets,
                                const int* __restrict__ len, int T, int B, int C, int Upad, int blank,
                                const float* __restrict__ alpha, const float* __restrict__ beta,
                                const double* __restrict__ offs_a, const double* __restrict__ offs_b,
                                const double* __restrict__ logp_in, int log2_domain, float* __restrict__ grad) {
  extern __shared__ float acc_all[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= T * B) return;
  const int t = row / B, b = row - t * B;
  float* g = grad + (size_t)row * C;
  const CtcItem it = ctc_item(len, label_offsets, b);
  if (it.skip || t >= it.L) {
    for (int k = lane; k < C; k += 32) g[k] = 0.f;
    return;
  }
  const float* p = logits + (size_t)row * C;
  const float lt = lse[row];
  const double logp = logp_in[b];
  if (logp == -INFINITY) {             // "No valid path found": dy = y
    for (int k = lane; k < C; k += 32) g[k] = expf(p[k] - lt);
    return;
  }
  // alpha + beta - log p = alpha~ + beta~ + shift, the scalar part evaluated in double
  // (rows and offsets written by ctc_lattice1_kernel are in base-2 units, log p is always natural)
  const float shift = (float)(offs_a[(size_t)b * T + t] + offs_b[(size_t)b * T + t] - (log2_domain ? logp * 1.4426950408889634 : logp));
  float* acc = acc_all + warp * C;
  for (int k = lane; k < C; k += 32) acc[k] = 0.f;
  __syncwarp();
  const float* a = alpha + ((size_t)b * T + t) * Upad;
  const float* be = beta + ((size_t)b * T + t) * Upad;
  const int* lab = labels + label_offsets[b];
  const int U = it.U;
  float blank_sum = 0.f;
  for (int u = lane; u < U; u += 32) {
    const float ab = a[u] + be[u];
    const float w = (ab == kNegInf) ? 0.f : (log2_domain ? exp2f(ab + shift) : expf(ab + shift));
    if (u & 1) {
      const int l = lab[u >> 1];
      atomicAdd(&acc[l], w);
    } else {
      blank_sum += w;
    }
  }
  blank_sum = warp_sum(blank_sum);
  __syncwarp();
  if (lane == 0) acc[blank] += blank_sum;
  __syncwarp();
  for (int k = lane; k < C; k += 32) g[k] = expf(p[k] - lt) - acc[k];
}

// one CTA per item.  dynamic smem: int path[T]; int counts[blockDim]
__global__ void ctc_greedy_kernel(const float* __restrict__ logits, const int* __restrict__ len,
                                  int T, int B, int C, int blank, int* __restrict__ out,
                                  int* __restrict__ out_len) {
  extern __shared__ int gsm[];
  int* path = gsm;
  int* counts = gsm + T;
  const int b = blockIdx.x;
  const int L = min(len[b], T);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int t = warp; t < L; t += nw) {
    const float* p = logits + ((size_t)t * B + b) * C;
    float best = kNegInf;
    int bi = 0x7fffffff;
    for (int k = lane; k < C; k += 32) {
      float v = p[k];
      if (v > best || (v == best && k < bi)) { best = v; bi = k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) path[t] = bi;
  }
  __syncthreads();
  const int seg = (L + blockDim.x - 1) / blockDim.x;
  const int t0 = min(L, (int)threadIdx.x * seg), t1 = min(L, t0 + seg);
  int cnt = 0;
  for (int t = t0; t < t1; ++t) {
    const int c = path[t];
    if (c != blank && (t == 0 || c != path[t - 1])) ++cnt;
  }
  counts[threadIdx.x] = cnt;
  __syncthreads();
  // exclusive scan (blockDim <= 1024): simple Hillis-Steele in shared memory
  for (int o = 1; o < blockDim.x; o <<= 1) {
    int v = (threadIdx.x >= o) ? counts[threadIdx.x - o] : 0;
    __syncthreads();
    counts[threadIdx.x] += v;
    __syncthreads();
  }
  int pos = counts[threadIdx.x] - cnt;
  int* o = out + (size_t)b * T;
  for (int t = t0; t < t1; ++t) {
    const int c = path[t];
    if (c != blank && (t == 0 || c != path[t - 1])) o[pos++] = c;
  }
  const int total = counts[blockDim.x - 1];
  if (threadIdx.x == 0) out_len[b] = total;
  __syncthreads();
  for (int t = total + threadIdx.x; t < T; t += blockDim.x) o[t] = -1;
}

}  // namespace rs

using namespace rs;

static inline int ctc_upad(int max_label_len) { return (int)align_up((size_t)(2 * max_label_len + 1), 32); }

extern "C" size_t rs_ctc_workspace_bytes(int T, int B, int C, int max_label_len) {
  (void)C;
  size_t upad = ctc_upad(max_label_len);
  size_t lse = align_up((size_t)T * B * sizeof(float), 256);
  size_t lat = align_up((size_t)T * B * upad * sizeof(float), 256);
  size_t offs = align_up((size_t)T * B * sizeof(double), 256);
  return lse + 2 * lat + 2 * offs + align_up((size_t)B * sizeof(double), 256);
}

extern "C" int rs_ctc_loss_grad(const float* logits_d, const int32_t* labels_d,
                                const int32_t* label_offsets_d, const int32_t* len_d, int T, int B,
                                int C, int max_label_len, int blank, int beta_skip, float* loss_d,
                                float* grad_d, void* ws_d, size_t ws_bytes, void* stream) {
  RS_REQUIRE(T > 0 && B > 0 && C > 1, RS_ERR_INVALID, "rs_ctc_loss_grad: bad shape T=%d B=%d C=%d", T, B, C);
  RS_REQUIRE(blank >= 0 && blank < C, RS_ERR_INVALID, "rs_ctc_loss_grad: blank %d outside [0,%d)", blank, C);
  RS_REQUIRE(max_label_len >= 0, RS_ERR_INVALID, "rs_ctc_loss_grad: max_label_len < 0");
  RS_REQUIRE(beta_skip == RS_CTC_BETA_SOURCE || beta_skip == RS_CTC_BETA_DEST, RS_ERR_INVALID,
             "rs_ctc_loss_grad: unknown beta_skip %d", beta_skip);
  RS_REQUIRE(ws_bytes >= rs_ctc_workspace_bytes(T, B, C, max_label_len), RS_ERR_WORKSPACE,
             "rs_ctc_loss_grad: workspace %zu < %zu", ws_bytes, rs_ctc_workspace_bytes(T, B, C, max_label_len));
  cudaStream_t st = (cudaStream_t)stream;
  const int upad = ctc_upad(max_label_len);
  char* ws = (char*)ws_d;
  float* lse = (float*)ws;
  size_t lse_b = align_up((size_t)T * B * sizeof(float), 256);
  size_t lat_b = align_up((size_t)T * B * upad * sizeof(float), 256);
  float* alpha = (float*)(ws + lse_b);
  float* beta = (float*)(ws + lse_b + lat_b);
  size_t offs_b_ = align_up((size_t)T * B * sizeof(double), 256);
  double* offs_a = (double*)(ws + lse_b + 2 * lat_b);
  double* offs_b = (double*)(ws + lse_b + 2 * lat_b + offs_b_);
  double* logp = (double*)(ws + lse_b + 2 * lat_b + 2 * offs_b_);

  const int rows = T * B;
  ctc_lse_kernel<<<cdiv(rows, 8), 256, 0, st>>>(logits_d, len_d, T, B, C, lse);
  RS_CHECK_LAUNCH();
  const int U = 2 * max_label_len + 1;
  int threads = (int)align_up((size_t)U, 32);
  if (threads > 1024) threads = 1024;
  size_t smem = (size_t)upad * sizeof(int) + 2 * (size_t)(upad + 2) * sizeof(float) + 64 * sizeof(float);
  RS_REQUIRE(smem <= 200 * 1024, RS_ERR_UNSUPPORTED, "rs_ctc_loss_grad: label length %d too large", max_label_len);
  // one state per thread when every lattice row fits a block (RS_CTC_LATTICE=0: the general kernel)
  static const bool one_env = [] { const char* v = getenv("RS_CTC_LATTICE"); return !(v && v[0] == '0'); }();
  const bool one = one_env && (int)align_up((size_t)U, 32) <= 1024;
  if (one) {
    const size_t smem1 = 2 * (size_t)(upad + 2) * sizeof(float) + 64 * sizeof(float) + 16;
    auto k = threads <= 256 ? ctc_lattice1_kernel<256> : ctc_lattice1_kernel<1024>;
    k<<<dim3(B, 2), threads, smem1, st>>>(logits_d, lse, labels_d, label_offsets_d, len_d, T, B, C, upad, blank,
                                         beta_skip == RS_CTC_BETA_DEST ? 1 : 0, alpha, beta, offs_a, offs_b, logp, loss_d);
  } else {
    if (smem > 48 * 1024)
      RS_CHECK_CUDA(cudaFuncSetAttribute(ctc_lattice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctc_lattice_kernel<<<dim3(B, 2), threads, smem, st>>>(logits_d, lse, labels_d, label_offsets_d, len_d, T, B, C,
                                                          upad, blank, beta_skip == RS_CTC_BETA_DEST ? 1 : 0,
                                                          alpha, beta, offs_a, offs_b, logp, loss_d);
  }
  RS_CHECK_LAUNCH();
  if (grad_d) {
    const int warps = 8;
    size_t gsmem = (size_t)warps * C * sizeof(float);
    RS_REQUIRE(gsmem <= 48 * 1024, RS_ERR_UNSUPPORTED, "rs_ctc_loss_grad: C=%d too large", C);
    ctc_grad_kernel<<<cdiv(rows, warps), warps * 32, gsmem, st>>>(logits_d, lse, labels_d, label_offsets_d, len_d,
                                                                  T, B, C, upad, blank, alpha, beta, offs_a, offs_b, logp, one ? 1 : 0, grad_d);
    RS_CHECK_LAUNCH();
  }
  return RS_OK;
}

extern "C" int rs_ctc_greedy_decode(const float* logits_d, const int32_t* len_d, int T, int B, int C,
                                    int blank, int32_t* out_d, int32_t* out_len_d, void* stream) {
  RS_REQUIRE(T > 0 && B > 0 && C > 0, RS_ERR_INVALID, "rs_ctc_greedy_decode: bad shape T=%d B=%d C=%d", T, B, C);
  const int threads = 256;
  size_t smem = ((size_t)T + threads) * sizeof(int);
  RS_REQUIRE(smem <= 200 * 1024, RS_ERR_UNSUPPORTED, "rs_ctc_greedy_decode: T=%d too large", T);
  if (smem > 48 * 1024)
    RS_CHECK_CUDA(cudaFuncSetAttribute(ctc_greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ctc_greedy_kernel<<<B, threads, smem, (cudaStream_t)stream>>>(logits_d, len_d, T, B, C, blank, out_d, out_len_d);
  RS_CHECK_LAUNCH();
  return RS_OK;
}
