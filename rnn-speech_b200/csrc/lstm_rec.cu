// Persistent recurrent LSTM kernels, fp32 FFMA version.
//
// Replaces the tf.nn.dynamic_rnn while-loop over BasicLSTMCell
// (/root/reference/models/AcousticModel.py:227-237, :277-278) for one layer:
//   g_t = gx_t + h_{t-1} @ Wh;  i,j,f,o = split(g_t)
//   c_t = c_{t-1} * sigmoid(f + 1) + sigmoid(i) * tanh(j);  h_t = tanh(c_t) * sigmoid(o)
//   rows with t >= len[b]: output 0, state frozen.
//
// Layout / ownership: the grid is ceil(H/8) co-resident CTAs (cooperative launch);
// CTA j owns hidden units [8j, 8j+8): its 32 gate columns of Wh stay in shared
// memory for the whole sequence, its c/h state stays on chip, and each step the
// freshly written h_t slice is exchanged through L2 behind one grid barrier.
#include "lstm_rec.cuh"

namespace rs {
namespace {

constexpr int U = 8;            // hidden units per CTA
constexpr int GC = 4 * U;       // gate columns per CTA
constexpr int KC = 256;         // K-chunk staged per pass
constexpr int NT = 256;         // threads
constexpr int NW = NT / 32;

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// All threads of the CTA call this.  Writes made by any thread of the CTA before
// the call are visible to every CTA after its matching wait returns.
__device__ __forceinline__ void grid_arrive(unsigned* bar) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(bar, 1u);
}
__device__ __forceinline__ void grid_wait(unsigned* bar, unsigned target) {
  if (threadIdx.x == 0) {
    while (ld_acquire(bar) < target) { __nanosleep(20); }
  }
  __syncthreads();
}

// Stage a [32 rows (batch chunk)] x [KC cols] block of a row-major [B, ld] matrix
// transposed into st[kk][b] (row stride 33).  Rows >= B / cols >= ncols are zero.
__device__ __forceinline__ void stage_T(float* __restrict__ stg, const float* __restrict__ src, int ld,
                                        int b0, int B, int c0, int ncols, bool coherent) {
  for (int e = threadIdx.x; e < 32 * (KC / 4); e += NT) {
    const int r = e / (KC / 4), q = (e % (KC / 4)) * 4;
    const int b = b0 + r, c = c0 + q;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b < B && src != nullptr) {
      const float* p = src + (size_t)b * ld + c;
      if (c + 3 < ncols && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
        v = coherent ? __ldcg(reinterpret_cast<const float4*>(p)) : *reinterpret_cast<const float4*>(p);
      } else {
        if (c < ncols) v.x = coherent ? __ldcg(p) : p[0];
        if (c + 1 < ncols) v.y = coherent ? __ldcg(p + 1) : p[1];
        if (c + 2 < ncols) v.z = coherent ? __ldcg(p + 2) : p[2];
        if (c + 3 < ncols) v.w = coherent ? __ldcg(p + 3) : p[3];
      }
    }
    stg[(q + 0) * 33 + r] = v.x;
    stg[(q + 1) * 33 + r] = v.y;
    stg[(q + 2) * 33 + r] = v.z;
    stg[(q + 3) * 33 + r] = v.w;
  }
}

// dynamic smem: Ws[H][GC] | stg[KC][33] | red[NW][GC][33] | cst[Bpad][U] | hst[Bpad][U]
__global__ void __launch_bounds__(NT, 1) lstm_rec_fwd_kernel(RecFwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int T = a.T, B = a.B, H = a.H;
  const int nbc = (B + 31) / 32, Bpad = nbc * 32;
  float* Ws = smem;
  float* stg = Ws + (size_t)H * GC;
  float* red = stg + KC * 33;
  float* cst = red + NW * GC * 33;
  float* hst = cst + Bpad * U;
  const int u0 = blockIdx.x * U;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned nctas = gridDim.x;

  // resident weights: Ws[k][g*U + u] = Wh[k][g*H + u0 + u]
  for (int e = tid; e < H * GC; e += NT) {
    const int k = e / GC, c = e % GC, g = c / U, u = c % U;
    Ws[e] = (u0 + u < H) ? a.Wh[(size_t)k * 4 * H + (size_t)g * H + u0 + u] : 0.f;
  }
  for (int e = tid; e < Bpad * U; e += NT) {
    const int b = e / U, u = e % U;
    const bool ok = b < B && u0 + u < H;
    cst[e] = (ok && a.c0) ? a.c0[(size_t)b * H + u0 + u] : 0.f;
    hst[e] = (ok && a.h0) ? a.h0[(size_t)b * H + u0 + u] : 0.f;
  }
  __syncthreads();

  // elementwise-phase mapping: 8 units fastest -> 32-byte segments per row
  const int eu = tid % U, eb = tid / U;   // eb in [0,32)
  for (int t = 0; t < T; ++t) {
    if (t > 0) grid_wait(a.barrier, nctas * (unsigned)t);
    const float* hprev = (t == 0) ? a.h0 : a.out + (size_t)(t - 1) * B * H;
    for (int bc = 0; bc < nbc; ++bc) {
      const int b0 = bc * 32;
      // prefetch this thread's 4 gate pre-activations from gx
      const int b_e = b0 + eb, unit = u0 + eu;
      const bool e_ok = b_e < B && unit < H;
      float gxv[4] = {0.f, 0.f, 0.f, 0.f};
      if (e_ok) {
        const float* gp = a.gx + ((size_t)t * B + b_e) * 4 * H + unit;
#pragma unroll
        for (int g = 0; g < 4; ++g) gxv[g] = __ldg(gp + (size_t)g * H);
      }
      float acc[GC];
#pragma unroll
      for (int c = 0; c < GC; ++c) acc[c] = 0.f;
      for (int kc = 0; kc < H; kc += KC) {
        stage_T(stg, hprev, H, b0, B, kc, H, /*coherent=*/t > 0);
        __syncthreads();
        const int kend = min(KC, H - kc);
        const int k0 = warp * (KC / NW);
#pragma unroll 4
        for (int kk = k0; kk < k0 + KC / NW; ++kk) {
          if (kk < kend) {
            const float hv = stg[kk * 33 + lane];
            const float4* w4 = reinterpret_cast<const float4*>(Ws + (size_t)(kc + kk) * GC);
#pragma unroll
            for (int q = 0; q < GC / 4; ++q) {
              const float4 w = w4[q];
              acc[4 * q + 0] = fmaf(hv, w.x, acc[4 * q + 0]);
              acc[4 * q + 1] = fmaf(hv, w.y, acc[4 * q + 1]);
              acc[4 * q + 2] = fmaf(hv, w.z, acc[4 * q + 2]);
              acc[4 * q + 3] = fmaf(hv, w.w, acc[4 * q + 3]);
            }
          }
        }
        __syncthreads();
      }
#pragma unroll
      for (int c = 0; c < GC; ++c) red[(warp * GC + c) * 33 + lane] = acc[c];
      __syncthreads();
      {
        float g4[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float s = gxv[g];
#pragma unroll
          for (int w = 0; w < NW; ++w) s += red[(w * GC + g * U + eu) * 33 + eb];
          g4[g] = s;
        }
        if (e_ok) {
          const float ig = sigmoidf_(g4[0]);
          const float jg = tanhf_(g4[1]);
          const float fg = sigmoidf_(g4[2] + 1.0f);
          const float og = sigmoidf_(g4[3]);
          const float c_old = cst[b_e * U + eu];
          const float c_new = c_old * fg + ig * jg;
          const float h_new = tanhf_(c_new) * og;
          const bool valid = t < a.len[b_e];
          const size_t row = (size_t)t * B + b_e;
          a.out[row * H + unit] = valid ? h_new : 0.f;
          if (valid) {
            cst[b_e * U + eu] = c_new;
            hst[b_e * U + eu] = h_new;
          }
          if (a.gates) {
            float* gp = a.gates + row * 4 * H + unit;
            gp[0] = ig; gp[(size_t)H] = jg; gp[(size_t)2 * H] = fg; gp[(size_t)3 * H] = og;
            a.cs[row * H + unit] = c_new;
          }
        }
      }
      __syncthreads();
    }
    if (t + 1 < T) grid_arrive(a.barrier);
  }
  for (int e = tid; e < Bpad * U; e += NT) {
    const int b = e / U, u = e % U;
    if (b < B && u0 + u < H) {
      if (a.cT) a.cT[(size_t)b * H + u0 + u] = cst[e];
      if (a.hT) a.hT[(size_t)b * H + u0 + u] = hst[e];
    }
  }
}

// dynamic smem: WsT[4H][U] | stg[KC][33] | red[NW][U][33] | dhs[Bpad][U] | dcs[Bpad][U]
__global__ void __launch_bounds__(NT, 1) lstm_rec_bwd_kernel(RecBwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int T = a.T, B = a.B, H = a.H, G = 4 * a.H;
  const int nbc = (B + 31) / 32, Bpad = nbc * 32;
  float* WsT = smem;
  float* stg = WsT + (size_t)G * U;
  float* red = stg + KC * 33;
  float* dhs = red + NW * U * 33;
  float* dcs = dhs + Bpad * U;
  const int u0 = blockIdx.x * U;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned nctas = gridDim.x;

  // WsT[col][u] = Wh[u0 + u][col]
  for (int e = tid; e < G * U; e += NT) {
    const int u = e / G, col = e % G;     // coalesced over col
    WsT[(size_t)col * U + u] = (u0 + u < H) ? a.Wh[(size_t)(u0 + u) * G + col] : 0.f;
  }
  for (int e = tid; e < Bpad * U; e += NT) { dhs[e] = 0.f; dcs[e] = 0.f; }
  __syncthreads();

  const int eu = tid % U, eb = tid / U;
  unsigned epoch = 0;
  for (int t = T - 1; t >= 0; --t) {
    // ---- elementwise: d(pre-activation gates) for the owned units --------------
    for (int bc = 0; bc < nbc; ++bc) {
      const int b_e = bc * 32 + eb, unit = u0 + eu;
      if (b_e < B && unit < H) {
        const size_t row = (size_t)t * B + b_e;
        float* gp = a.gates + row * G + unit;
        const bool valid = t < a.len[b_e];
        float di = 0.f, dj = 0.f, df = 0.f, dob = 0.f;
        if (valid) {
          const float ig = gp[0], jg = gp[(size_t)H], fg = gp[(size_t)2 * H], og = gp[(size_t)3 * H];
          const float c_t = a.cs[row * H + unit];
          const float c_prev = (t > 0) ? a.cs[(row - B) * H + unit]
                                       : (a.c0 ? a.c0[(size_t)b_e * H + unit] : 0.f);
          const float dh_tot = dhs[b_e * U + eu] + a.dout[row * H + unit];
          const float tc = tanhf_(c_t);
          dob = dh_tot * tc * og * (1.f - og);
          const float dc_tot = dcs[b_e * U + eu] + dh_tot * og * (1.f - tc * tc);
          di = dc_tot * jg * ig * (1.f - ig);
          dj = dc_tot * ig * (1.f - jg * jg);
          df = dc_tot * c_prev * fg * (1.f - fg);
          dcs[b_e * U + eu] = dc_tot * fg;
        }
        // rows with t >= len[b]: no loss term depends on them, dh = dc = 0 there
        gp[0] = di; gp[(size_t)H] = dj; gp[(size_t)2 * H] = df; gp[(size_t)3 * H] = dob;
      }
    }
    if (t == 0) break;     // dh_{-1} is not needed (no gradient into the carried state)
    grid_arrive(a.barrier);
    ++epoch;
    grid_wait(a.barrier, nctas * epoch);
    // ---- dh_{t-1}[b][u] = sum_col dgates_t[b][col] * Wh[u][col] -----------------
    const float* dg = a.gates + (size_t)t * B * G;
    for (int bc = 0; bc < nbc; ++bc) {
      const int b0 = bc * 32;
      float acc[U];
#pragma unroll
      for (int u = 0; u < U; ++u) acc[u] = 0.f;
      for (int kc = 0; kc < G; kc += KC) {
        stage_T(stg, dg, G, b0, B, kc, G, /*coherent=*/true);
        __syncthreads();
        const int kend = min(KC, G - kc);
        const int k0 = warp * (KC / NW);
#pragma unroll 4
        for (int kk = k0; kk < k0 + KC / NW; ++kk) {
          if (kk < kend) {
            const float v = stg[kk * 33 + lane];
            const float4* w4 = reinterpret_cast<const float4*>(WsT + (size_t)(kc + kk) * U);
            const float4 w0 = w4[0], w1 = w4[1];
            acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]);
            acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
            acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]);
            acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
          }
        }
        __syncthreads();
      }
#pragma unroll
      for (int u = 0; u < U; ++u) red[(warp * U + u) * 33 + lane] = acc[u];
      __syncthreads();
      {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += red[(w * U + eu) * 33 + eb];
        dhs[(b0 + eb) * U + eu] = s;
      }
      __syncthreads();
    }
  }
}

size_t fwd_smem_bytes(int H, int B) {
  const int Bpad = (B + 31) / 32 * 32;
  return ((size_t)H * GC + KC * 33 + NW * GC * 33 + 2 * (size_t)Bpad * U) * sizeof(float);
}
size_t bwd_smem_bytes(int H, int B) {
  const int Bpad = (B + 31) / 32 * 32;
  return ((size_t)4 * H * U + KC * 33 + NW * U * 33 + 2 * (size_t)Bpad * U) * sizeof(float);
}

template <typename K, typename A>
int launch_coop(K kernel, const A& args, int grid, size_t smem, cudaStream_t st, const char* name) {
  int dev = 0, coop = 0, nsm = 0, per_sm = 0;
  RS_CHECK_CUDA(cudaGetDevice(&dev));
  RS_CHECK_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  RS_REQUIRE(coop, RS_ERR_UNSUPPORTED, "%s: device lacks cooperative launch", name);
  RS_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RS_CHECK_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  RS_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, NT, smem));
  RS_REQUIRE(per_sm * nsm >= grid, RS_ERR_UNSUPPORTED, "%s: %d CTAs cannot be co-resident (%d SMs x %d)", name,
             grid, nsm, per_sm);
  A a = args;
  void* kargs[] = {(void*)&a};
  RS_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)kernel, dim3(grid), dim3(NT), kargs, smem, st));
  count_launch();
  return RS_OK;
}

}  // namespace

int lstm_rec_forward(const RecFwdArgs& a, cudaStream_t st) {
  RS_REQUIRE(a.H > 0 && a.H <= kRecMaxH, RS_ERR_UNSUPPORTED, "lstm_rec_forward: hidden_size %d > %d", a.H, kRecMaxH);
  RS_REQUIRE(a.B > 0 && a.B <= kRecMaxB, RS_ERR_UNSUPPORTED, "lstm_rec_forward: batch %d > %d", a.B, kRecMaxB);
  RS_REQUIRE(a.T > 0, RS_ERR_INVALID, "lstm_rec_forward: T=%d", a.T);
  RS_CHECK_CUDA(cudaMemsetAsync(a.barrier, 0, sizeof(unsigned), st));
  return launch_coop(lstm_rec_fwd_kernel, a, cdiv(a.H, U), fwd_smem_bytes(a.H, a.B), st, "lstm_rec_forward");
}

int lstm_rec_backward(const RecBwdArgs& a, cudaStream_t st) {
  RS_REQUIRE(a.H > 0 && a.H <= kRecMaxH, RS_ERR_UNSUPPORTED, "lstm_rec_backward: hidden_size %d > %d", a.H, kRecMaxH);
  RS_REQUIRE(a.B > 0 && a.B <= kRecMaxB, RS_ERR_UNSUPPORTED, "lstm_rec_backward: batch %d > %d", a.B, kRecMaxB);
  RS_REQUIRE(a.T > 0, RS_ERR_INVALID, "lstm_rec_backward: T=%d", a.T);
  RS_CHECK_CUDA(cudaMemsetAsync(a.barrier, 0, sizeof(unsigned), st));
  return launch_coop(lstm_rec_bwd_kernel, a, cdiv(a.H, U), bwd_smem_bytes(a.H, a.B), st, "lstm_rec_backward");
}

}  // namespace rs
