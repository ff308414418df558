// Persistent warp-specialised tcgen05 GEMM (TMA -> smem ring -> tcgen05.mma -> TMEM ->
// epilogue warps), used for every batched projection of the acoustic model:
//   gx = xin @ K[:H] (+b), logits = top @ w_o (+b), rnn_in = x @ w_i (+b), dxin = dgates @ K[:H]^T.
// One CTA per SM, 128 x 128 output tiles, BK = 64, 3-stage ring, two TMEM accumulator
// stages so that the epilogue of tile i overlaps the MMAs of tile i+1.
//   warp 0: TMA producer   warp 1: MMA issuer   warp 2: TMEM allocator   warps 4-7: epilogue
// fp32-grade accuracy from bf16 tensor cores: products = 3 issues A_hi*B_hi + A_hi*B_lo +
// A_lo*B_hi into the same accumulator (see tc_common.cuh).
#include "gemm_tc.cuh"
#include "tc_common.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace rs {
namespace {

// Output tile 128 x BN, BN = 256 or 128.  A tcgen05.mma with both operands in shared memory reads
// (128 + BN) x 16 bf16 per instruction; at BN = 128 that is 128 B per clock -- all of the shared-memory
// bandwidth -- so the 128-wide tile cannot reach the tensor pipe's rate and the 256-wide one (96 B per
// clock) is the default.  BN = 128 (3 stages) stays for narrow outputs.
constexpr int BM = 128, BK = 64;
constexpr int TILE_A_BYTES = BM * BK * 2;   // 16 KB
constexpr int NTHREADS = 256;
// (BN, stages): (128, 3) and (256, 2) fill an SM's shared memory; (128, 2) leaves room -- 129 KB, 256 TMEM columns --
// to share an SM with a backward recurrent CTA (50 KB, 256 columns, 128 registers), whose tensor pipe is ~5 % busy.
template <int BN, int ST> struct Cfg {
  static constexpr int STAGES = ST;
  static constexpr int TILE_B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = 2 * TILE_A_BYTES + 2 * TILE_B_BYTES;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

struct KParams {
  int M, N, K, products;
  int tiles_m, tiles_n;
  int base_ctas;            // elastic launches: CTAs that work when the machine is shared
  GemmTcOut out;
};

// MN = true: both operands MN-major ("TN" product C = A^T B of two matrices stored [K][M] and [K][N] row-major --
// the weight-gradient GEMMs x^T dgates, whose operands the other kernels write with K = (t, b) as the ROW index):
// every 64-column block of an operand tile is its own TMA box {64 mn, 64 k}.
template <int BN, int ST, bool MN>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, KParams p) {
  constexpr int STAGES = Cfg<BN, ST>::STAGES, TILE_B_BYTES = Cfg<BN, ST>::TILE_B_BYTES, STAGE_BYTES = Cfg<BN, ST>::STAGE_BYTES;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = p.tiles_m * p.tiles_n;
  const int nkb = (p.K + BK - 1) / BK;
  // how many CTAs of this grid work (all of them, or the first base_ctas of an elastic launch): decided once per
  // launch by whichever CTA comes first, so that every CTA strides over the tiles by the same count
  __shared__ int s_nctas;
  int nctas = gridDim.x;
  if (p.out.elastic) {
    if (threadIdx.x == 0) {
      const int want = (*reinterpret_cast<volatile int*>(p.out.elastic) != 0) ? (int)gridDim.x : p.base_ctas;
      const int old = atomicCAS(p.out.elastic + 1 + p.out.elastic_id, 0, want);
      s_nctas = old ? old : want;
    }
    __syncthreads();
    nctas = s_nctas;
    if ((int)blockIdx.x >= nctas) return;
  }
  const bool use_blo = (p.products & 1) != 0, use_alo = (p.products & 2) != 0;   // extra terms A_hi*B_lo / A_lo*B_hi

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA_hi); tc::tma_prefetch_desc(&tmB_hi);
    if (use_alo) tc::tma_prefetch_desc(&tmA_lo);
    if (use_blo) tc::tma_prefetch_desc(&tmB_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&tfull_bar[s], 1); tc::mbar_init(&tempty_bar[s], 4); }
    tc::fence_mbar_init();
  }
  if (warp == 2) tc::tmem_alloc(&tmem_slot, 2 * BN);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    {   // converged warp; one elected lane issues each TMA / arrive
      uint32_t it = 0;
      const uint32_t tx = (uint32_t)(TILE_A_BYTES + TILE_B_BYTES + (use_alo ? TILE_A_BYTES : 0) + (use_blo ? TILE_B_BYTES : 0));
      for (int tile = blockIdx.x; tile < ntiles; tile += nctas) {
        const int m0 = (tile / p.tiles_n) * BM, n0 = (tile % p.tiles_n) * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          tc::mbar_wait(&empty_bar[s], ph ^ 1);
          unsigned char* st = smem + (size_t)s * STAGE_BYTES;
          tc::mbar_arrive_expect_tx_warp(&full_bar[s], tx);
          if constexpr (MN) {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) {
              tc::tma_load_2d_warp(st + i * 8192, &tmA_hi, m0 + 64 * i, kb * BK, &full_bar[s]);
              if (use_alo) tc::tma_load_2d_warp(st + TILE_A_BYTES + i * 8192, &tmA_lo, m0 + 64 * i, kb * BK, &full_bar[s]);
            }
#pragma unroll
            for (int i = 0; i < BN / 64; ++i) {
              tc::tma_load_2d_warp(st + 2 * TILE_A_BYTES + i * 8192, &tmB_hi, n0 + 64 * i, kb * BK, &full_bar[s]);
              if (use_blo) tc::tma_load_2d_warp(st + 2 * TILE_A_BYTES + TILE_B_BYTES + i * 8192, &tmB_lo, n0 + 64 * i, kb * BK, &full_bar[s]);
            }
          } else {
          tc::tma_load_2d_warp(st, &tmA_hi, kb * BK, m0, &full_bar[s]);
          tc::tma_load_2d_warp(st + 2 * TILE_A_BYTES, &tmB_hi, kb * BK, n0, &full_bar[s]);
          if (use_alo) tc::tma_load_2d_warp(st + TILE_A_BYTES, &tmA_lo, kb * BK, m0, &full_bar[s]);
          if (use_blo) tc::tma_load_2d_warp(st + 2 * TILE_A_BYTES + TILE_B_BYTES, &tmB_lo, kb * BK, n0, &full_bar[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    {   // converged warp, elect.sync inside each issue
      const uint32_t idesc = MN ? tc::instr_desc_bf16_mn(BM, BN) : tc::instr_desc_bf16(BM, BN);
      uint32_t it = 0, tl = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += nctas, ++tl) {
        const int as = tl & 1;
        tc::mbar_wait(&tempty_bar[as], ((tl >> 1) & 1) ^ 1);
        tc::tc_fence_after();
        const uint32_t d = tmem + (uint32_t)(as * BN);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          tc::mbar_wait(&full_bar[s], ph);
          tc::tc_fence_after();
          const uint32_t sa = tc::smem_u32(smem + (size_t)s * STAGE_BYTES);
          // K-major: 16 K elements = 32 bytes along the row; MN-major: 16 K rows = two 8-row groups = 2048 bytes
          constexpr uint32_t KSTEP = MN ? (2048 >> 4) : 2;
          const uint64_t dah = MN ? tc::smem_desc_sw128_mn(sa, 8192, 1024) : tc::smem_desc_sw128(sa);
          const uint64_t dal = MN ? tc::smem_desc_sw128_mn(sa + TILE_A_BYTES, 8192, 1024) : tc::smem_desc_sw128(sa + TILE_A_BYTES);
          const uint64_t dbh = MN ? tc::smem_desc_sw128_mn(sa + 2 * TILE_A_BYTES, 8192, 1024) : tc::smem_desc_sw128(sa + 2 * TILE_A_BYTES);
          const uint64_t dbl = MN ? tc::smem_desc_sw128_mn(sa + 2 * TILE_A_BYTES + TILE_B_BYTES, 8192, 1024)
                                  : tc::smem_desc_sw128(sa + 2 * TILE_A_BYTES + TILE_B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            tc::mma_bf16_ss_warp(d, dah + KSTEP * k, dbh + KSTEP * k, idesc, (uint32_t)((kb | k) != 0));
            if (use_blo) tc::mma_bf16_ss_warp(d, dah + KSTEP * k, dbl + KSTEP * k, idesc, 1u);
            if (use_alo) tc::mma_bf16_ss_warp(d, dal + KSTEP * k, dbh + KSTEP * k, idesc, 1u);
          }
          tc::mma_commit_warp(&empty_bar[s]);
        }
        tc::mma_commit_warp(&tfull_bar[as]);
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;                       // TMEM lane quadrant of this warp
    const GemmTcOut& o = p.out;
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += nctas, ++tl) {
      const int as = tl & 1;
      const int m0 = (tile / p.tiles_n) * BM, n0 = (tile % p.tiles_n) * BN;
      tc::mbar_wait(&tfull_bar[as], (tl >> 1) & 1);
      tc::tc_fence_after();
      const int m = m0 + q * 32 + lane;
      const bool row_ok = m < p.M;
      // GEMM_OUT_REC: this thread's row is a permuted gate row; its bias entry is g*H + unit
      float rec_bias = 0.f;
      if (o.mode == GEMM_OUT_REC && row_ok && o.bias) {
        const int r4u = 4 * o.recU, sl = m / r4u, r = m - sl * r4u;
        rec_bias = __ldg(o.bias + (size_t)(r & 3) * o.recH + sl * o.recU + (r >> 2));
      }
#pragma unroll 1
      for (int c = 0; c < BN / 16; ++c) {
        float v[16];
        tc::tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c * 16), v);
        tc::tmem_ld_wait();
        const int nb = n0 + c * 16;
        if (row_ok && nb < p.N) {
        if (o.mode == GEMM_OUT_REC) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += rec_bias;
        } else if (o.bias) {
#pragma unroll
          for (int i = 0; i < 16; ++i) if (nb + i < p.N) v[i] += __ldg(o.bias + nb + i);
        }
        if (o.mode == GEMM_OUT_F32) {
          float* cp = o.C + (size_t)m * o.ldc + nb;
          if (nb + 15 < p.N && ((reinterpret_cast<uintptr_t>(cp) & 15) == 0)) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              float4 x = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
              if (o.accumulate) {
                const float4 y = *reinterpret_cast<const float4*>(cp + i);
                x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
              }
              *reinterpret_cast<float4*>(cp + i) = x;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (nb + i < p.N) cp[i] = o.accumulate ? cp[i] + v[i] : v[i];
          }
        } else if (o.mode == GEMM_OUT_REC) {
          // column n = t*B + b -> gx[(t*4H + m)*Bpad + b]: 16 consecutive b of one step are 64 contiguous bytes
          const int t = nb / o.recB, b0 = nb - t * o.recB;
          float* cp = o.C + ((size_t)t * 4 * o.recH + m) * (size_t)o.recBpad + b0;
          if (b0 + 15 < o.recB && nb + 15 < p.N && ((b0 | o.recBpad) & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(cp + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int n = nb + i;
              if (n < p.N) {
                const int tt = n / o.recB, bb = n - tt * o.recB;
                o.C[((size_t)tt * 4 * o.recH + m) * (size_t)o.recBpad + bb] = v[i];
              }
            }
          }
        } else {
          if (o.drop_thr != 0u && o.drop_thr != 0xffffffffu) {
            const uint64_t e0 = (uint64_t)m * (uint64_t)p.N + (uint64_t)nb;
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = dropout_keep(o.drop_key, o.drop_stream, e0 + i, o.drop_thr) ? v[i] * o.drop_inv : 0.f;
          }
          __nv_bfloat16* hp = o.Chi + (size_t)m * o.ldc + nb;
          __nv_bfloat16* lp = o.Clo + (size_t)m * o.ldc + nb;
          if (nb + 15 < p.N && ((reinterpret_cast<uintptr_t>(hp) & 15) == 0)) {
            uint32_t h[8], l[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              __nv_bfloat16 h0, l0, h1, l1;
              tc::split_bf16(v[2 * i], h0, l0);
              tc::split_bf16(v[2 * i + 1], h1, l1);
              h[i] = tc::pack_bf16(h0, h1);
              l[i] = tc::pack_bf16(l0, l1);
            }
            *reinterpret_cast<uint4*>(hp) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(hp + 8) = make_uint4(h[4], h[5], h[6], h[7]);
            *reinterpret_cast<uint4*>(lp) = make_uint4(l[0], l[1], l[2], l[3]);
            *reinterpret_cast<uint4*>(lp + 8) = make_uint4(l[4], l[5], l[6], l[7]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (nb + i < p.N) {
                __nv_bfloat16 h0, l0;
                tc::split_bf16(v[i], h0, l0);
                hp[i] = h0; lp[i] = l0;
              }
          }
        }
        }
        __syncwarp();
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty_bar[as]);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem, 2 * BN);
}

__global__ void split_planes_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ hi,
                                    __nv_bfloat16* __restrict__ lo, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    __nv_bfloat16 h, l;
    tc::split_bf16(in[i], h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
  }
}

// 32x32 smem tile transpose; in [R,C] -> out [C,R]
__global__ void split_planes_T_kernel(const float* __restrict__ in, int R, int C, int ld_in,
                                      __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int ld_out) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? in[(size_t)r * ld_in + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < R) {
      __nv_bfloat16 h, l;
      tc::split_bf16(tile[threadIdx.x][i], h, l);
      hi[(size_t)c * ld_out + r] = h;
      if (lo) lo[(size_t)c * ld_out + r] = l;
    }
  }
}

__global__ void transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, int R, int C, int ld_in,
                                      __nv_bfloat16* __restrict__ out, int ld_out) {
  __shared__ __nv_bfloat16 tile[64][66];
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  for (int i = threadIdx.y; i < 64; i += blockDim.y)
    for (int jj = threadIdx.x; jj < 64; jj += 32) {
      const int r = r0 + i, c = c0 + jj;
      tile[i][jj] = (r < R && c < C) ? in[(size_t)r * ld_in + c] : __float2bfloat16(0.f);
    }
  __syncthreads();
  for (int i = threadIdx.y; i < 64; i += blockDim.y)
    for (int jj = threadIdx.x; jj < 64; jj += 32) {
      const int c = c0 + i, r = r0 + jj;
      if (c < C && r < R) out[(size_t)c * ld_out + r] = tile[jj][i];
    }
}

// one warp per row
__global__ void rowsum_planes_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, int R,
                                     int C, int ld, float* __restrict__ out, int accumulate) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) {
    s += __bfloat162float(hi[(size_t)row * ld + c]);
    if (lo) s += __bfloat162float(lo[(size_t)row * ld + c]);
  }
  s = warp_sum(s);
  if (lane == 0) out[row] = accumulate ? out[row] + s : s;
}

// Column sums of a (hi, lo) plane pair, bit-reproducible: block (x, y) sums a slab of rows of 64 columns (thread (cx, ry):
// 8 columns -- one 16-byte load per plane and row --, rows ry, ry + 32, ... of the slab), writes the 64 partial sums to
// scratch[y][C], and the block that arrives last at the column block's counter adds the slabs in slab order: one writer
// per column, a fixed summation order whatever the block schedule.  (One block per 64 columns over ALL rows -- the first
// version -- took 172 us per 4096 x 3072 call, 10 % of a step's kernel time: too little memory parallelism.)
constexpr int kColsumSlabs = 8;
__global__ void __launch_bounds__(256) colsum_planes_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                                            int R, int C, int ld, float* __restrict__ out, int accumulate,
                                                            float* __restrict__ scratch, unsigned* __restrict__ counters) {
  __shared__ float red[32][65];
  __shared__ unsigned s_ticket;
  const int cx = threadIdx.x & 7, ry = threadIdx.x >> 3;
  const int c = blockIdx.x * 64 + 8 * cx;
  const int rows_per = (R + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per, r1 = min(R, r0 + rows_per);
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.f;
  if (c + 7 < C) {
#pragma unroll 4
    for (int r = r0 + ry; r < r1; r += 32) {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(hi + (size_t)r * ld + c));
      const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) { s[2 * i] += __uint_as_float(w[i] << 16); s[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u); }
      if (lo) {
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(lo + (size_t)r * ld + c));
        const uint32_t v[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { s[2 * i] += __uint_as_float(v[i] << 16); s[2 * i + 1] += __uint_as_float(v[i] & 0xffff0000u); }
      }
    }
  } else {
    for (int r = r0 + ry; r < r1; r += 32)
      for (int i = 0; i < 8; ++i)
        if (c + i < C) {
          s[i] += __bfloat162float(hi[(size_t)r * ld + c + i]);
          if (lo) s[i] += __bfloat162float(lo[(size_t)r * ld + c + i]);
        }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[ry][8 * cx + i] = s[i];
  __syncthreads();
  const int cc = blockIdx.x * 64 + threadIdx.x;
  if (threadIdx.x < 64 && cc < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += red[i][threadIdx.x];
    scratch[(size_t)blockIdx.y * C + cc] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(&counters[blockIdx.x], 1u);
  __syncthreads();
  if (s_ticket != gridDim.y - 1) return;
  __threadfence();
  if (threadIdx.x < 64 && cc < C) {
    float t = 0.f;
    for (unsigned y = 0; y < gridDim.y; ++y) t += __ldcg(scratch + (size_t)y * C + cc);
    out[cc] = accumulate ? out[cc] + t : t;
  }
  if (threadIdx.x == 0) counters[blockIdx.x] = 0u;          // ready for the next call on this scratch
}

int make_tmap(CUtensorMap* map, const __nv_bfloat16* base, int rows, int cols, int ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  RS_REQUIRE(fn != nullptr, RS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  RS_REQUIRE((ld % 8) == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0, RS_ERR_INVALID,
             "TMA operand needs ld %% 8 == 0 and a 16-byte aligned base (ld=%d)", ld);
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RS_REQUIRE(r == CUDA_SUCCESS, RS_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d (rows=%d cols=%d ld=%d)", (int)r,
             rows, cols, ld);
  return RS_OK;
}

}  // namespace

// MN-major operand: the matrix is [krows][cols] row-major (cols = the M / N index, contiguous); box {64 cols, 64 k rows}
int make_tmap_mn(CUtensorMap* map, const __nv_bfloat16* base, int krows, int cols, int ld) {
  EncodeTiledFn fn = encode_fn();
  RS_REQUIRE(fn != nullptr, RS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  RS_REQUIRE((ld % 8) == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0, RS_ERR_INVALID,
             "TMA operand needs ld %% 8 == 0 and a 16-byte aligned base (ld=%d)", ld);
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)krows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)BK};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RS_REQUIRE(r == CUDA_SUCCESS, RS_ERR_CUDA, "cuTensorMapEncodeTiled (mn) failed with %d (krows=%d cols=%d ld=%d)", (int)r,
             krows, cols, ld);
  return RS_OK;
}

int tmap_2d_bf16(void* map64, const __nv_bfloat16* base, int rows, int cols, int ld, int box_rows) {
  return make_tmap(reinterpret_cast<CUtensorMap*>(map64), base, rows, cols, ld, box_rows);
}

// Plain (unswizzled) 2-D map for TMA stores of a [box_rows x box_cols] shared-memory tile.
int tmap_store_bf16(void* map64, const __nv_bfloat16* base, int rows, int cols, int ld, int box_rows, int box_cols) {
  EncodeTiledFn fn = encode_fn();
  RS_REQUIRE(fn != nullptr, RS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  RS_REQUIRE((ld % 8) == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0 && (box_cols % 8) == 0, RS_ERR_INVALID,
             "TMA store map needs ld %% 8 == 0, box_cols %% 8 == 0 and a 16-byte aligned base");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(map64), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(base),
                  gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RS_REQUIRE(r == CUDA_SUCCESS, RS_ERR_CUDA, "cuTensorMapEncodeTiled (store) failed with %d", (int)r);
  return RS_OK;
}

// Plain 3-D store map: dims {d0, d1, d2} elements with byte strides {2, s1, s2}; box {b0, b1, b2}.
int tmap_store3_bf16(void* map64, const __nv_bfloat16* base, int d0, int d1, int d2, size_t s1_bytes, size_t s2_bytes,
                     int b0, int b1, int b2) {
  EncodeTiledFn fn = encode_fn();
  RS_REQUIRE(fn != nullptr, RS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t gdim[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t gstr[2] = {(cuuint64_t)s1_bytes, (cuuint64_t)s2_bytes};
  cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(map64), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(base),
                  gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RS_REQUIRE(r == CUDA_SUCCESS, RS_ERR_CUDA, "cuTensorMapEncodeTiled (store3) failed with %d", (int)r);
  return RS_OK;
}

// 4-D map over a row-interleaved plane pair [rows][2][cols] bf16 (hi | lo per row): one box
// {64 k, box_rows, 2 planes, box_kb K-blocks} lands as box_kb stacked SWIZZLE_128B operand tiles
// [kb][plane][box_rows][64] -- the stacked (hi rows; lo rows) B operand of the TMEM-resident kernels.
int tmap_stacked_bf16(void* map64, const __nv_bfloat16* base, int rows, int cols, int box_rows, int box_kb) {
  EncodeTiledFn fn = encode_fn();
  RS_REQUIRE(fn != nullptr, RS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  RS_REQUIRE((cols % 64) == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0, RS_ERR_INVALID,
             "stacked TMA map needs cols %% 64 == 0 and a 16-byte aligned base");
  cuuint64_t gdim[4] = {64, (cuuint64_t)rows, 2, (cuuint64_t)(cols / 64)};
  cuuint64_t gstr[3] = {(cuuint64_t)cols * 2 * 2, (cuuint64_t)cols * 2, 128};
  cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 2, (cuuint32_t)box_kb};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(map64), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(base),
                  gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RS_REQUIRE(r == CUDA_SUCCESS, RS_ERR_CUDA, "cuTensorMapEncodeTiled (stacked) failed with %d (rows=%d cols=%d)", (int)r, rows, cols);
  return RS_OK;
}

// The same stacked box over two separate plane arrays [rows][ld] (hi at base_hi, lo plane_stride_bytes later).
int tmap_stacked2_bf16(void* map64, const __nv_bfloat16* base_hi, size_t plane_stride_bytes, int rows, int cols, int ld,
                       int box_rows, int box_kb) {
  EncodeTiledFn fn = encode_fn();
  RS_REQUIRE(fn != nullptr, RS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  RS_REQUIRE((cols % 64) == 0 && (ld % 8) == 0 && (reinterpret_cast<uintptr_t>(base_hi) & 15) == 0 &&
             (plane_stride_bytes % 16) == 0 && plane_stride_bytes > 0, RS_ERR_INVALID,
             "stacked TMA map needs cols %% 64 == 0, ld %% 8 == 0, 16-byte aligned planes");
  cuuint64_t gdim[4] = {64, (cuuint64_t)rows, 2, (cuuint64_t)(cols / 64)};
  cuuint64_t gstr[3] = {(cuuint64_t)ld * 2, (cuuint64_t)plane_stride_bytes, 128};
  cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 2, (cuuint32_t)box_kb};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(map64), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(base_hi),
                  gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RS_REQUIRE(r == CUDA_SUCCESS, RS_ERR_CUDA, "cuTensorMapEncodeTiled (stacked2) failed with %d (rows=%d cols=%d)", (int)r, rows, cols);
  return RS_OK;
}

namespace {

template <int BN, int ST, bool MN>
int gemm_launch(const SplitMat& A, const SplitMat& B, int M, int N, int K, int products, const GemmTcOut& out,
                cudaStream_t st) {
  // bf16x3 degrades gracefully to the planes that exist: hi*hi (+ hi*B_lo) (+ A_lo*hi)
  products = (products == 3 ? ((B.lo ? 1 : 0) | (A.lo ? 2 : 0)) : 0);
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  int rc;
  if (MN) {
    if ((rc = make_tmap_mn(&ta_hi, A.hi, K, M, A.ld)) != RS_OK) return rc;
    if ((rc = make_tmap_mn(&tb_hi, B.hi, K, N, B.ld)) != RS_OK) return rc;
  } else {
    if ((rc = make_tmap(&ta_hi, A.hi, M, K, A.ld, BM)) != RS_OK) return rc;
    if ((rc = make_tmap(&tb_hi, B.hi, N, K, B.ld, BN)) != RS_OK) return rc;
  }
  ta_lo = ta_hi; tb_lo = tb_hi;
  if (products & 2) if ((rc = MN ? make_tmap_mn(&ta_lo, A.lo, K, M, A.ld) : make_tmap(&ta_lo, A.lo, M, K, A.ld, BM)) != RS_OK) return rc;
  if (products & 1) if ((rc = MN ? make_tmap_mn(&tb_lo, B.lo, K, N, B.ld) : make_tmap(&tb_lo, B.lo, N, K, B.ld, BN)) != RS_OK) return rc;
  KParams p;
  p.M = M; p.N = N; p.K = K; p.products = products;
  p.tiles_m = cdiv(M, BM); p.tiles_n = cdiv(N, BN);
  p.out = out;
  const int ntiles = p.tiles_m * p.tiles_n;
  int grid = sm_count();
  if (out.max_ctas > 0 && grid > out.max_ctas) grid = out.max_ctas;
  if (out.tiles_per_cta > 0) grid = cdiv(ntiles, out.tiles_per_cta);
  if (grid > ntiles) grid = ntiles;
  p.base_ctas = grid;
  if (out.elastic) { grid = sm_count(); if (grid > ntiles) grid = ntiles; if (grid < p.base_ctas) grid = p.base_ctas; }
  const size_t smem = (size_t)Cfg<BN, ST>::STAGES * Cfg<BN, ST>::STAGE_BYTES + 1024;
  static bool attr_done[kMaxDevices] = {};
  const int dev = device_slot();
  if (!attr_done[dev]) {
    RS_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, ST, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done[dev] = true;
  }
  gemm_tc_kernel<BN, ST, MN><<<grid, NTHREADS, smem, st>>>(ta_hi, ta_lo, tb_hi, tb_lo, p);
  RS_CHECK_LAUNCH();
  return RS_OK;
}

}  // namespace

static int g_force_bn = -1;     // test hook override of RS_GEMM_BN

int gemm_tc_nt(const SplitMat& A, const SplitMat& B, int M, int N, int K, int products, const GemmTcOut& out,
               cudaStream_t st) {
  if (M <= 0 || N <= 0) return RS_OK;
  RS_REQUIRE(K > 0, RS_ERR_INVALID, "gemm_tc_nt: K=%d", K);
  static const int env_bn = [] { const char* v = getenv("RS_GEMM_BN"); return v ? atoi(v) : 0; }();
  const int force_bn = g_force_bn >= 0 ? g_force_bn : env_bn;
  // 128 x 256 tiles read a quarter fewer operand bytes per flop, but only pay off while there are at least as many
  // of them as CTAs (measured: tests/gpu_diag.py gemmbench -- the 768 x 3072 weight-gradient GEMMs prefer 128 x 128)
  int ctas = sm_count();
  if (out.max_ctas > 0 && out.max_ctas < ctas && !out.elastic) ctas = out.max_ctas;
  const bool wide = force_bn ? force_bn == 256 : (N > 128 && cdiv(M, BM) * cdiv(N, 256) >= ctas);
  if (out.coresident) return gemm_launch<128, 2, false>(A, B, M, N, K, products, out, st);
  return wide ? gemm_launch<256, 2, false>(A, B, M, N, K, products, out, st) : gemm_launch<128, 3, false>(A, B, M, N, K, products, out, st);
}

// C[M,N] = A^T B with A stored [K][M] and B stored [K][N] (row-major, ld in elements): both operands MN-major.
int gemm_tc_tn(const SplitMat& A, const SplitMat& B, int M, int N, int K, int products, const GemmTcOut& out,
               cudaStream_t st) {
  if (M <= 0 || N <= 0) return RS_OK;
  RS_REQUIRE(K > 0, RS_ERR_INVALID, "gemm_tc_tn: K=%d", K);
  RS_REQUIRE(out.mode == GEMM_OUT_F32 && !out.coresident, RS_ERR_INVALID, "gemm_tc_tn: fp32 output only");
  static const int env_bn = [] { const char* v = getenv("RS_GEMM_BN"); return v ? atoi(v) : 0; }();
  const int force_bn = g_force_bn >= 0 ? g_force_bn : env_bn;
  int ctas = sm_count();
  if (out.max_ctas > 0 && out.max_ctas < ctas && !out.elastic) ctas = out.max_ctas;
  const bool wide = force_bn ? force_bn == 256 : (N > 128 && cdiv(M, BM) * cdiv(N, 256) >= ctas);
  return wide ? gemm_launch<256, 2, true>(A, B, M, N, K, products, out, st) : gemm_launch<128, 3, true>(A, B, M, N, K, products, out, st);
}

int split_planes(const float* in, __nv_bfloat16* hi, __nv_bfloat16* lo, int64_t n, cudaStream_t st) {
  if (n <= 0) return RS_OK;
  int grid = (int)((n + 255) / 256);
  const int cap = sm_count() * 16;
  if (grid > cap) grid = cap;
  split_planes_kernel<<<grid, 256, 0, st>>>(in, hi, lo, n);
  RS_CHECK_LAUNCH();
  return RS_OK;
}

int split_planes_transposed(const float* in, int R, int C, int ld_in, __nv_bfloat16* hi, __nv_bfloat16* lo, int ld_out,
                            cudaStream_t st) {
  if (R <= 0 || C <= 0) return RS_OK;
  split_planes_T_kernel<<<dim3(cdiv(C, 32), cdiv(R, 32)), dim3(32, 8), 0, st>>>(in, R, C, ld_in, hi, lo, ld_out);
  RS_CHECK_LAUNCH();
  return RS_OK;
}

int transpose_bf16(const __nv_bfloat16* in, int R, int C, int ld_in, __nv_bfloat16* out, int ld_out, cudaStream_t st) {
  if (R <= 0 || C <= 0) return RS_OK;
  transpose_bf16_kernel<<<dim3(cdiv(C, 64), cdiv(R, 64)), dim3(32, 8), 0, st>>>(in, R, C, ld_in, out, ld_out);
  RS_CHECK_LAUNCH();
  return RS_OK;
}

// scratch layout: [kColsumCounters column-block counters][kColsumSlabs x C partial sums].  The counters come FIRST, at a
// place that does not depend on C: calls with different C share one scratch, and a counter must never sit where another
// call's partial sums go.
constexpr int kColsumCounters = 1024;
size_t colsum_scratch_bytes(int C) { return ((size_t)kColsumCounters + (size_t)kColsumSlabs * C) * sizeof(float); }

int colsum_planes(const __nv_bfloat16* hi, const __nv_bfloat16* lo, int R, int C, int ld, float* out, int accumulate,
                  void* scratch, cudaStream_t st) {
  if (C <= 0) return RS_OK;
  RS_REQUIRE((ld % 8) == 0 && (reinterpret_cast<uintptr_t>(hi) & 15) == 0 && (!lo || (reinterpret_cast<uintptr_t>(lo) & 15) == 0),
             RS_ERR_INVALID, "colsum_planes: planes must be 16-byte aligned with a row stride that is a multiple of 8");
  RS_REQUIRE(scratch != nullptr, RS_ERR_INVALID, "colsum_planes: scratch (colsum_scratch_bytes, counters zeroed) is required");
  RS_REQUIRE(cdiv(C, 64) <= kColsumCounters, RS_ERR_UNSUPPORTED, "colsum_planes: %d columns", C);
  unsigned* counters = reinterpret_cast<unsigned*>(scratch);
  float* part = reinterpret_cast<float*>(scratch) + kColsumCounters;
  const int slabs = R >= 64 * kColsumSlabs ? kColsumSlabs : 1;
  colsum_planes_kernel<<<dim3(cdiv(C, 64), slabs), 256, 0, st>>>(hi, lo, R, C, ld, out, accumulate, part, counters);
  RS_CHECK_LAUNCH();
  return RS_OK;
}

int rowsum_planes(const __nv_bfloat16* hi, const __nv_bfloat16* lo, int R, int C, int ld, float* out, int accumulate,
                  cudaStream_t st) {
  if (R <= 0) return RS_OK;
  rowsum_planes_kernel<<<cdiv(R, 8), 256, 0, st>>>(hi, lo, R, C, ld, out, accumulate);
  RS_CHECK_LAUNCH();
  return RS_OK;
}

}  // namespace rs

using namespace rs;

#ifdef RS_DIAG   // test / measurement hooks: in librnnspeech_b200_diag.so only
// Test hook: C[M,N] = A[M,K] * B[N,K]^T (+bias) from fp32 inputs; scratch_d must hold
// 2*(M*Kp + N*Kp) bf16 with Kp = K rounded up to 8.
// products >= 16 selects the MN-major form (A_d is [K][M], B_d is [K][N], M and N multiples of 8; products - 16 = 1 | 3).
extern "C" int rs_gemm_tc_test(const float* A_d, const float* B_d, const float* bias_d, float* C_d, int M, int N,
                               int K, int products, void* scratch_d, size_t scratch_bytes, void* stream) {
  RS_REQUIRE(A_d && B_d && C_d && scratch_d, RS_ERR_INVALID, "rs_gemm_tc_test: NULL argument");
  if (products >= 16) {
    RS_REQUIRE(M % 8 == 0 && N % 8 == 0, RS_ERR_INVALID, "rs_gemm_tc_test (MN-major): M and N must be multiples of 8");
    const size_t need_mn = 2 * ((size_t)M * K + (size_t)N * K) * sizeof(__nv_bfloat16) + 64;
    RS_REQUIRE(scratch_bytes >= need_mn, RS_ERR_WORKSPACE, "rs_gemm_tc_test: scratch %zu < %zu", scratch_bytes, need_mn);
    cudaStream_t st2 = (cudaStream_t)stream;
    __nv_bfloat16* ah2 = (__nv_bfloat16*)scratch_d;
    __nv_bfloat16* al2 = ah2 + (size_t)M * K;
    __nv_bfloat16* bh2 = al2 + (size_t)M * K;
    __nv_bfloat16* bl2 = bh2 + (size_t)N * K;
    int rc2;
    if ((rc2 = split_planes(A_d, ah2, al2, (int64_t)M * K, st2)) != RS_OK) return rc2;
    if ((rc2 = split_planes(B_d, bh2, bl2, (int64_t)N * K, st2)) != RS_OK) return rc2;
    SplitMat A2{ah2, al2, K, M, M}, B2{bh2, bl2, K, N, N};
    GemmTcOut o2{};
    o2.mode = GEMM_OUT_F32; o2.C = C_d; o2.ldc = N; o2.bias = bias_d; o2.accumulate = 0;
    return gemm_tc_tn(A2, B2, M, N, K, products - 16, o2, st2);
  }
  RS_REQUIRE(K % 8 == 0, RS_ERR_INVALID, "rs_gemm_tc_test: K must be a multiple of 8");
  const size_t need = 2 * ((size_t)M * K + (size_t)N * K) * sizeof(__nv_bfloat16) + 64;
  RS_REQUIRE(scratch_bytes >= need, RS_ERR_WORKSPACE, "rs_gemm_tc_test: scratch %zu < %zu", scratch_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* ah = (__nv_bfloat16*)scratch_d;
  __nv_bfloat16* al = ah + (size_t)M * K;
  __nv_bfloat16* bh = al + (size_t)M * K;
  __nv_bfloat16* bl = bh + (size_t)N * K;
  int rc;
  if ((rc = split_planes(A_d, ah, al, (int64_t)M * K, st)) != RS_OK) return rc;
  if ((rc = split_planes(B_d, bh, bl, (int64_t)N * K, st)) != RS_OK) return rc;
  SplitMat A{ah, al, M, K, K}, B{bh, bl, N, K, K};
  GemmTcOut o{};
  o.mode = GEMM_OUT_F32; o.C = C_d; o.ldc = N; o.bias = bias_d; o.accumulate = 0;
  return gemm_tc_nt(A, B, M, N, K, products, o, st);
}

// Measurement hook: split once, then time `reps` GEMMs C = A B^T (K-major fp32 inputs) with CUDA events on `stream`.
// bn = 0 | 128 | 256 forces the tile width; max_ctas / tiles_per_cta / accumulate as in GemmTcOut.
extern "C" int rs_gemm_tc_bench(const float* A_d, const float* B_d, float* C_d, int M, int N, int K, int products,
                                int bn, int max_ctas, int tiles_per_cta, int accumulate, int reps, void* scratch_d,
                                size_t scratch_bytes, float* ms_out, void* stream) {
  RS_REQUIRE(A_d && B_d && C_d && scratch_d && ms_out && reps > 0, RS_ERR_INVALID, "rs_gemm_tc_bench: bad argument");
  RS_REQUIRE(K % 8 == 0, RS_ERR_INVALID, "rs_gemm_tc_bench: K must be a multiple of 8");
  const size_t need = 2 * ((size_t)M * K + (size_t)N * K) * sizeof(__nv_bfloat16) + 64;
  RS_REQUIRE(scratch_bytes >= need, RS_ERR_WORKSPACE, "rs_gemm_tc_bench: scratch %zu < %zu", scratch_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* ah = (__nv_bfloat16*)scratch_d;
  __nv_bfloat16* al = ah + (size_t)M * K;
  __nv_bfloat16* bh = al + (size_t)M * K;
  __nv_bfloat16* bl = bh + (size_t)N * K;
  int rc;
  if ((rc = split_planes(A_d, ah, al, (int64_t)M * K, st)) != RS_OK) return rc;
  if ((rc = split_planes(B_d, bh, bl, (int64_t)N * K, st)) != RS_OK) return rc;
  SplitMat A{ah, al, M, K, K}, B{bh, bl, N, K, K};
  GemmTcOut o{};
  o.mode = GEMM_OUT_F32; o.C = C_d; o.ldc = N; o.accumulate = accumulate; o.max_ctas = max_ctas; o.tiles_per_cta = tiles_per_cta;
  cudaEvent_t e0, e1;
  RS_CHECK_CUDA(cudaEventCreate(&e0));
  RS_CHECK_CUDA(cudaEventCreate(&e1));
  g_force_bn = bn;
  rc = gemm_tc_nt(A, B, M, N, K, products, o, st);          // warm-up
  RS_CHECK_CUDA(cudaEventRecord(e0, st));
  for (int i = 0; i < reps && rc == RS_OK; ++i) rc = gemm_tc_nt(A, B, M, N, K, products, o, st);
  RS_CHECK_CUDA(cudaEventRecord(e1, st));
  g_force_bn = -1;
  if (rc != RS_OK) return rc;
  RS_CHECK_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  RS_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  *ms_out = ms / reps;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return RS_OK;
}
#endif  // RS_DIAG
