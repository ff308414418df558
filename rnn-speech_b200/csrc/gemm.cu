// Generic fp32 GEMM (FFMA) + column sums.  Any shape, any alignment; the large
// aligned shapes of the hot path are routed to the tcgen05 kernels instead.
#include "gemm.cuh"

namespace rs {

namespace {
constexpr int BM = 128, BN = 128, BK = 16;

__device__ __forceinline__ float4 ld4_guard(const float* __restrict__ base, int r, int c, int R, int Cc,
                                            int ld, bool vec_ok) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r >= R) return v;
  const float* p = base + (size_t)r * ld + c;
  if (vec_ok && c + 3 < Cc) return *reinterpret_cast<const float4*>(p);
  if (c < Cc) v.x = p[0];
  if (c + 1 < Cc) v.y = p[1];
  if (c + 2 < Cc) v.z = p[2];
  if (c + 3 < Cc) v.w = p[3];
  return v;
}

template <int TA, int TB>
__global__ void __launch_bounds__(256)
sgemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
             float* __restrict__ C, int ldc, const float* __restrict__ bias, int accumulate, int vecA,
             int vecB, int vecC) {
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int e = tid + 256 * r;
      if (TA == 0) {   // A [M,K]: 4 consecutive k of one row
        const int row = e >> 2, kq = (e & 3) * 4;
        ra[r] = ld4_guard(A, m0 + row, k0 + kq, M, K, lda, vecA);
      } else {         // A [K,M]: 4 consecutive m of one k
        const int krow = e >> 5, mq = (e & 31) * 4;
        ra[r] = ld4_guard(A, k0 + krow, m0 + mq, K, M, lda, vecA);
      }
      if (TB == 0) {   // B [K,N]
        const int krow = e >> 5, nq = (e & 31) * 4;
        rb[r] = ld4_guard(B, k0 + krow, n0 + nq, K, N, ldb, vecB);
      } else {         // B [N,K]
        const int col = e >> 2, kq = (e & 3) * 4;
        rb[r] = ld4_guard(B, n0 + col, k0 + kq, N, K, ldb, vecB);
      }
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int e = tid + 256 * r;
      if (TA == 0) {
        const int row = e >> 2, kq = (e & 3) * 4;
        As[buf][kq + 0][row] = ra[r].x; As[buf][kq + 1][row] = ra[r].y;
        As[buf][kq + 2][row] = ra[r].z; As[buf][kq + 3][row] = ra[r].w;
      } else {
        const int krow = e >> 5, mq = (e & 31) * 4;
        *reinterpret_cast<float4*>(&As[buf][krow][mq]) = ra[r];
      }
      if (TB == 0) {
        const int krow = e >> 5, nq = (e & 31) * 4;
        *reinterpret_cast<float4*>(&Bs[buf][krow][nq]) = rb[r];
      } else {
        const int col = e >> 2, kq = (e & 3) * 4;
        Bs[buf][kq + 0][col] = rb[r].x; Bs[buf][kq + 1][col] = rb[r].y;
        Bs[buf][kq + 2][col] = rb[r].z; Bs[buf][kq + 3][col] = rb[r].w;
      }
    }
  };

  const int nk = (K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int n = n0 + jh * 64 + tx * 4;
      if (n >= N) continue;
      float v[4] = {acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]};
      float* cp = C + (size_t)m * ldc + n;
      if (vecC && n + 3 < N) {
        if (bias) {
          const float4 bb = *reinterpret_cast<const float4*>(bias + n);
          v[0] += bb.x; v[1] += bb.y; v[2] += bb.z; v[3] += bb.w;
        }
        if (accumulate) {
          const float4 cc = *reinterpret_cast<const float4*>(cp);
          v[0] += cc.x; v[1] += cc.y; v[2] += cc.z; v[3] += cc.w;
        }
        *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (n + q < N) {
            float x = v[q];
            if (bias) x += bias[n + q];
            if (accumulate) x += cp[q];
            cp[q] = x;
          }
        }
      }
    }
  }
}

// grid: ceil(N/32) CTAs, block (32, 8)
__global__ void colsum_kernel(const float* __restrict__ A, int M, int N, int lda, float* __restrict__ out,
                              int accumulate) {
  __shared__ float red[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (n < N)
    for (int m = threadIdx.y; m < M; m += 8) s += A[(size_t)m * lda + n];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    out[n] = accumulate ? out[n] + t : t;
  }
}
}  // namespace

int sgemm(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
          float* C, int ldc, const float* bias, int accumulate, cudaStream_t st) {
  if (M <= 0 || N <= 0) return RS_OK;
  RS_REQUIRE(K > 0, RS_ERR_INVALID, "sgemm: K=%d", K);
  dim3 grid(cdiv(N, BN), cdiv(M, BM));
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const int vecA = (lda % 4 == 0) && al16(A), vecB = (ldb % 4 == 0) && al16(B);
  const int vecC = (ldc % 4 == 0) && al16(C) && (!bias || al16(bias));
  if (!transA && !transB)
    sgemm_kernel<0, 0><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, accumulate, vecA, vecB, vecC);
  else if (transA && !transB)
    sgemm_kernel<1, 0><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, accumulate, vecA, vecB, vecC);
  else if (!transA && transB)
    sgemm_kernel<0, 1><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, accumulate, vecA, vecB, vecC);
  else
    sgemm_kernel<1, 1><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, accumulate, vecA, vecB, vecC);
  RS_CHECK_LAUNCH();
  return RS_OK;
}

int colsum(const float* A, int M, int N, int lda, float* out, int accumulate, cudaStream_t st) {
  if (N <= 0) return RS_OK;
  colsum_kernel<<<cdiv(N, 32), dim3(32, 8), 0, st>>>(A, M, N, lda, out, accumulate);
  RS_CHECK_LAUNCH();
  return RS_OK;
}

}  // namespace rs
