// Global-norm clip + TF-Adam update on the flat parameter buffer.
//
// Replaces tf.clip_by_global_norm + AdamOptimizer.apply_gradients
// (/root/reference/models/AcousticModel.py:388,404-406); see oracle/optim.py.
// HBM-bound: rs_sumsq reads g once (4 B/param); rs_clip_adam_step reads g, theta,
// m, v and writes theta, m, v (28 B/param), vectorised 128-bit accesses,
// grid = a multiple of the SM count, grid-stride.
#include "common.cuh"

namespace rs {
namespace {

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, int64_t n,
                                                    double* __restrict__ out) {
  double acc = 0.0;
  const int64_t n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = g4[i];
    acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride)
    acc += (double)g[i] * g[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += red[i];
    atomicAdd(out, s);
  }
}

__device__ __forceinline__ void adam1(float& th, float g, float& m, float& v, float scale, float lr_t, float b1,
                                      float b2, float eps) {
  g *= scale;
  m += (g - m) * (1.0f - b1);
  v += (g * g - v) * (1.0f - b2);
  th -= lr_t * m / (sqrtf(v) + eps);
}

__global__ void __launch_bounds__(256)
clip_adam_kernel(float* __restrict__ th, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 int64_t n, const double* __restrict__ sumsq, float clip, float lr_t, float b1, float b2,
                 float eps) {
  const float norm = (float)sqrt(*sumsq);
  const float scale = clip / fmaxf(norm, clip);      // tf.clip_by_global_norm
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float4* th4 = reinterpret_cast<float4*>(th);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 t = th4[i], gg = g4[i], mm = m4[i], vv = v4[i];
    adam1(t.x, gg.x, mm.x, vv.x, scale, lr_t, b1, b2, eps);
    adam1(t.y, gg.y, mm.y, vv.y, scale, lr_t, b1, b2, eps);
    adam1(t.z, gg.z, mm.z, vv.z, scale, lr_t, b1, b2, eps);
    adam1(t.w, gg.w, mm.w, vv.w, scale, lr_t, b1, b2, eps);
    th4[i] = t; m4[i] = mm; v4[i] = vv;
  }
  for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride)
    adam1(th[i], g[i], m[i], v[i], scale, lr_t, b1, b2, eps);
}

}  // namespace
}  // namespace rs

using namespace rs;

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int rs_sumsq(const float* g_d, int64_t n, double* sumsq_d, void* stream) {
  RS_REQUIRE(g_d && sumsq_d && n > 0, RS_ERR_INVALID, "rs_sumsq: bad argument");
  RS_REQUIRE(al16(g_d), RS_ERR_INVALID, "rs_sumsq: g_d must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  RS_CHECK_CUDA(cudaMemsetAsync(sumsq_d, 0, sizeof(double), st));
  int grid = sm_count() * 8;
  int64_t need = (n / 4 + 255) / 256;
  if (need < 1) need = 1;
  if (grid > need) grid = (int)need;
  sumsq_kernel<<<grid, 256, 0, st>>>(g_d, n, sumsq_d);
  RS_CHECK_LAUNCH();
  return RS_OK;
}

extern "C" int rs_clip_adam_step(float* params_d, const float* grads_d, float* m_d, float* v_d, int64_t n,
                                 const double* sumsq_d, float clip, float lr, float beta1, float beta2, float eps,
                                 int64_t step, void* stream) {
  RS_REQUIRE(params_d && grads_d && m_d && v_d && sumsq_d && n > 0, RS_ERR_INVALID, "rs_clip_adam_step: bad argument");
  RS_REQUIRE(step >= 1, RS_ERR_INVALID, "rs_clip_adam_step: step must be >= 1 (got %lld)", (long long)step);
  RS_REQUIRE(clip > 0.f, RS_ERR_INVALID, "rs_clip_adam_step: clip must be positive");
  RS_REQUIRE(al16(params_d) && al16(grads_d) && al16(m_d) && al16(v_d), RS_ERR_INVALID,
             "rs_clip_adam_step: buffers must be 16-byte aligned");
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)step)) / (1.0 - pow((double)beta1, (double)step));
  int grid = sm_count() * 8;
  int64_t need = (n / 4 + 255) / 256;
  if (need < 1) need = 1;
  if (grid > need) grid = (int)need;
  clip_adam_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(params_d, grads_d, m_d, v_d, n, sumsq_d, clip,
                                                           (float)lr_t, beta1, beta2, eps);
  RS_CHECK_LAUNCH();
  return RS_OK;
}
