// Optional batch normalisation of the LSTM stack's input (sm_100a).
//
// Replaces tf.nn.moments(rnn_inputs, [1], keep_dims=True) + tf.nn.batch_normalization(..., None, None, 1e-3)
// (/root/reference/models/AcousticModel.py:253-259; config.ini batch_normalization, default False): for every
// (time step, hidden unit) the mean and POPULATION variance over the batch axis, y = (x - mean) / sqrt(var + 1e-3),
// no scale / offset, padded frames included.  oracle/model.py restates it (forward lines 111-115, backward 205-211).
//
//   bn_forward_kernel   x [T,B,H] -> x_hat in place, 1/sqrt(var+eps) [T,H] kept for backward
//   bn_backward_kernel  d [T,B,H] (gradient wrt x_hat) -> gradient wrt x in place:
//                       dx = istd * (d - mean_b(d) - x_hat * mean_b(d * x_hat))
// One thread per (t, h), consecutive threads on consecutive h (coalesced rows), two passes over the <= 256 batch rows.
#include "common.cuh"

namespace rs {
namespace {

constexpr float kBnEps = 1e-3f;

__global__ void bn_forward_kernel(float* __restrict__ x, float* __restrict__ istd_out, int B, int H) {
  const int t = blockIdx.x, h = blockIdx.y * blockDim.x + threadIdx.x;
  if (h >= H) return;
  float* col = x + (size_t)t * B * H + h;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += col[(size_t)b * H];
  const float mean = s / (float)B;
  float v = 0.f;
  for (int b = 0; b < B; ++b) { const float d = col[(size_t)b * H] - mean; v = fmaf(d, d, v); }
  const float istd = rsqrtf(v / (float)B + kBnEps);
  for (int b = 0; b < B; ++b) col[(size_t)b * H] = (col[(size_t)b * H] - mean) * istd;
  if (istd_out) istd_out[(size_t)t * H + h] = istd;
}

__global__ void bn_backward_kernel(float* __restrict__ d, const float* __restrict__ xhat,
                                   const float* __restrict__ istd_in, int B, int H) {
  const int t = blockIdx.x, h = blockIdx.y * blockDim.x + threadIdx.x;
  if (h >= H) return;
  float* dc = d + (size_t)t * B * H + h;
  const float* xc = xhat + (size_t)t * B * H + h;
  float s1 = 0.f, s2 = 0.f;
  for (int b = 0; b < B; ++b) { const float g = dc[(size_t)b * H]; s1 += g; s2 = fmaf(g, xc[(size_t)b * H], s2); }
  const float m1 = s1 / (float)B, m2 = s2 / (float)B, istd = istd_in[(size_t)t * H + h];
  for (int b = 0; b < B; ++b) dc[(size_t)b * H] = istd * (dc[(size_t)b * H] - m1 - xc[(size_t)b * H] * m2);
}

}  // namespace

int bn_forward(float* x, float* istd, int T, int B, int H, cudaStream_t st) {
  bn_forward_kernel<<<dim3(T, cdiv(H, 128)), 128, 0, st>>>(x, istd, B, H);
  RS_CHECK_LAUNCH();
  return RS_OK;
}

int bn_backward(float* d, const float* xhat, const float* istd, int T, int B, int H, cudaStream_t st) {
  bn_backward_kernel<<<dim3(T, cdiv(H, 128)), 128, 0, st>>>(d, xhat, istd, B, H);
  RS_CHECK_LAUNCH();
  return RS_OK;
}

}  // namespace rs
