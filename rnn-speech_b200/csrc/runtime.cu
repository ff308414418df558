// Small step-protocol pieces of the C ABI, so that a training step launches no framework kernels:
// accumulator zeroing, the display-loss bookkeeping of models/AcousticModel.py:361-383, and the single
// host-to-device copy of a staged mini-batch.
#include "common.cuh"

namespace rs {
namespace {

// dst += (1/n) sum_i v[i] / (div ? div[i] : 1);  *count += 1 when given
// (mean_loss = mean(ctc_loss / seq_len), accumulated over mini-batches: models/AcousticModel.py:361-366, :378-383;
//  a padded row with len 0 gives inf / NaN exactly as the reference's division does)
__global__ void accumulate_mean_kernel(const float* __restrict__ v, const int32_t* __restrict__ div, int n,
                                       float* __restrict__ dst, float* __restrict__ count) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += div ? v[i] / (float)div[i] : v[i];
  acc = warp_sum(acc);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    *dst += s / (float)n;
    if (count) *count += 1.0f;
  }
}

}  // namespace
}  // namespace rs

using namespace rs;

extern "C" int rs_accumulate_mean(const float* v_d, const int32_t* div_d, int n, float* dst_d, float* count_d,
                                  void* stream) {
  RS_REQUIRE(v_d && dst_d && n > 0, RS_ERR_INVALID, "rs_accumulate_mean: bad argument");
  accumulate_mean_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(v_d, div_d, n, dst_d, count_d);
  RS_CHECK_LAUNCH();
  return RS_OK;
}

extern "C" int rs_memset_zero(void* dst_d, size_t bytes, void* stream) {
  RS_REQUIRE(dst_d || bytes == 0, RS_ERR_INVALID, "rs_memset_zero: NULL destination");
  if (bytes) RS_CHECK_CUDA(cudaMemsetAsync(dst_d, 0, bytes, (cudaStream_t)stream));
  return RS_OK;
}

extern "C" int rs_memcpy_h2d_async(void* dst_d, const void* src_host, size_t bytes, void* stream) {
  RS_REQUIRE((dst_d && src_host) || bytes == 0, RS_ERR_INVALID, "rs_memcpy_h2d_async: NULL argument");
  if (bytes) RS_CHECK_CUDA(cudaMemcpyAsync(dst_d, src_host, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return RS_OK;
}
