// Shared helpers for the rnnspeech_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <math.h>
#include "../../include/rnnspeech_b200.h"

namespace rs {

void set_error(const char* fmt, ...);

#define RS_CHECK_CUDA(expr)                                                             \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      rs::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return RS_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

#define RS_REQUIRE(cond, code, ...)  \
  do {                               \
    if (!(cond)) {                   \
      rs::set_error(__VA_ARGS__);    \
      return (code);                 \
    }                                \
  } while (0)

// every kernel launch of the library goes through this: checks the launch and counts it
// (rs_launch_count() is what bench.py reports as gpu_launches)
void count_launch();
#define RS_CHECK_LAUNCH()              \
  do {                                 \
    rs::count_launch();                \
    RS_CHECK_CUDA(cudaGetLastError()); \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

int sm_count();
// ordinal of the current device, clamped to [0, kMaxDevices): index of the per-device caches of function attributes and
// occupancy results (they are per-device state: a per-process cache would be wrong for a second device in the process)
constexpr int kMaxDevices = 64;
int device_slot();

// ---- dropout mask: counter hash restated bit for bit by oracle/model.py ----
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  uint64_t z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// key = splitmix64(seed); stream_id = 2*layer + (0 input | 1 output);
// idx = (t*B + b)*H + h;  keep iff top 24 bits < keep * 2^24
__host__ __device__ __forceinline__ bool dropout_keep(uint64_t key, uint32_t stream_id, uint64_t idx,
                                                      uint32_t thr24) {
  uint64_t z = splitmix64(key ^ (((uint64_t)stream_id << 40) | idx));
  return (uint32_t)(z >> 40) < thr24;
}

// full-precision libm forms: the parity gate (greedy labels vs an fp32/fp64 CPU
// oracle over ~1000 recurrent steps) does not leave room for ex2.approx error.
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float tanhf_(float x) { return tanhf(x); }

// MUFU-based forms for the latency-critical tensor-core recurrent kernels: ex2.approx /
// rcp.approx, absolute error ~1e-6 on (0,1) / (-1,1) outputs for |x| <= 16.
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) {
  x = fminf(fmaxf(x, -15.0f), 15.0f);
  return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace rs
