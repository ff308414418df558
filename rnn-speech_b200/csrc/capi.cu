// Library-wide pieces of the C ABI: version, error string, device query.
#include "common.cuh"
#include <string.h>
#include <atomic>

namespace rs {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
unsigned long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  return n;
}

}  // namespace rs

namespace rs { unsigned long long launch_count(); }
extern "C" int rs_version(void) { return 100; }
extern "C" uint64_t rs_launch_count(void) { return (uint64_t)rs::launch_count(); }
extern "C" const char* rs_last_error(void) { return rs::g_err; }
extern "C" int rs_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    rs::set_error("rs_sm_count: no CUDA device");
    return RS_ERR_CUDA;
  }
  return n;
}

// MFCC lives in features_mfcc.cu
