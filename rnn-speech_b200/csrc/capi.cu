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

int device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  return dev % kMaxDevices;
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

// CRC-32C (Castagnoli) of a HOST buffer, continuing from `crc` (0 to start): the checksum of TensorFlow's tensor-bundle
// files (table blocks and tensor payloads), used by rnn-speech_b200/tf_checkpoint.py::write_bundle.  Slicing-by-8.
namespace {
struct Crc32cTable {
  uint32_t v[8][256];
  Crc32cTable() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      v[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int t = 1; t < 8; ++t) v[t][i] = (v[t - 1][i] >> 8) ^ v[0][v[t - 1][i] & 0xff];
  }
};
}  // namespace

extern "C" uint32_t rs_crc32c(const void* data, size_t n, uint32_t crc) {
  static const Crc32cTable tab;                     // initialised once, thread-safe (C++11 static local)
  const uint32_t (*table)[256] = tab.v;
  const unsigned char* p = (const unsigned char*)data;
  uint32_t c = ~crc;
  while (n >= 8) {
    uint32_t lo, hi;
    memcpy(&lo, p, 4);
    memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = table[7][lo & 0xff] ^ table[6][(lo >> 8) & 0xff] ^ table[5][(lo >> 16) & 0xff] ^ table[4][lo >> 24] ^
        table[3][hi & 0xff] ^ table[2][(hi >> 8) & 0xff] ^ table[1][(hi >> 16) & 0xff] ^ table[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) c = (c >> 8) ^ table[0][(c ^ *p++) & 0xff];
  return ~c;
}

// MFCC lives in features_mfcc.cu
