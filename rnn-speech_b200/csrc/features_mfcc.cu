// MFCC features for sm_100a.
//
// Replaces AudioProcessor._extract_mfcc = librosa.feature.mfcc(sig, sr, hop_length=round(.01 sr),
// n_fft=round(.025 sr)) (/root/reference/util/audioprocessor.py:63-75): reflect-padded centred
// frames, periodic Hann, an n_fft-point DFT (400 at 16 kHz -- not a power of two, so a direct
// DFT against a shared-memory twiddle table), power, 128 Slaney mel bands (area-normalised),
// 10 log10 with amin 1e-10 and the utterance-wide top_db = 80 clamp, orthonormal DCT-II, first
// n_mfcc coefficients.  librosa is absent from the reference tree: restated from its published
// algorithm (oracle/features.py::mfcc), parity unpinned upstream.
//
//   mfcc_logmel_kernel  one warp per frame: windowed frame -> smem, lanes over DFT bins,
//                       power -> smem, lanes over mel bands -> 10 log10 -> workspace; per-CTA max
//   mfcc_dct_kernel     utterance max (from the per-CTA maxima) -> clamp -> DCT -> [.,.,n_mfcc]
// fp32 arithmetic (librosa's own pipeline is float32 / complex64); tables built in double on host.
#include "common.cuh"
#include <string.h>
#include <mutex>
#include <map>
#include <vector>

namespace rs {
namespace {

constexpr int kNmels = 128;
constexpr int kMaxNfft = 1024;
constexpr int kWarps = 8;
constexpr double kPi = 3.14159265358979323846;

struct MfccTables {
  int n_fft, hop, nbins;
  std::vector<float> window, cosw, sinw;   // [n_fft]
  std::vector<float> melw;                 // [128][nbins]
  std::vector<int> mstart, mend;           // [128]
  std::vector<float> dct;                  // [128 (k)][128 (n)] orthonormal DCT-II rows
};

double hz_to_mel(double f) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
double mel_to_hz(double m) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

const MfccTables* get_tables(int sr) {
  static std::mutex mu;
  static std::map<int, MfccTables*> cache;
  std::lock_guard<std::mutex> g(mu);
  auto it = cache.find(sr);
  if (it != cache.end()) return it->second;
  MfccTables* t = new MfccTables();
  t->n_fft = (int)nearbyint(0.025 * sr);
  t->hop = (int)nearbyint(0.01 * sr);
  const int N = t->n_fft, nb = N / 2 + 1;
  t->nbins = nb;
  t->window.resize(N); t->cosw.resize(N); t->sinw.resize(N);
  for (int i = 0; i < N; ++i) {
    t->window[i] = (float)(0.5 - 0.5 * cos(2.0 * kPi * i / N));     // periodic Hann
    t->cosw[i] = (float)cos(2.0 * kPi * i / N);
    t->sinw[i] = (float)sin(2.0 * kPi * i / N);
  }
  // librosa.filters.mel(sr, n_fft, n_mels=128, fmin=0, fmax=sr/2, htk=False, norm=1)
  std::vector<double> mel_f(kNmels + 2);
  const double m0 = hz_to_mel(0.0), m1 = hz_to_mel(sr / 2.0);
  for (int i = 0; i < kNmels + 2; ++i) mel_f[i] = mel_to_hz(m0 + (m1 - m0) * i / (kNmels + 1));
  t->melw.assign((size_t)kNmels * nb, 0.f);
  t->mstart.assign(kNmels, nb); t->mend.assign(kNmels, 0);
  for (int m = 0; m < kNmels; ++m) {
    const double enorm = 2.0 / (mel_f[m + 2] - mel_f[m]);
    for (int k = 0; k < nb; ++k) {
      const double f = (double)sr / 2.0 * k / (nb - 1);
      const double lower = (f - mel_f[m]) / (mel_f[m + 1] - mel_f[m]);
      const double upper = (mel_f[m + 2] - f) / (mel_f[m + 2] - mel_f[m + 1]);
      const double w = fmax(0.0, fmin(lower, upper)) * enorm;
      if (w > 0.0) {
        t->melw[(size_t)m * nb + k] = (float)w;
        if (k < t->mstart[m]) t->mstart[m] = k;
        if (k + 1 > t->mend[m]) t->mend[m] = k + 1;
      }
    }
    if (t->mend[m] == 0) t->mstart[m] = 0;
  }
  t->dct.resize((size_t)kNmels * kNmels);
  for (int k = 0; k < kNmels; ++k)
    for (int n = 0; n < kNmels; ++n) {
      double v = cos(kPi * k * (2 * n + 1) / (2.0 * kNmels)) * sqrt(2.0 / kNmels);
      if (k == 0) v *= 1.0 / sqrt(2.0);
      t->dct[(size_t)k * kNmels + n] = (float)v;
    }
  cache[sr] = t;
  return t;
}

struct DevTables {   // offsets (in floats) inside the workspace table block
  const float *window, *cosw, *sinw, *melw, *dct;
  const int *mstart, *mend;
};

// grid (ceil(Tfull / 8), B); dynamic smem: cos[N] sin[N] win[N] | per warp: x[N] pw[nbins]
__global__ void __launch_bounds__(kWarps * 32)
mfcc_logmel_kernel(const float* __restrict__ pcm, const int64_t* __restrict__ offsets, DevTables tab, int N, int hop,
                   int nbins, int Tstride, int nblk, float* __restrict__ logmel, float* __restrict__ blkmax,
                   int* __restrict__ nframes_out) {
  extern __shared__ float sm[];
  float* scos = sm;
  float* ssin = scos + N;
  float* swin = ssin + N;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sx = swin + N + warp * (N + nbins);
  float* spw = sx + N;
  __shared__ float wmax[kWarps];
  const int b = blockIdx.y;
  const int64_t off = offsets[b];
  const int64_t n = offsets[b + 1] - off;
  const int T = (int)(1 + n / hop);
  if (blockIdx.x == 0 && threadIdx.x == 0) nframes_out[b] = T;
  for (int i = threadIdx.x; i < N; i += blockDim.x) { scos[i] = tab.cosw[i]; ssin[i] = tab.sinw[i]; swin[i] = tab.window[i]; }
  __syncthreads();
  const int t = blockIdx.x * kWarps + warp;
  float vmax = -INFINITY;
  if (t < T) {
    const float* x = pcm + off;
    const int pad = N / 2;
    for (int i = lane; i < N; i += 32) {
      int64_t s = (int64_t)t * hop + i - pad;            // np.pad(mode='reflect')
      if (s < 0) s = -s;
      if (s >= n) s = 2 * (n - 1) - s;
      s = s < 0 ? 0 : (s >= n ? n - 1 : s);
      sx[i] = x[s] * swin[i];
    }
    __syncwarp();
    for (int k = lane; k < nbins; k += 32) {
      float re = 0.f, im = 0.f;
      int idx = 0;
      for (int i = 0; i < N; ++i) {
        const float v = sx[i];
        re = fmaf(v, scos[idx], re);
        im = fmaf(v, ssin[idx], im);
        idx += k;
        if (idx >= N) idx -= N;
      }
      spw[k] = re * re + im * im;
    }
    __syncwarp();
    for (int m = lane; m < kNmels; m += 32) {
      const float* w = tab.melw + (size_t)m * nbins;
      float acc = 0.f;
      const int k1 = tab.mend[m];
      for (int k = tab.mstart[m]; k < k1; ++k) acc = fmaf(spw[k], w[k], acc);
      const float db = 10.0f * log10f(fmaxf(acc, 1e-10f));   // power_to_db(ref=1, amin=1e-10)
      logmel[((size_t)b * Tstride + t) * kNmels + m] = db;
      vmax = fmaxf(vmax, db);
    }
  }
  vmax = warp_max(vmax);
  if (lane == 0) wmax[warp] = vmax;
  __syncthreads();
  if (threadIdx.x == 0) {
    float mx = -INFINITY;
    for (int w = 0; w < kWarps; ++w) mx = fmaxf(mx, wmax[w]);
    blkmax[(size_t)b * nblk + blockIdx.x] = mx;
  }
}

// grid (ceil(Tmax / 8), B), one warp per output frame
__global__ void __launch_bounds__(kWarps * 32)
mfcc_dct_kernel(const float* __restrict__ logmel, const float* __restrict__ blkmax, const int* __restrict__ nframes,
                DevTables tab, int Tstride, int nblk, int Tmax, int B, int n_mfcc, int time_major, float top_db,
                float* __restrict__ out) {
  __shared__ float srow[kWarps][kNmels];
  __shared__ float smax;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int T = nframes[b];
  if (threadIdx.x < 32) {
    float mx = -INFINITY;
    for (int i = lane; i < nblk; i += 32) mx = fmaxf(mx, blkmax[(size_t)b * nblk + i]);
    mx = warp_max(mx);
    if (lane == 0) smax = mx;
  }
  __syncthreads();
  const int t = blockIdx.x * kWarps + warp;
  if (t >= Tmax) return;
  float* o = time_major ? out + ((size_t)t * B + b) * n_mfcc : out + ((size_t)b * Tmax + t) * n_mfcc;
  if (t >= T) {
    for (int k = lane; k < n_mfcc; k += 32) o[k] = 0.f;
    return;
  }
  const float floor_db = smax - top_db;
  for (int m = lane; m < kNmels; m += 32) srow[warp][m] = fmaxf(logmel[((size_t)b * Tstride + t) * kNmels + m], floor_db);
  __syncwarp();
  for (int k = lane; k < n_mfcc; k += 32) {
    const float* d = tab.dct + (size_t)k * kNmels;
    float acc = 0.f;
#pragma unroll 8
    for (int m = 0; m < kNmels; ++m) acc = fmaf(srow[warp][m], d[m], acc);
    o[k] = acc;
  }
}

size_t table_floats(int N, int nbins) { return (size_t)3 * N + (size_t)kNmels * nbins + (size_t)kNmels * kNmels + 2 * kNmels; }

}  // namespace
}  // namespace rs

using namespace rs;

extern "C" int64_t rs_mfcc_num_frames(int64_t n, int sr) {
  const int hop = (int)nearbyint(0.01 * sr);
  return hop > 0 ? 1 + n / hop : 0;
}

extern "C" size_t rs_mfcc_workspace_bytes(int B, int64_t max_samples, int sr) {
  const int N = (int)nearbyint(0.025 * sr), nbins = N / 2 + 1;
  const int64_t Tfull = rs_mfcc_num_frames(max_samples, sr);
  const int64_t nblk = (Tfull + kWarps - 1) / kWarps;
  return align_up(table_floats(N, nbins) * sizeof(float), 256) + align_up((size_t)B * Tfull * kNmels * sizeof(float), 256) +
         align_up((size_t)B * nblk * sizeof(float), 256);
}

extern "C" int rs_mfcc_forward(const float* pcm_d, const int64_t* offsets_d, int B, int64_t max_samples, int sr,
                               int Tmax, int n_mfcc, int time_major, float* out_d, int32_t* nframes_d, void* ws_d,
                               size_t ws_bytes, void* stream) {
  RS_REQUIRE(B > 0 && Tmax > 0 && sr > 0 && max_samples > 0, RS_ERR_INVALID, "rs_mfcc_forward: bad arguments");
  RS_REQUIRE(n_mfcc > 0 && n_mfcc <= kNmels, RS_ERR_INVALID, "rs_mfcc_forward: n_mfcc %d outside [1,128]", n_mfcc);
  const MfccTables* t = get_tables(sr);
  const int N = t->n_fft, nbins = t->nbins;
  RS_REQUIRE(N >= 2 && N <= kMaxNfft && t->hop > 0, RS_ERR_UNSUPPORTED, "rs_mfcc_forward: n_fft %d unsupported (sr %d)", N, sr);
  RS_REQUIRE(max_samples > N / 2, RS_ERR_INVALID, "rs_mfcc_forward: signal shorter than the reflect padding");
  RS_REQUIRE(ws_bytes >= rs_mfcc_workspace_bytes(B, max_samples, sr), RS_ERR_WORKSPACE, "rs_mfcc_forward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int Tfull = (int)rs_mfcc_num_frames(max_samples, sr);
  const int nblk = cdiv(Tfull, kWarps);
  // tables -> workspace
  std::vector<float> host(table_floats(N, nbins));
  float* p = host.data();
  memcpy(p, t->window.data(), N * 4); p += N;
  memcpy(p, t->cosw.data(), N * 4); p += N;
  memcpy(p, t->sinw.data(), N * 4); p += N;
  memcpy(p, t->melw.data(), (size_t)kNmels * nbins * 4); p += (size_t)kNmels * nbins;
  memcpy(p, t->dct.data(), (size_t)kNmels * kNmels * 4); p += (size_t)kNmels * kNmels;
  memcpy(p, t->mstart.data(), kNmels * 4); p += kNmels;
  memcpy(p, t->mend.data(), kNmels * 4);
  char* ws = (char*)ws_d;
  float* tb = (float*)ws;
  RS_CHECK_CUDA(cudaMemcpyAsync(tb, host.data(), host.size() * 4, cudaMemcpyHostToDevice, st));
  RS_CHECK_CUDA(cudaStreamSynchronize(st));   // `host` is a local buffer
  DevTables d;
  d.window = tb; d.cosw = tb + N; d.sinw = tb + 2 * N; d.melw = tb + 3 * N;
  d.dct = d.melw + (size_t)kNmels * nbins;
  d.mstart = (const int*)(d.dct + (size_t)kNmels * kNmels);
  d.mend = d.mstart + kNmels;
  float* logmel = (float*)(ws + align_up(table_floats(N, nbins) * sizeof(float), 256));
  float* blkmax = (float*)((char*)logmel + align_up((size_t)B * Tfull * kNmels * sizeof(float), 256));
  const size_t smem = ((size_t)3 * N + (size_t)kWarps * (N + nbins)) * sizeof(float);
  RS_CHECK_CUDA(cudaFuncSetAttribute(mfcc_logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mfcc_logmel_kernel<<<dim3(nblk, B), kWarps * 32, smem, st>>>(pcm_d, offsets_d, d, N, t->hop, nbins, Tfull, nblk,
                                                               logmel, blkmax, nframes_d);
  RS_CHECK_LAUNCH();
  mfcc_dct_kernel<<<dim3(cdiv(Tmax, kWarps), B), kWarps * 32, 0, st>>>(logmel, blkmax, nframes_d, d, Tfull, nblk, Tmax, B,
                                                                      n_mfcc, time_major, 80.0f, out_d);
  RS_CHECK_LAUNCH();
  return RS_OK;
}
