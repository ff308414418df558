// Fused feature extraction for sm_100a.
//
// Replaces AudioProcessor._extract_fbank / _extract_mfcc
// (/root/reference/util/audioprocessor.py:63-161) for a whole batch per call.
//
// fbank (120-dim):
//   fbank_logmel_kernel  PCM tile -> shared memory (coalesced, read once) ->
//                        pre-emphasis + Hamming -> two real frames packed into one
//                        512-point complex FFT per warp (shared-memory radix-2
//                        butterflies) -> |X|^2/512 -> 40 HTK-mel triangles -> 10 log10
//                        -> log-mel [B,T,40] in the workspace + per-CTA column sums
//   fbank_delta_kernel   per-utterance mean (from the partial sums) -> mean-norm ->
//                        delta -> delta-delta -> [.,.,120] (batch- or time-major),
//                        zero fill up to Tmax
// The reference computes in float64; these kernels compute in fp32 with tables
// built in double on the host (tolerances are stated in tests/test_features_gpu.py).
#include "common.cuh"
#include <mutex>
#include <vector>
#include <map>

namespace rs {

static constexpr int kNfft = 512;
static constexpr int kBins = kNfft / 2 + 1;  // 257
static constexpr int kNfilt = 40;
static constexpr int kFramesPerCta = 32;
static constexpr int kFbankWarps = 8;
static constexpr double kPi = 3.14159265358979323846;

// python3 round(): round half to even
static inline int py_round(double x) {
  double r = nearbyint(x);  // default rounding mode = to nearest even
  return (int)r;
}

struct FrameParams { int frame_length, frame_step; };
static inline FrameParams frame_params(int sr) {
  FrameParams p;
  p.frame_length = py_round(0.025 * sr);
  p.frame_step = py_round(0.01 * sr);
  return p;
}
static inline int64_t fbank_num_frames(int64_t n, const FrameParams& p) {
  int64_t d = n - p.frame_length;
  if (d < 0) d = -d;
  return (d + p.frame_step - 1) / p.frame_step;   // ceil(|n - fl| / step)
}

// ---- host-built tables (double math, then rounded to fp32) ------------------
struct FbankTables {
  double window[kNfft];         // Hamming(frame_length) cropped / zero-padded to 512
  double tw_re[kNfft / 2];      // exp(-2 pi i k / 512)
  double tw_im[kNfft / 2];
  double melw[kNfilt * kBins];  // dense [40][257]
  int mstart[kNfilt];
  int mend[kNfilt];
};

static void build_fbank_tables(int sr, FbankTables* t) {
  FrameParams fp = frame_params(sr);
  for (int i = 0; i < kNfft; ++i) {
    double w = 0.0;
    if (i < fp.frame_length) {
      w = (fp.frame_length == 1) ? 1.0 : 0.54 - 0.46 * cos(2.0 * kPi * i / (fp.frame_length - 1));
    }
    t->window[i] = w;
  }
  for (int k = 0; k < kNfft / 2; ++k) {
    t->tw_re[k] = cos(-2.0 * kPi * k / kNfft);
    t->tw_im[k] = sin(-2.0 * kPi * k / kNfft);
  }
  // util/audioprocessor.py:107-133
  double high_mel = 2595.0 * log10(1.0 + ((double)sr / 2.0) / 700.0);
  double bins[kNfilt + 2];
  for (int i = 0; i < kNfilt + 2; ++i) {
    // np.linspace(0, high_mel, 42): start + i*step with step = (stop-start)/41, last point exact
    double mel = (i == kNfilt + 1) ? high_mel : i * (high_mel / (kNfilt + 1));
    double hz = 700.0 * (pow(10.0, mel / 2595.0) - 1.0);
    bins[i] = floor((kNfft + 1) * hz / sr);
  }
  for (int i = 0; i < kNfilt * kBins; ++i) t->melw[i] = 0.0;
  for (int m = 1; m <= kNfilt; ++m) {
    int lo = (int)bins[m - 1], ce = (int)bins[m], hi = (int)bins[m + 1];
    for (int k = lo; k < ce && k < kBins; ++k)
      t->melw[(m - 1) * kBins + k] = (k - bins[m - 1]) / (bins[m] - bins[m - 1]);
    for (int k = ce; k < hi && k < kBins; ++k)
      t->melw[(m - 1) * kBins + k] = (bins[m + 1] - k) / (bins[m + 1] - bins[m]);
    t->mstart[m - 1] = lo < kBins ? lo : kBins;
    t->mend[m - 1] = hi < kBins ? hi : kBins;
  }
}

static const FbankTables* get_fbank_tables(int sr) {
  static std::mutex mu;
  static std::map<int, FbankTables*> cache;
  std::lock_guard<std::mutex> g(mu);
  auto it = cache.find(sr);
  if (it != cache.end()) return it->second;
  FbankTables* t = new FbankTables();
  build_fbank_tables(sr, t);
  cache[sr] = t;
  return t;
}

// ---- kernels ------------------------------------------------------------------

__device__ __forceinline__ int bitrev9(int x) { return (int)(__brev((unsigned)x) >> 23); }

// grid (ceil(Tfull_max / 32), B), 256 threads.
// The reference computes this stage in float64 (its zero padding promotes the signal,
// util/audioprocessor.py:96-97) after a float32 pre-emphasis; so does this kernel: the
// FFT, power, mel sums and log10 run in fp64 (0.4 GFLOP per batch of 32 -- negligible
// against B200's fp64 rate), which keeps log-mel of weak bins exact instead of carrying
// fp32 FFT leakage error.
// dynamic smem: float pcm[span + 1]; double fft[8 warps][2][512]; double pw[8][2][260]; double col[8][40]
__global__ void __launch_bounds__(kFbankWarps * 32)
fbank_logmel_kernel(const float* __restrict__ pcm, const int64_t* __restrict__ offsets,
                    const FbankTables* __restrict__ tab, int frame_length, int frame_step,
                    int Tstride, int nblk, double* __restrict__ logmel, double* __restrict__ partial,
                    int* __restrict__ nframes_out) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int b = blockIdx.y;
  const int64_t off = offsets[b];
  const int64_t n = offsets[b + 1] - off;
  int64_t d = n - frame_length; if (d < 0) d = -d;
  const int T = (int)((d + frame_step - 1) / frame_step);
  if (blockIdx.x == 0 && threadIdx.x == 0) nframes_out[b] = T;
  const int t0 = blockIdx.x * kFramesPerCta;
  double* part = partial + ((size_t)b * nblk + blockIdx.x) * kNfilt;
  if (t0 >= T) {
    if (threadIdx.x < kNfilt) part[threadIdx.x] = 0.0;
    return;
  }
  const int nfr = min(kFramesPerCta, T - t0);
  const int nuse = min(frame_length, kNfft);       // rfft(frames, 512) crops or zero-pads
  const int span = (kFramesPerCta - 1) * frame_step + nuse;
  double* sfft = reinterpret_cast<double*>(sm_raw);              // [8][2][512]
  double* scol = sfft + kFbankWarps * 2 * kNfft;                 // [8][40] per-warp column sums
  float* spcm = reinterpret_cast<float*>(scol + kFbankWarps * kNfilt);   // [span + 1], spcm[i] = x[s0 - 1 + i]
  const int64_t s0 = (int64_t)t0 * frame_step;
  const float* x = pcm + off;
  for (int i = threadIdx.x; i < span + 1; i += blockDim.x) {
    int64_t s = s0 - 1 + i;
    spcm[i] = (s >= 0 && s < n) ? x[s] : 0.f;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = lane; i < kNfilt; i += 32) scol[warp * kNfilt + i] = 0.0;
  __syncthreads();

  double* re = sfft + warp * 2 * kNfft;
  double* im = re + kNfft;
  const double* __restrict__ win = tab->window;
  const double* __restrict__ twr = tab->tw_re;
  const double* __restrict__ twi = tab->tw_im;

  for (int pair = warp; pair * 2 < nfr; pair += kFbankWarps) {
    const int fa = pair * 2, fb = fa + 1;
    const bool has_b = fb < nfr;
    // load (float32 pre-emphasis, float64 window) in bit-reversed order: frame A -> re, frame B -> im
    for (int i = lane; i < kNfft; i += 32) {
      double va = 0.0, vb = 0.0;
      if (i < nuse) {
        const double w = win[i];
        {
          const int p = fa * frame_step + i + 1;         // index into spcm (shifted by one)
          const int64_t s = s0 + fa * frame_step + i;    // absolute sample
          // y[0] = x[0]; y[s] = x[s] - 0.97 x[s-1] in float32 with separately rounded product
          // (numpy: sig[1:] - 0.97 * sig[:-1] on a float32 array); zero beyond the signal
          const float prev = (s > 0) ? spcm[p - 1] : 0.f;
          const float y = (s < n) ? __fsub_rn(spcm[p], __fmul_rn(0.97f, prev)) : 0.f;
          va = (double)y * w;
        }
        if (has_b) {
          const int p = fb * frame_step + i + 1;
          const int64_t s = s0 + fb * frame_step + i;
          const float prev = (s > 0) ? spcm[p - 1] : 0.f;
          const float y = (s < n) ? __fsub_rn(spcm[p], __fmul_rn(0.97f, prev)) : 0.f;
          vb = (double)y * w;
        }
      }
      const int r = bitrev9(i);
      re[r] = va;
      im[r] = vb;
    }
    __syncwarp();
    // 9 radix-2 DIT stages
#pragma unroll 1
    for (int s = 1; s <= 9; ++s) {
      const int half = 1 << (s - 1);
      const int tstep = kNfft >> s;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int i = lane + 32 * r;
        const int pos = i & (half - 1);
        const int a = ((i >> (s - 1)) << s) + pos;
        const int bb = a + half;
        const double wr = twr[pos * tstep], wi = twi[pos * tstep];
        const double xr = re[bb], xi = im[bb];
        const double tr = wr * xr - wi * xi, ti = wr * xi + wi * xr;
        const double ar = re[a], ai = im[a];
        re[a] = ar + tr; im[a] = ai + ti;
        re[bb] = ar - tr; im[bb] = ai - ti;
      }
      __syncwarp();
    }
    // unpack the two real spectra, power = |X|^2 / 512.  Bin k only needs the FFT outputs k and 512 - k, and no other
    // bin needs them, so the two power spectra overwrite re[0..256] / im[0..256] in place (no second buffer: the
    // CTA's shared memory drops from 122 KB to 89 KB and two CTAs fit on an SM).
    for (int k = lane; k < kBins; k += 32) {
      const int nk = (kNfft - k) & (kNfft - 1);
      const double zr = re[k], zi = im[k], yr = re[nk], yi = im[nk];
      const double ar = 0.5 * (zr + yr), ai = 0.5 * (zi - yi);
      const double br = 0.5 * (zi + yi), bi = -0.5 * (zr - yr);
      re[k] = (ar * ar + ai * ai) * (1.0 / kNfft);
      im[k] = (br * br + bi * bi) * (1.0 / kNfft);
    }
    __syncwarp();
    const double* pw = re;
    const double* pwb = im;
    // 40 triangular filters; one lane owns filter m for BOTH frames of the pair so the
    // per-warp column sums are accumulated in a fixed order (deterministic).
    for (int m = lane; m < kNfilt; m += 32) {
      const double* w = tab->melw + m * kBins;
      double acc_a = 0.0, acc_b = 0.0;
      const int k1 = tab->mend[m];
      for (int k = tab->mstart[m]; k < k1; ++k) {
        const double wk = w[k];
        acc_a = fma(pw[k], wk, acc_a);
        acc_b = fma(pwb[k], wk, acc_b);
      }
      if (acc_a == 0.0) acc_a = 2.220446049250313e-16;   // util/audioprocessor.py:135
      if (acc_b == 0.0) acc_b = 2.220446049250313e-16;
      const double va = 10.0 * log10(acc_a);
      double colsum = va;
      logmel[((size_t)b * Tstride + (t0 + fa)) * kNfilt + m] = va;
      if (has_b) {
        const double vb = 10.0 * log10(acc_b);
        logmel[((size_t)b * Tstride + (t0 + fb)) * kNfilt + m] = vb;
        colsum += vb;
      }
      scol[warp * kNfilt + m] += colsum;
    }
    __syncwarp();
  }
  __syncthreads();
  if (threadIdx.x < kNfilt) {
    double s = 0.0;
    for (int w = 0; w < kFbankWarps; ++w) s += scol[w * kNfilt + threadIdx.x];
    part[threadIdx.x] = s;
  }
}

// ------------------------------------------------------------------------------------
// fbank_logmel2_kernel: the same stage with the 512-point FFT in REGISTERS.
//
// The radix-2 kernel above makes nine passes over a warp's 8 KB of complex fp64 data in shared memory (plus a
// bit-reversed scatter that lands 32 lanes on one bank): ~1150 shared-memory wavefronts per FFT before conflicts, ~3000
// with them, and the SM's one shared-memory pipe is what the kernel waits for (ncu round 1: sm throughput 17 %, no
// other pipe above 10 %).  Here 512 = 16 x 16 x 2:
//   1. lane b loads x[b + 32 j], j = 0..15 (consecutive lanes read consecutive samples) and runs a 16-point FFT over j
//      in its own registers: Y[b][k1];
//   2. twiddle: Z[b][k1] = Y[b][k1] W512^(b k1), the powers of W512^b by repeated multiplication;
//   3. ONE transpose through shared memory (rows padded to 33): lane (k1, b0) collects Z[2 b1 + b0][k1], b1 = 0..15,
//      and runs the second 16-point FFT over b1: G[b0][k1][c1];
//   4. the last radix-2 step X[k1 + 16 c1 + 256 c0] = G[0] +- W32^c1 G[1] is folded into the read of the unpack stage
//      (the lane that unpacks bin k also unpacks bin 256 - k: the four values it reads are exactly the ones the two
//      bins need, so the power spectra overwrite them in place).
// Shared-memory traffic per FFT: 64 + 64 wavefronts for the transpose, 64 for G, ~40 for the unpack.
// dynamic smem: double fft[8 warps][2][528]; double col[8][40]; double window[512]; float pcm[span + 1]
// ------------------------------------------------------------------------------------
static constexpr int kFftLd = 528;      // 16 rows of 33 (the transpose), >= 512 + 1 (G and the power spectra)

// cos / sin of c pi / 16, c = 0..16.  Not recursive: a recursive constexpr function is not inlined, and the calls -- their
// arguments are constants only after the loops around them are unrolled -- stay in the kernel as real calls (ncu, first
// version: a quarter of the stall samples in that subroutine).  A switch over literals folds away.
__host__ __device__ __forceinline__ constexpr double cos_pi16(int c) {
  switch (c) {
    case 0: return 1.0;
    case 1: return 0.98078528040323044913;
    case 2: return 0.92387953251128675613;
    case 3: return 0.83146961230254523708;
    case 4: return 0.70710678118654752440;
    case 5: return 0.55557023301960222474;
    case 6: return 0.38268343236508977173;
    case 7: return 0.19509032201612826785;
    case 8: return 0.0;
    case 9: return -0.19509032201612826785;
    case 10: return -0.38268343236508977173;
    case 11: return -0.55557023301960222474;
    case 12: return -0.70710678118654752440;
    case 13: return -0.83146961230254523708;
    case 14: return -0.92387953251128675613;
    case 15: return -0.98078528040323044913;
    default: return -1.0;
  }
}
__host__ __device__ __forceinline__ constexpr double sin_pi16(int c) { return c <= 8 ? cos_pi16(8 - c) : cos_pi16(c - 8); }
__host__ __device__ constexpr int bitrev4(int i) { return ((i & 1) << 3) | ((i & 2) << 1) | ((i & 4) >> 1) | ((i & 8) >> 3); }

// 16-point DFT (forward, e^{-2 pi i jk/16}) of a lane's registers, decimation in frequency; the result is left in
// bit-reversed order: X[k] = v[bitrev4(k)].  Everything is unrolled: indices and twiddles are compile-time constants.
__device__ __forceinline__ void fft16_dif(double (&re)[16], double (&im)[16]) {
#pragma unroll
  for (int span = 8; span >= 1; span >>= 1) {
#pragma unroll
    for (int g = 0; g < 16; g += 2 * span) {
#pragma unroll
      for (int k = 0; k < span; ++k) {
        const int i = g + k, j = i + span;
        const int m = k * (8 / span);                      // W16^m = cos(m pi/8) - i sin(m pi/8)
        const double tr = re[i] - re[j], ti = im[i] - im[j];
        re[i] += re[j]; im[i] += im[j];
        if (m == 0) { re[j] = tr; im[j] = ti; }
        else if (m == 4) { re[j] = ti; im[j] = -tr; }      // times -i
        else {
          const double c = cos_pi16(2 * m), sn = sin_pi16(2 * m);
          re[j] = tr * c + ti * sn;                        // (tr + i ti)(c - i sn)
          im[j] = ti * c - tr * sn;
        }
      }
    }
  }
}

template <int MINB>      // CTAs per SM the register budget is cut for (2: 128 registers and a few spilled doubles; 1: 255)
__global__ void __launch_bounds__(kFbankWarps * 32, MINB)
fbank_logmel2_kernel(const float* __restrict__ pcm, const int64_t* __restrict__ offsets,
                     const FbankTables* __restrict__ tab, int frame_length, int frame_step,
                     int Tstride, int nblk, double* __restrict__ logmel, double* __restrict__ partial,
                     int* __restrict__ nframes_out) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int b = blockIdx.y;
  const int64_t off = offsets[b];
  const int64_t n = offsets[b + 1] - off;
  int64_t d = n - frame_length; if (d < 0) d = -d;
  const int T = (int)((d + frame_step - 1) / frame_step);
  if (blockIdx.x == 0 && threadIdx.x == 0) nframes_out[b] = T;
  const int t0 = blockIdx.x * kFramesPerCta;
  double* part = partial + ((size_t)b * nblk + blockIdx.x) * kNfilt;
  if (t0 >= T) {
    if (threadIdx.x < kNfilt) part[threadIdx.x] = 0.0;
    return;
  }
  const int nfr = min(kFramesPerCta, T - t0);
  const int nuse = min(frame_length, kNfft);       // rfft(frames, 512) crops or zero-pads
  const int span = (kFramesPerCta - 1) * frame_step + nuse;
  double* sfft = reinterpret_cast<double*>(sm_raw);              // [8][2][528]
  double* scol = sfft + kFbankWarps * 2 * kFftLd;                // [8][40] per-warp column sums
  double* swin = scol + kFbankWarps * kNfilt;                    // [512]
  float* spcm = reinterpret_cast<float*>(swin + kNfft);          // [span + 1], spcm[i] = x[s0 - 1 + i]
  const int64_t s0 = (int64_t)t0 * frame_step;
  const float* x = pcm + off;
  // eight loads in flight per thread (one at a time, the loop is 21 exposed DRAM latencies: 18 % of the first version's
  // stall samples)
  for (int base = 0; base < span + 1; base += 8 * kFbankWarps * 32) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = base + j * kFbankWarps * 32 + threadIdx.x;
      const int64_t s = s0 - 1 + i;
      v[j] = (i < span + 1 && s >= 0 && s < n) ? __ldg(x + s) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = base + j * kFbankWarps * 32 + threadIdx.x;
      if (i < span + 1) spcm[i] = v[j];
    }
  }
  for (int i = threadIdx.x; i < kNfft; i += blockDim.x) swin[i] = tab->window[i];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = lane; i < kNfilt; i += 32) scol[warp * kNfilt + i] = 0.0;
  __syncthreads();

  double* re = sfft + warp * 2 * kFftLd;
  double* im = re + kFftLd;
  const double w1r = tab->tw_re[lane], w1i = tab->tw_im[lane];
  const int k1 = lane & 15, b0 = lane >> 4;

  for (int pair = warp; pair * 2 < nfr; pair += kFbankWarps) {
    const int fa = pair * 2, fb = fa + 1;
    const bool has_b = fb < nfr;
    double vr[16], vi[16];
    // load (float32 pre-emphasis, float64 window): frame A -> real parts, frame B -> imaginary parts
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int i = lane + 32 * j;
      double va = 0.0, vb = 0.0;
      if (i < nuse) {
        {
          const int p = fa * frame_step + i + 1;         // index into spcm (shifted by one)
          const int64_t s = s0 + fa * frame_step + i;    // absolute sample
          // y[0] = x[0]; y[s] = x[s] - 0.97 x[s-1] in float32 with separately rounded product
          // (numpy: sig[1:] - 0.97 * sig[:-1] on a float32 array); zero beyond the signal
          const float prev = (s > 0) ? spcm[p - 1] : 0.f;
          const float y = (s < n) ? __fsub_rn(spcm[p], __fmul_rn(0.97f, prev)) : 0.f;
          va = (double)y * swin[i];
        }
        if (has_b) {
          const int p = fb * frame_step + i + 1;
          const int64_t s = s0 + fb * frame_step + i;
          const float prev = (s > 0) ? spcm[p - 1] : 0.f;
          const float y = (s < n) ? __fsub_rn(spcm[p], __fmul_rn(0.97f, prev)) : 0.f;
          vb = (double)y * swin[i];
        }
      }
      vr[j] = va; vi[j] = vb;
    }
    fft16_dif(vr, vi);                                  // Y[lane][k] = v[bitrev4(k)]
    // twiddle by W512^(lane k) and transpose: row k of the buffer holds Z[.][k]
    {
      double pr = 1.0, pi = 0.0;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const double yr = vr[bitrev4(k)], yi = vi[bitrev4(k)];
        re[k * 33 + lane] = yr * pr - yi * pi;
        im[k * 33 + lane] = yr * pi + yi * pr;
        const double qr = pr * w1r - pi * w1i, qi = pr * w1i + pi * w1r;
        pr = qr; pi = qi;
      }
    }
    __syncwarp();
#pragma unroll
    for (int b1 = 0; b1 < 16; ++b1) { vr[b1] = re[k1 * 33 + 2 * b1 + b0]; vi[b1] = im[k1 * 33 + 2 * b1 + b0]; }
    __syncwarp();
    fft16_dif(vr, vi);                                  // G[b0][k1][c] = v[bitrev4(c)]
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      // lanes with b0 = 1 carry the odd half: times W32^c = cos(c pi/16) - i sin(c pi/16)
      const double wr = b0 ? cos_pi16(c) : 1.0, wi = b0 ? -sin_pi16(c) : 0.0;
      const double gr = vr[bitrev4(c)], gi = vi[bitrev4(c)];
      re[b0 * 256 + c * 16 + k1] = gr * wr - gi * wi;
      im[b0 * 256 + c * 16 + k1] = gr * wi + gi * wr;
    }
    __syncwarp();
    // X[k] = G0[k & 255] +- G1[k & 255] (minus for k >= 256).  Unpack the two real spectra, power = |X|^2 / 512: bins k
    // and 256 - k together (their inputs are G0 / G1 at k and at 256 - k), in place -- frame A's spectrum in re[0..256],
    // frame B's in im[0..256]
    for (int k = lane; k <= 128; k += 32) {
      const int k2 = 256 - k, k2i = k2 & 255;
      const double ar = re[k], ai = im[k], br = re[256 + k], bi = im[256 + k];
      const double cr = re[k2i], ci = im[k2i], dr = re[256 + k2i], di = im[256 + k2i];
      // bin k: Z = X[k], Y = X[512 - k];  bin 256 - k: Z' = X[256 - k], Y' = X[256 + k]
      const double zr = ar + br, zi = ai + bi, y2r = ar - br, y2i = ai - bi;
      double yr, yi, z2r, z2i;
      if (k == 0) { yr = zr; yi = zi; z2r = y2r; z2i = y2i; }         // X[0] and X[256] pair with themselves
      else { yr = cr - dr; yi = ci - di; z2r = cr + dr; z2i = ci + di; }
      {
        const double fr = 0.5 * (zr + yr), fi = 0.5 * (zi - yi), gr = 0.5 * (zi + yi), gi = -0.5 * (zr - yr);
        re[k] = (fr * fr + fi * fi) * (1.0 / kNfft);
        im[k] = (gr * gr + gi * gi) * (1.0 / kNfft);
      }
      {
        const double fr = 0.5 * (z2r + y2r), fi = 0.5 * (z2i - y2i), gr = 0.5 * (z2i + y2i), gi = -0.5 * (z2r - y2r);
        re[k2] = (fr * fr + fi * fi) * (1.0 / kNfft);
        im[k2] = (gr * gr + gi * gi) * (1.0 / kNfft);
      }
    }
    __syncwarp();
    const double* pw = re;
    const double* pwb = im;
    // 40 triangular filters; one lane owns filter m for BOTH frames of the pair so the
    // per-warp column sums are accumulated in a fixed order (deterministic).
    for (int m = lane; m < kNfilt; m += 32) {
      const double* w = tab->melw + m * kBins;
      double acc_a = 0.0, acc_b = 0.0;
      const int kend = tab->mend[m];
      for (int k = tab->mstart[m]; k < kend; ++k) {
        const double wk = w[k];
        acc_a = fma(pw[k], wk, acc_a);
        acc_b = fma(pwb[k], wk, acc_b);
      }
      if (acc_a == 0.0) acc_a = 2.220446049250313e-16;   // util/audioprocessor.py:135
      if (acc_b == 0.0) acc_b = 2.220446049250313e-16;
      const double va = 10.0 * log10(acc_a);
      double colsum = va;
      logmel[((size_t)b * Tstride + (t0 + fa)) * kNfilt + m] = va;
      if (has_b) {
        const double vb = 10.0 * log10(acc_b);
        logmel[((size_t)b * Tstride + (t0 + fb)) * kNfilt + m] = vb;
        colsum += vb;
      }
      scol[warp * kNfilt + m] += colsum;
    }
    __syncwarp();
  }
  __syncthreads();
  if (threadIdx.x < kNfilt) {
    double s = 0.0;
    for (int w = 0; w < kFbankWarps; ++w) s += scol[w * kNfilt + threadIdx.x];
    part[threadIdx.x] = s;
  }
}

// grid (ceil(Tmax / 32), B), 256 threads
__global__ void __launch_bounds__(256)
fbank_delta_kernel(const double* __restrict__ logmel, const double* __restrict__ partial,
                   const int* __restrict__ nframes, int Tstride, int nblk, int Tmax, int B,
                   int delta_mode, int time_major, float* __restrict__ out) {
  constexpr int TT = 32;
  __shared__ double xs[(TT + 32) * kNfilt];
  __shared__ double d1[(TT + 16) * kNfilt];
  __shared__ double mean[kNfilt];
  const int b = blockIdx.y;
  const int T = nframes[b];
  const int t0 = blockIdx.x * TT;
  const int Tout = min(T, Tmax);
  auto out_row = [&](int t) -> float* {
    return time_major ? out + ((size_t)t * B + b) * RS_FBANK_DIM : out + ((size_t)b * Tmax + t) * RS_FBANK_DIM;
  };
  if (t0 >= Tout) {   // pure zero fill
    for (int i = threadIdx.x; i < TT * RS_FBANK_DIM; i += blockDim.x) {
      int t = t0 + i / RS_FBANK_DIM;
      if (t < Tmax) out_row(t)[i % RS_FBANK_DIM] = 0.f;
    }
    return;
  }
  if (threadIdx.x < kNfilt) {
    // deterministic fixed-order sum of the per-CTA partials (double accumulate)
    double s = 0.0;
    for (int i = 0; i < nblk; ++i) s += partial[((size_t)b * nblk + i) * kNfilt + threadIdx.x];
    mean[threadIdx.x] = s / (double)T + 1e-8;
  }
  __syncthreads();
  // halos: delta-delta at frame t reads d1 in [t-8, t+8] ('interp' edge frames use a
  // centre up to 4 frames away), and d1 at frame i reads x in [i-8, i+8].
  const int xlo = t0 - 16;    // xs row r <-> frame xlo + r
  for (int i = threadIdx.x; i < (TT + 32) * kNfilt; i += blockDim.x) {
    const int t = xlo + i / kNfilt, m = i % kNfilt;
    xs[i] = (t >= 0 && t < T) ? logmel[((size_t)b * Tstride + t) * kNfilt + m] - mean[m] : 0.0;
  }
  __syncthreads();
  const int half = 4;
  const double inv = delta_mode == RS_DELTA_INTERP ? (1.0 / 60.0) : (1.0 / 20.0);
  // centre index used for output frame t
  auto centre = [&](int t) -> int {
    if (delta_mode == RS_DELTA_INTERP) return min(max(t, half), T - 1 - half);
    return t;
  };
  auto clampi = [&](int t) -> int { return min(max(t, 0), T - 1); };
  const int dlo = t0 - 8;     // d1 row r <-> frame dlo + r
  for (int i = threadIdx.x; i < (TT + 16) * kNfilt; i += blockDim.x) {
    const int t = dlo + i / kNfilt, m = i % kNfilt;
    double v = 0.0;
    if (t >= 0 && t < T) {
      const int c = centre(t);
#pragma unroll
      for (int j = -half; j <= half; ++j) v += (double)j * xs[(clampi(c + j) - xlo) * kNfilt + m];
      v *= inv;
    }
    d1[i] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TT * kNfilt; i += blockDim.x) {
    const int t = t0 + i / kNfilt, m = i % kNfilt;
    if (t >= Tmax) continue;
    float* o = out_row(t);
    if (t < Tout) {
      const int c = centre(t);
      double v = 0.0;
#pragma unroll
      for (int j = -half; j <= half; ++j) v += (double)j * d1[(clampi(c + j) - dlo) * kNfilt + m];
      o[m] = (float)xs[(t - xlo) * kNfilt + m];
      o[kNfilt + m] = (float)d1[(t - dlo) * kNfilt + m];
      o[2 * kNfilt + m] = (float)(v * inv);
    } else {
      o[m] = 0.f; o[kNfilt + m] = 0.f; o[2 * kNfilt + m] = 0.f;
    }
  }
}

}  // namespace rs

using namespace rs;

static size_t fbank_tables_bytes() { return align_up(sizeof(FbankTables), 256); }

extern "C" size_t rs_fbank_workspace_bytes(int B, int64_t max_samples, int sr) {
  FrameParams fp = frame_params(sr);
  int64_t Tfull = fbank_num_frames(max_samples, fp);
  int64_t nblk = (Tfull + kFramesPerCta - 1) / kFramesPerCta;
  if (nblk < 1) nblk = 1;
  return fbank_tables_bytes() + align_up((size_t)B * Tfull * kNfilt * sizeof(double), 256) +
         align_up((size_t)B * nblk * kNfilt * sizeof(double), 256);
}

// Host-side table dump for the CPU tests (no device work): dense mel weights
// [40*257] and the frame parameters the kernels will use.
extern "C" int rs_fbank_tables_host(int sr, float* melw_out, float* window_out, int* frame_length,
                                    int* frame_step) {
  RS_REQUIRE(sr > 0, RS_ERR_INVALID, "rs_fbank_tables_host: sr must be positive");
  const FbankTables* t = get_fbank_tables(sr);
  if (melw_out) for (int i = 0; i < kNfilt * kBins; ++i) melw_out[i] = (float)t->melw[i];
  if (window_out) for (int i = 0; i < kNfft; ++i) window_out[i] = (float)t->window[i];
  FrameParams fp = frame_params(sr);
  if (frame_length) *frame_length = fp.frame_length;
  if (frame_step) *frame_step = fp.frame_step;
  return RS_OK;
}

extern "C" int64_t rs_fbank_num_frames(int64_t n, int sr) {
  return fbank_num_frames(n, frame_params(sr));
}

extern "C" int rs_fbank_forward(const float* pcm_d, const int64_t* offsets_d, int B, int64_t max_samples,
                                int sr, int Tmax, int delta_mode, int time_major, float* out_d,
                                int32_t* nframes_d, void* ws_d, size_t ws_bytes, void* stream) {
  RS_REQUIRE(B > 0 && Tmax > 0 && sr > 0 && max_samples > 0, RS_ERR_INVALID,
             "rs_fbank_forward: bad arguments B=%d Tmax=%d sr=%d max_samples=%lld", B, Tmax, sr, (long long)max_samples);
  RS_REQUIRE(delta_mode == RS_DELTA_INTERP || delta_mode == RS_DELTA_EDGE, RS_ERR_INVALID,
             "rs_fbank_forward: unknown delta_mode %d", delta_mode);
  RS_REQUIRE(ws_bytes >= rs_fbank_workspace_bytes(B, max_samples, sr), RS_ERR_WORKSPACE,
             "rs_fbank_forward: workspace %zu < %zu", ws_bytes, rs_fbank_workspace_bytes(B, max_samples, sr));
  cudaStream_t st = (cudaStream_t)stream;
  FrameParams fp = frame_params(sr);
  RS_REQUIRE(fp.frame_step > 0 && fp.frame_length > 0, RS_ERR_INVALID, "rs_fbank_forward: sr %d too small", sr);
  const int64_t Tfull64 = fbank_num_frames(max_samples, fp);
  RS_REQUIRE(Tfull64 > 0 && Tfull64 < (1 << 30), RS_ERR_INVALID, "rs_fbank_forward: frame count %lld", (long long)Tfull64);
  const int Tfull = (int)Tfull64;
  const int nblk = cdiv(Tfull, kFramesPerCta);
  char* ws = (char*)ws_d;
  FbankTables* tab_d = (FbankTables*)ws;
  double* logmel = (double*)(ws + fbank_tables_bytes());
  double* partial = (double*)(ws + fbank_tables_bytes() + align_up((size_t)B * Tfull * kNfilt * sizeof(double), 256));
  const FbankTables* tab_h = get_fbank_tables(sr);
  RS_CHECK_CUDA(cudaMemcpyAsync(tab_d, tab_h, sizeof(FbankTables), cudaMemcpyHostToDevice, st));

  const int nuse = fp.frame_length < kNfft ? fp.frame_length : kNfft;
  const int span = (kFramesPerCta - 1) * fp.frame_step + nuse;
  // register FFT (fbank_logmel2_kernel); RS_FBANK_FFT=0: the shared-memory radix-2 kernel
  static const bool reg_fft = [] { const char* v = getenv("RS_FBANK_FFT"); return !(v && v[0] == '0'); }();
  const int fft_ld = reg_fft ? kFftLd : kNfft;
  size_t smem = (size_t)(kFbankWarps * 2 * fft_ld + kFbankWarps * kNfilt + (reg_fft ? kNfft : 0)) * sizeof(double) +
                (size_t)(span + 1 + 3) * sizeof(float);
  RS_REQUIRE(smem <= 220 * 1024, RS_ERR_UNSUPPORTED, "rs_fbank_forward: sr %d needs %zu B smem", sr, smem);
  static const int occ = [] { const char* v = getenv("RS_FBANK_OCC"); return v ? atoi(v) : 2; }();
  auto kern = reg_fft ? (occ == 1 ? fbank_logmel2_kernel<1> : fbank_logmel2_kernel<2>) : fbank_logmel_kernel;
  RS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  kern<<<dim3(nblk, B), kFbankWarps * 32, smem, st>>>(pcm_d, offsets_d, tab_d, fp.frame_length, fp.frame_step, Tfull, nblk,
                                                      logmel, partial, nframes_d);
  RS_CHECK_LAUNCH();
  fbank_delta_kernel<<<dim3(cdiv(Tmax, 32), B), 256, 0, st>>>(logmel, partial, nframes_d, Tfull, nblk, Tmax, B,
                                                              delta_mode, time_major, out_d);
  RS_CHECK_LAUNCH();
  return RS_OK;
}
