// Audio front door on the device: decoded PCM -> float32 mono -> 22 050 Hz.
//
// Replaces what the reference reaches through librosa.load(file_name, mono=True)
// (util/audioprocessor.py:49): buf_to_float (int16 * 2^-15), to_mono (mean over
// channels) and resampy.resample(..., filter='kaiser_best') followed by
// fix_length(ceil(n * ratio)).  librosa / resampy are third-party and absent;
// the algorithm restated here is resampy's band-limited sinc interpolation
// (Smith's method): a one-sided Kaiser-windowed sinc table with 64 zero
// crossings and 512 samples per crossing, linearly interpolated between table
// entries, walked with stride min(1, ratio) * 512 on both wings of each output
// sample.  oracle/resample.py is the CPU restatement the tests check against.
//
// HBM-bound streaming work: 2 (int16) or 4 (float) bytes read per input sample
// and 4 written per output sample; the input window of a block is staged once
// in shared memory (coalesced), the 512 KB filter table stays in L2.
#include "common.cuh"

namespace rs {

constexpr int kNumZeros = 64;
constexpr int kNumTable = 512;                         // 2^precision, precision = 9
constexpr int kNwin = kNumZeros * kNumTable + 1;       // 32 769 one-sided taps
constexpr double kBeta = 14.769656459379492;           // resampy 'kaiser_best'
constexpr double kRolloff = 0.9475937167399596;
constexpr int kResampleThreads = 256;
constexpr int kOutPerThread = 4;
constexpr int kOutPerBlock = kResampleThreads * kOutPerThread;

// modified Bessel function of the first kind, order 0: sum_k ((x/2)^k / k!)^2 (all terms positive)
__host__ __device__ inline double bessel_i0(double x) {
  const double q = 0.25 * x * x;
  double term = 1.0, sum = 1.0;
  for (int k = 1; k < 500; ++k) {
    term *= q / ((double)k * (double)k);
    sum += term;
    if (term < 1e-17 * sum) break;
  }
  return sum;
}

// tap i of the one-sided filter: kaiser(2n+1, beta)[n + i] * rolloff * sinc(rolloff * i / 512)
__host__ __device__ inline double filter_tap(int i, double inv_i0_beta) {
  const double n = (double)(kNwin - 1);
  const double r = (double)i / n;
  const double arg = 1.0 - r * r;
  const double taper = bessel_i0(kBeta * sqrt(arg > 0.0 ? arg : 0.0)) * inv_i0_beta;
  const double xs = kRolloff * ((double)i / (double)kNumTable);
  const double y = 3.141592653589793238462643383279502884 * (xs == 0.0 ? 1e-20 : xs);
  return taper * (kRolloff * (sin(y) / y));
}

// table[i] = {gain * win[i], gain * (win[i+1] - win[i])}, delta of the last tap = 0
__global__ void resample_table_kernel(double2* __restrict__ table, double gain) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kNwin) return;
  const double inv = 1.0 / bessel_i0(kBeta);
  const double w0 = gain * filter_tap(i, inv);
  const double w1 = (i + 1 < kNwin) ? gain * filter_tap(i + 1, inv) : w0;
  table[i] = make_double2(w0, w1 - w0);
}

struct ResamplePlan {
  double sample_ratio, scale, time_increment;
  int index_step;      // int(scale * 512)
  int wing;            // most taps one wing can have: 64 / scale, rounded up, + 1
  int window;          // input samples staged per block
};

static ResamplePlan make_plan(int sr_in, int sr_out) {
  ResamplePlan p;
  p.sample_ratio = (double)sr_out / (double)sr_in;
  p.scale = p.sample_ratio < 1.0 ? p.sample_ratio : 1.0;
  p.time_increment = 1.0 / p.sample_ratio;
  p.index_step = (int)(p.scale * kNumTable);
  p.wing = p.index_step > 0 ? (kNwin + p.index_step - 1) / p.index_step + 1 : 0;
  p.window = (int)ceil((double)kOutPerBlock * p.time_increment) + 2 * p.wing + 4;
  return p;
}

template <typename T> struct PcmLoad;
template <> struct PcmLoad<float> {
  // to_mono on already-scaled samples: float32 sum over the channels, divided by their number
  static __device__ __forceinline__ float at(const float* base, int64_t frame, int nch) {
    if (nch == 1) return __ldg(base + frame);
    const float* p = base + frame * nch;
    float s = 0.0f;
    for (int c = 0; c < nch; ++c) s += __ldg(p + c);
    return s / (float)nch;
  }
};
template <> struct PcmLoad<int16_t> {
  // buf_to_float then to_mono: float32 sum of x_c * 2^-15 over the channels, divided by their number
  static __device__ __forceinline__ float at(const int16_t* base, int64_t frame, int nch) {
    const int16_t* p = base + frame * nch;
    float s = 0.0f;
    for (int c = 0; c < nch; ++c) s += (float)__ldg(p + c) * (1.0f / 32768.0f);
    return nch > 1 ? s / (float)nch : s;
  }
};

// resampy advances its read position with `time_register += time_increment` once per output sample, and the rounding
// of that running float64 sum decides on which side of an input sample an output lands whenever t / ratio is an
// integer.  With ratio < 1 the two sides are NOT equivalent (the table stride int(ratio * 512) is truncated), so the
// running sum is reproduced add for add: one thread per utterance walks it and leaves a checkpoint for every block of
// the main kernel, which continues from there.  (A dependent chain of n_out double adds: ~1 ms for 10 s of audio.)
__global__ void resample_time_kernel(const int64_t* __restrict__ offsets, int B, ResamplePlan plan, int nblk,
                                     double* __restrict__ ckpt) {
  const int b = blockIdx.x;
  if (b >= B || threadIdx.x != 0) return;
  const int64_t n_orig = offsets[b + 1] - offsets[b];
  const int64_t n_out = (int64_t)((double)n_orig * plan.sample_ratio);
  double tr = 0.0;
  double* c = ckpt + (size_t)b * nblk;
  for (int64_t t0 = 0; t0 < n_out; t0 += kOutPerBlock) {
    *c++ = tr;
    const int64_t stop = (n_out - t0 < kOutPerBlock) ? n_out - t0 : kOutPerBlock;
    for (int64_t j = 0; j < stop; ++j) tr += plan.time_increment;
  }
}

// grid (blocks over output samples, B); block b handles outputs [blk*kOutPerBlock, +kOutPerBlock) of one utterance
template <typename T>
__global__ void __launch_bounds__(kResampleThreads)
resample_kernel(const T* __restrict__ pcm, const int64_t* __restrict__ offsets, int nch,
                const double2* __restrict__ table, ResamplePlan plan, int nblk, const double* __restrict__ ckpt,
                float* __restrict__ out, const int64_t* __restrict__ out_offsets) {
  extern __shared__ float xs[];
  __shared__ double times[kOutPerBlock];
  const int b = blockIdx.y;
  const int64_t in0 = offsets[b], n_orig = offsets[b + 1] - in0;
  const int64_t o0 = out_offsets[b], n_fix = out_offsets[b + 1] - o0;
  const int64_t n_out = (int64_t)((double)n_orig * plan.sample_ratio);       // int(n * ratio): what resampy writes
  const int64_t t0 = (int64_t)blockIdx.x * kOutPerBlock;
  if (t0 >= n_fix) return;
  // ckpt == nullptr (ratio >= 1): t * time_increment stands in for the running sum.  The two can only disagree on
  // which neighbour of an exact integer position they pick, and with the table stride at exactly 512 both choices
  // address the same taps (x[n] with the tap at 0 vs x[n-1]'s successor at the end of its interval): the filter is
  // continuous there, so nothing is lost and the serial replay is skipped.
  const double start = ckpt == nullptr ? (double)t0 * plan.time_increment
                                       : (t0 < n_out ? ckpt[(size_t)b * nblk + blockIdx.x] : 0.0);
  if (ckpt == nullptr) {
    for (int j = threadIdx.x; j < kOutPerBlock; j += kResampleThreads) times[j] = (double)(t0 + j) * plan.time_increment;
  } else if (threadIdx.x == 0) {             // the time register of every output of this block, add for add
    double tr = start;
    for (int j = 0; j < kOutPerBlock; ++j) {
      times[j] = tr;
      tr += plan.time_increment;
    }
  }
  // input window of this block: [w0, w0 + window)
  const int64_t w0 = (int64_t)start - plan.wing - 1;
  const T* base = pcm + in0 * (int64_t)nch;
  for (int i = threadIdx.x; i < plan.window; i += kResampleThreads) {
    const int64_t f = w0 + i;
    xs[i] = (f >= 0 && f < n_orig) ? PcmLoad<T>::at(base, f, nch) : 0.0f;
  }
  __syncthreads();
  const int64_t nwin = kNwin;
  for (int j = 0; j < kOutPerThread; ++j) {
    const int64_t t = t0 + j * kResampleThreads + threadIdx.x;
    if (t >= n_fix) break;
    if (t >= n_out) {                       // librosa.util.fix_length: the sample resampy does not produce is zero
      out[o0 + t] = 0.0f;
      continue;
    }
    const double time_register = times[j * kResampleThreads + threadIdx.x];
    const int64_t n = (int64_t)time_register;
    const float* xc = xs + (n - w0);        // xc[0] = x[n]
    double acc = 0.0;
    // left wing: x[n], x[n-1], ...
    double frac = plan.scale * (time_register - (double)n);
    double index_frac = frac * kNumTable;
    int offset = (int)index_frac;
    double eta = index_frac - offset;
    int64_t cnt = (nwin - offset) / plan.index_step;
    if (n + 1 < cnt) cnt = n + 1;
    for (int i = 0; i < (int)cnt; ++i) {
      const double2 w = __ldg(table + offset + i * plan.index_step);
      acc += (w.x + eta * w.y) * (double)xc[-i];
    }
    // right wing: x[n+1], x[n+2], ...
    frac = plan.scale - frac;
    index_frac = frac * kNumTable;
    offset = (int)index_frac;
    eta = index_frac - offset;
    cnt = (nwin - offset) / plan.index_step;
    if (n_orig - n - 1 < cnt) cnt = n_orig - n - 1;
    for (int k = 0; k < (int)cnt; ++k) {
      const double2 w = __ldg(table + offset + k * plan.index_step);
      acc += (w.x + eta * w.y) * (double)xc[k + 1];
    }
    out[o0 + t] = (float)acc;
  }
}

// Upsampling by a rational ratio sr_out / sr_in = (P / g') with P = sr_out / gcd <= 1024: the interpolation phase
// (table offset and blend factor) of output t repeats every P outputs, so each thread computes kPhaseOutputs outputs
// that are S = P * ceil(256 / P) apart and loads every table entry once for all of them (the table loads from L2 bound
// the per-output kernel: ncu lts throughput 72 %).  Read position and phase come from exact integer arithmetic,
// t * sr_in = n * sr_out + r: frac = r / sr_out.  Block blk handles outputs [blk * 4S, (blk + 1) * 4S).
constexpr int kPhaseOutputs = 4;
template <typename T>
__global__ void __launch_bounds__(kResampleThreads)
resample_periodic_kernel(const T* __restrict__ pcm, const int64_t* __restrict__ offsets, int nch,
                         const double2* __restrict__ table, ResamplePlan plan, int sr_in, int sr_out, int S, int step_in,
                         int window, float* __restrict__ out, const int64_t* __restrict__ out_offsets) {
  extern __shared__ float xs[];
  const int b = blockIdx.y;
  const int64_t in0 = offsets[b], n_orig = offsets[b + 1] - in0;
  const int64_t o0 = out_offsets[b], n_fix = out_offsets[b + 1] - o0;
  const int64_t n_out = (int64_t)((double)n_orig * plan.sample_ratio);
  const int64_t t0 = (int64_t)blockIdx.x * (kPhaseOutputs * S);
  if (t0 >= n_fix) return;
  const int64_t w0 = (t0 * sr_in) / sr_out - plan.wing - 1;
  const T* base = pcm + in0 * (int64_t)nch;
  for (int i = threadIdx.x; i < window; i += kResampleThreads) {
    const int64_t f = w0 + i;
    xs[i] = (f >= 0 && f < n_orig) ? PcmLoad<T>::at(base, f, nch) : 0.0f;      // zeros outside the utterance
  }
  __syncthreads();
  for (int u = threadIdx.x; u < S; u += kResampleThreads) {
    const int64_t t = t0 + u;
    if (t >= n_fix) break;
    const int64_t num = t * sr_in;
    const int64_t n = num / sr_out;
    const double frac = (double)(num - n * sr_out) / (double)sr_out;
    const float* xc = xs + (n - w0);        // xc[j * step_in] = x[n_j] of output t + j * S
    double acc[kPhaseOutputs];
#pragma unroll
    for (int j = 0; j < kPhaseOutputs; ++j) acc[j] = 0.0;
    // left wing; taps that would fall before the utterance meet the zeros staged above
    double index_frac = frac * kNumTable;
    int offset = (int)index_frac;
    double eta = index_frac - offset;
    int cnt = (kNwin - offset) / kNumTable;
    for (int i = 0; i < cnt; ++i) {
      const double2 w = __ldg(table + offset + i * kNumTable);
      const double wgt = w.x + eta * w.y;
#pragma unroll
      for (int j = 0; j < kPhaseOutputs; ++j) acc[j] += wgt * (double)xc[j * step_in - i];
    }
    // right wing
    index_frac = (1.0 - frac) * kNumTable;
    offset = (int)index_frac;
    eta = index_frac - offset;
    cnt = (kNwin - offset) / kNumTable;
    for (int k = 0; k < cnt; ++k) {
      const double2 w = __ldg(table + offset + k * kNumTable);
      const double wgt = w.x + eta * w.y;
#pragma unroll
      for (int j = 0; j < kPhaseOutputs; ++j) acc[j] += wgt * (double)xc[j * step_in + k + 1];
    }
#pragma unroll
    for (int j = 0; j < kPhaseOutputs; ++j) {
      const int64_t tj = t + (int64_t)j * S;
      if (tj < n_out) out[o0 + tj] = (float)acc[j];
      else if (tj < n_fix) out[o0 + tj] = 0.0f;        // librosa.util.fix_length
    }
  }
}

// interleaved int16 / float32 -> float32 mono, no rate change (file already at the target rate)
template <typename T>
__global__ void pcm_to_mono_kernel(const T* __restrict__ pcm, int64_t frames, int nch, float* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < frames; f += stride)
    out[f] = PcmLoad<T>::at(pcm, f, nch);
}

}  // namespace rs

using namespace rs;

static size_t resample_table_bytes() { return align_up((size_t)kNwin * sizeof(double2), 256); }
static int resample_blocks(int64_t max_out_samples) { return (int)((max_out_samples + kOutPerBlock - 1) / kOutPerBlock); }

extern "C" size_t rs_resample_workspace_bytes(int B, int64_t max_out_samples) {
  if (B < 1 || max_out_samples < 1) return resample_table_bytes();
  return resample_table_bytes() + align_up((size_t)B * resample_blocks(max_out_samples) * sizeof(double), 256);
}

extern "C" int64_t rs_resample_num_samples(int64_t n_samples, int sr_in, int sr_out) {
  if (n_samples < 0 || sr_in <= 0 || sr_out <= 0) return -1;
  if (sr_in == sr_out) return n_samples;
  return (int64_t)ceil((double)n_samples * ((double)sr_out / (double)sr_in));
}

extern "C" int rs_resample_filter_host(double* win_out, int* num_table) {
  RS_REQUIRE(win_out != nullptr, RS_ERR_INVALID, "rs_resample_filter_host: null output");
  const double inv = 1.0 / bessel_i0(kBeta);
  for (int i = 0; i < kNwin; ++i) win_out[i] = filter_tap(i, inv);
  if (num_table) *num_table = kNumTable;
  return RS_OK;
}

extern "C" int rs_resample_forward(const void* pcm_d, int pcm_format, int channels, const int64_t* offsets_d, int B,
                                   int64_t max_out_samples, int sr_in, int sr_out, float* out_d,
                                   const int64_t* out_offsets_d, void* ws_d, size_t ws_bytes, void* stream) {
  RS_REQUIRE(B > 0 && max_out_samples > 0 && sr_in > 0 && sr_out > 0, RS_ERR_INVALID,
             "rs_resample_forward: bad arguments B=%d max_out_samples=%lld sr %d -> %d", B, (long long)max_out_samples,
             sr_in, sr_out);
  RS_REQUIRE(pcm_format == RS_PCM_F32 || pcm_format == RS_PCM_S16, RS_ERR_INVALID,
             "rs_resample_forward: unknown pcm_format %d", pcm_format);
  RS_REQUIRE(channels >= 1 && channels <= 8, RS_ERR_INVALID, "rs_resample_forward: %d channels (1..8)", channels);
  RS_REQUIRE(B <= 65535, RS_ERR_UNSUPPORTED, "rs_resample_forward: at most 65535 utterances per call, got %d", B);
  RS_REQUIRE(sr_in != sr_out, RS_ERR_INVALID, "rs_resample_forward: sr_in == sr_out (use rs_pcm16_to_f32 / no call)");
  RS_REQUIRE(max_out_samples < ((int64_t)1 << 40), RS_ERR_INVALID, "rs_resample_forward: max_out_samples %lld",
             (long long)max_out_samples);
  RS_REQUIRE(ws_bytes >= rs_resample_workspace_bytes(B, max_out_samples), RS_ERR_WORKSPACE,
             "rs_resample_forward: workspace %zu < %zu", ws_bytes, rs_resample_workspace_bytes(B, max_out_samples));
  const ResamplePlan plan = make_plan(sr_in, sr_out);
  const size_t smem = (size_t)plan.window * sizeof(float);
  RS_REQUIRE(plan.index_step >= 1 && smem <= 200 * 1024, RS_ERR_UNSUPPORTED,
             "rs_resample_forward: ratio %d -> %d needs a %zu B window", sr_in, sr_out, smem);
  cudaStream_t st = (cudaStream_t)stream;
  double2* table = (double2*)ws_d;
  resample_table_kernel<<<cdiv(kNwin, 256), 256, 0, st>>>(table, plan.sample_ratio < 1.0 ? plan.sample_ratio : 1.0);
  RS_CHECK_LAUNCH();
  // rational upsampling with a short phase period: the shared-weights kernel
  int g = sr_in, r = sr_out;
  while (r) { const int tmp = g % r; g = r; r = tmp; }
  const int period = sr_out / g;
  if (plan.sample_ratio > 1.0 && period <= 1024 && max_out_samples < ((int64_t)1 << 40) / sr_in) {
    const int S = period * ((kResampleThreads + period - 1) / period);
    const int step_in = (S / period) * (sr_in / g);
    const int window = (int)ceil((double)(kPhaseOutputs * S) * plan.time_increment) + 2 * plan.wing + 4;
    const size_t psmem = (size_t)window * sizeof(float);
    const int per_block = kPhaseOutputs * S;
    const dim3 pgrid((unsigned)((max_out_samples + per_block - 1) / per_block), (unsigned)B);
    if (pcm_format == RS_PCM_S16)
      resample_periodic_kernel<int16_t><<<pgrid, kResampleThreads, psmem, st>>>(
          (const int16_t*)pcm_d, offsets_d, channels, table, plan, sr_in, sr_out, S, step_in, window, out_d, out_offsets_d);
    else
      resample_periodic_kernel<float><<<pgrid, kResampleThreads, psmem, st>>>(
          (const float*)pcm_d, offsets_d, channels, table, plan, sr_in, sr_out, S, step_in, window, out_d, out_offsets_d);
    RS_CHECK_LAUNCH();
    return RS_OK;
  }
  const int nblk = resample_blocks(max_out_samples);
  double* ckpt = nullptr;
  if (plan.sample_ratio < 1.0) {            // see resample_time_kernel: only a truncated table stride makes it matter
    ckpt = (double*)((char*)ws_d + resample_table_bytes());
    resample_time_kernel<<<B, 32, 0, st>>>(offsets_d, B, plan, nblk, ckpt);
    RS_CHECK_LAUNCH();
  }
  const dim3 grid((unsigned)nblk, (unsigned)B);
  if (pcm_format == RS_PCM_S16) {
    RS_CHECK_CUDA(cudaFuncSetAttribute(resample_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    resample_kernel<int16_t><<<grid, kResampleThreads, smem, st>>>((const int16_t*)pcm_d, offsets_d, channels, table,
                                                                  plan, nblk, ckpt, out_d, out_offsets_d);
  } else {
    RS_CHECK_CUDA(cudaFuncSetAttribute(resample_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    resample_kernel<float><<<grid, kResampleThreads, smem, st>>>((const float*)pcm_d, offsets_d, channels, table, plan, nblk, ckpt,
                                                                out_d, out_offsets_d);
  }
  RS_CHECK_LAUNCH();
  return RS_OK;
}

extern "C" int rs_pcm16_to_f32(const int16_t* pcm_d, int64_t frames, int channels, float* out_d, void* stream) {
  RS_REQUIRE(frames >= 0 && channels >= 1 && channels <= 8, RS_ERR_INVALID, "rs_pcm16_to_f32: frames=%lld channels=%d",
             (long long)frames, channels);
  if (frames == 0) return RS_OK;
  int64_t blocks = (frames + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  pcm_to_mono_kernel<int16_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pcm_d, frames, channels, out_d);
  RS_CHECK_LAUNCH();
  return RS_OK;
}

extern "C" int rs_pcm_f32_to_mono(const float* pcm_d, int64_t frames, int channels, float* out_d, void* stream) {
  RS_REQUIRE(frames >= 0 && channels >= 1 && channels <= 8, RS_ERR_INVALID, "rs_pcm_f32_to_mono: frames=%lld channels=%d",
             (long long)frames, channels);
  if (frames == 0) return RS_OK;
  int64_t blocks = (frames + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  pcm_to_mono_kernel<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pcm_d, frames, channels, out_d);
  RS_CHECK_LAUNCH();
  return RS_OK;
}
