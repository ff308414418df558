// Levenshtein distance between the decoded and the reference label sequences (sm_100a).
//
// Replaces tf.edit_distance(prediction, sparse_labels, normalize=True) -- the reference's error_rate
// (/root/reference/models/AcousticModel.py:370), which drives its learning-rate schedule (stt.py:219-226).
// with the distance on the device the training step needs no host round trip for the metric.
//
// One warp per utterance.  Row i of the DP table over the truth positions j:
//     E[j]   = min(prev[j] + 1, prev[j-1] + (truth[j-1] != hyp[i-1]))          (deletion / substitution)
//     cur[j] = min_{k <= j} (E[k] + (j - k))                                     (runs of insertions)
// The second line is a prefix minimum of E[k] - k: each lane scans a contiguous segment, a shuffle scan carries
// the minima across lanes.  rate = distance / len(truth)  (inf for an empty truth and a non-empty hypothesis, 0
// for two empty sequences, as TF does).
#include "common.cuh"

namespace rs {
namespace {

constexpr int kWarps = 4;

__global__ void __launch_bounds__(kWarps * 32)
edit_distance_kernel(const int* __restrict__ hyp, const int* __restrict__ hyp_len, int hyp_ld,
                     const int* __restrict__ truth, const int* __restrict__ truth_off, int B, int row_len,
                     int* __restrict__ dist, float* __restrict__ rate) {
  extern __shared__ int rows[];                       // [kWarps][2][row_len]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kWarps + warp;
  if (b >= B) return;
  const int* t = truth + truth_off[b];
  const int N = truth_off[b + 1] - truth_off[b];
  const int M = max(hyp_len[b], 0);
  const int* h = hyp + (size_t)b * hyp_ld;
  int* prev = rows + (size_t)warp * 2 * row_len;
  int* cur = prev + row_len;
  for (int j = lane; j <= N; j += 32) prev[j] = j;
  __syncwarp();
  const int S = (N + 31) / 32;                        // truth positions per lane
  const int j0 = 1 + lane * S, j1 = min(N, j0 + S - 1);
  for (int i = 1; i <= M; ++i) {
    const int hi = h[i - 1];
    // local pass: v_j = E[j] - j and its running minimum
    int run = 0x3fffffff;
    for (int j = j0; j <= j1; ++j) {
      const int e = min(prev[j] + 1, prev[j - 1] + (t[j - 1] != hi ? 1 : 0));
      run = min(run, e - j);
      cur[j] = run;                                   // local prefix minimum for now
    }
    // exclusive prefix minimum of the segment minima over the lanes, seeded with E[0] - 0 = i
    int incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl = min(incl, up);
    }
    int carry = __shfl_up_sync(0xffffffffu, incl, 1);
    carry = (lane == 0) ? i : min(carry, i);
    for (int j = j0; j <= j1; ++j) cur[j] = min(cur[j], carry) + j;
    if (lane == 0) cur[0] = i;
    __syncwarp();
    int* tmp = prev; prev = cur; cur = tmp;
  }
  if (lane == 0) {
    const int d = prev[N];
    dist[b] = d;
    if (rate) rate[b] = N > 0 ? (float)d / (float)N : (d > 0 ? INFINITY : 0.f);
  }
}

}  // namespace
}  // namespace rs

using namespace rs;

extern "C" int rs_edit_distance(const int32_t* hyp_d, const int32_t* hyp_len_d, int hyp_ld, const int32_t* truth_d,
                                const int32_t* truth_offsets_d, int B, int max_truth_len, int32_t* dist_d,
                                float* rate_d, void* stream) {
  RS_REQUIRE(hyp_d && hyp_len_d && truth_d && truth_offsets_d && dist_d, RS_ERR_INVALID, "rs_edit_distance: NULL argument");
  RS_REQUIRE(B > 0 && hyp_ld > 0 && max_truth_len >= 0, RS_ERR_INVALID, "rs_edit_distance: bad shape");
  const int row_len = max_truth_len + 1;
  const size_t smem = (size_t)kWarps * 2 * row_len * sizeof(int);
  RS_REQUIRE(smem <= 48 * 1024, RS_ERR_UNSUPPORTED, "rs_edit_distance: truth length %d too large", max_truth_len);
  edit_distance_kernel<<<cdiv(B, kWarps), kWarps * 32, smem, (cudaStream_t)stream>>>(hyp_d, hyp_len_d, hyp_ld, truth_d,
                                                                                     truth_offsets_d, B, row_len, dist_d, rate_d);
  RS_CHECK_LAUNCH();
  return RS_OK;
}
