// MFCC (librosa.feature.mfcc semantics, util/audioprocessor.py:63-75).
// Placeholder translation unit: implemented in a later milestone of this round.
#include "common.cuh"

extern "C" size_t rs_mfcc_workspace_bytes(int B, int64_t max_samples, int sr) {
  (void)B; (void)max_samples; (void)sr;
  return 0;
}

extern "C" int64_t rs_mfcc_num_frames(int64_t n, int sr) {
  const int hop = (int)nearbyint(0.01 * sr);
  return hop > 0 ? 1 + n / hop : 0;
}

extern "C" int rs_mfcc_forward(const float* pcm_d, const int64_t* offsets_d, int B, int64_t max_samples, int sr,
                               int Tmax, int n_mfcc, int time_major, float* out_d, int32_t* nframes_d, void* ws_d,
                               size_t ws_bytes, void* stream) {
  (void)pcm_d; (void)offsets_d; (void)B; (void)max_samples; (void)sr; (void)Tmax; (void)n_mfcc; (void)time_major;
  (void)out_d; (void)nframes_d; (void)ws_d; (void)ws_bytes; (void)stream;
  rs::set_error("rs_mfcc_forward: not implemented yet");
  return RS_ERR_UNSUPPORTED;
}
