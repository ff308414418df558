// Internal definitions shared by the acoustic-model translation units.
#pragma once
#include "common.cuh"
#include "lstm_rec_tc.cuh"
#include <vector>

struct rs_am {
  int L, H, F, C, B, Tmax;
  int normalization;             // batch-norm of the stack's input (models/AcousticModel.py:253-259), default off
  int64_t n_params;
  int64_t off_input_w, off_input_b, off_output_w, off_output_b;
  int64_t off_kernel[64], off_bias[64];
  // optional per-launch timing of the recurrent kernels (CUDA events on the launching stream):
  // tev[fwd|bwd][layer] holds (start, stop) pairs, one pair per (chunk) launch of the last call
  int timing;
  std::vector<cudaEvent_t> tev[4][64];   // [fwd rec | bwd rec | bwd weight-gradient GEMMs | bwd dx GEMMs][layer]
  int tev_used[4][64];          // events recorded by the last call (2 per launch)
  cudaEvent_t tev_base[2];      // recorded on the caller's stream at the top of the call (time origin of a trace)
  int tev_base_ready;
  // tensor-core path (H % 64 == 0, B <= 64, weights fit in shared memory); else FFMA kernels
  int use_tc;
  // Streams of the pipelined schedule (lstm_tc.cu): one per layer for the chunked recurrent launches, one for the
  // chunk GEMMs between layers, one for the weight-gradient work; events come from a pool that is reused every call.
  cudaStream_t lane[64], gemm_st, side, tr_st;
  int streams_ready;
  std::vector<cudaEvent_t> evpool;
  size_t ev_next;
  int chunk;                     // time steps per chunked launch (RS_TC_CHUNK; 0 = one launch per layer)
  int chunk_is_default;          // RS_TC_CHUNK not set: the forward phase schedule picks its own chunk length
  int window;                    // recurrent launches allowed in flight (RS_TC_WINDOW)
  unsigned long long* dbg_fwd;   // optional device buffers [T][8] for kernel timelines (layer 0)
  unsigned long long* dbg_bwd;
  rs::RecTcGeom tc;
  // Weight planes (bf16 hi/lo, packed for the kernels) live in the caller's workspace.  With a non-zero
  // params_version (rs_am_set_params_version) the caller vouches that the parameters only change when the version
  // does and that nobody else writes the workspace between calls: the planes are then re-packed once per version
  // instead of once per forward / backward call.
  unsigned long long params_version;
  unsigned long long packed_version[2];      // [forward set | backward set]
  const void* packed_ws[2];
  const void* packed_params[2];
};


namespace rs {

// Timed (start, stop) event pairs around the recurrent launches; tev_begin() at the top of a forward / backward call.
inline void tev_begin(rs_am* am, int dir, cudaStream_t st = nullptr) {
  for (int l = 0; l < 64; ++l) am->tev_used[dir][l] = 0;
  if (dir == 1) for (int l = 0; l < 64; ++l) am->tev_used[2][l] = am->tev_used[3][l] = 0;
  if (am->timing) {
    if (!am->tev_base_ready) { cudaEventCreate(&am->tev_base[0]); cudaEventCreate(&am->tev_base[1]); am->tev_base_ready = 1; }
    cudaEventRecord(am->tev_base[dir], st);
  }
}
inline int tev_record(rs_am* am, int dir, int l, cudaStream_t st) {
  if (!am->timing) return RS_OK;
  int& u = am->tev_used[dir][l];
  if ((size_t)u >= am->tev[dir][l].size()) {
    cudaEvent_t e;
    RS_CHECK_CUDA(cudaEventCreate(&e));
    am->tev[dir][l].push_back(e);
  }
  RS_CHECK_CUDA(cudaEventRecord(am->tev[dir][l][u], st));
  ++u;
  return RS_OK;
}

// batchnorm.cu: x [T,B,H] -> x_hat in place (+ 1/std [T,H]); d -> gradient wrt x in place
int bn_forward(float* x, float* istd, int T, int B, int H, cudaStream_t st);
int bn_backward(float* d, const float* xhat, const float* istd, int T, int B, int H, cudaStream_t st);

// Tensor-core path (lstm_tc.cu).  Same contract as rs_am_forward / rs_am_backward.
size_t am_tc_reserve_bytes(const rs_am* am);
size_t am_tc_workspace_bytes(const rs_am* am);
int am_tc_forward(rs_am* am, const float* params_d, const float* x_d, const int32_t* len_d, int T,
                  const float* state_in_d, float* state_out_d, float keep_in, float keep_out, uint64_t seed,
                  float* logits_d, void* reserve_d, void* ws_d, size_t ws_bytes, cudaStream_t st);
int am_tc_backward(rs_am* am, const float* params_d, const float* x_d, const int32_t* len_d, int T, float keep_in,
                   float keep_out, uint64_t seed, const float* dlogits_d, void* reserve_d, float* grads_d,
                   void* ws_d, size_t ws_bytes, cudaStream_t st);

}  // namespace rs
