// Internal definitions shared by the acoustic-model translation units.
#pragma once
#include "common.cuh"
#include "lstm_rec_tc.cuh"

struct rs_am {
  int L, H, F, C, B, Tmax;
  int64_t n_params;
  int64_t off_input_w, off_input_b, off_output_w, off_output_b;
  int64_t off_kernel[64], off_bias[64];
  // optional per-kernel timing of the recurrent kernels (CUDA events on the launch stream)
  int timing;
  cudaEvent_t ev[2][64][2];     // [fwd|bwd][layer][start|stop]
  int ev_valid[2][64];
  // tensor-core path (H % 64 == 0, B <= 64, weights fit in shared memory); else FFMA kernels
  int use_tc;
  // side stream for the weight-gradient work that overlaps the next layer's recurrence
  cudaStream_t side;
  cudaEvent_t ev_rec[64], ev_side[64], ev_fork;
  int side_ready;
  unsigned long long* dbg_fwd;   // optional device buffers [T][8] for kernel timelines (layer 0)
  unsigned long long* dbg_bwd;
  rs::RecTcGeom tc;
};


namespace rs {

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
