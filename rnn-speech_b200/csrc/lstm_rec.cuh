// Persistent recurrent LSTM kernels (one launch per layer per direction).
#pragma once
#include "common.cuh"

namespace rs {

struct RecFwdArgs {
  const float* gx;      // [T,B,4H]  x~ @ K[:H] + bias   (gate order i,j,f,o)
  const float* Wh;      // [H,4H]    K[H:]  (recurrent half, row-major)
  const int* len;       // [B]
  const float* c0;      // [B,H] or nullptr (zeros)
  const float* h0;      // [B,H] or nullptr
  float* cT;            // [B,H] or nullptr
  float* hT;            // [B,H] or nullptr
  float* out;           // [T,B,H]  h_t for valid rows, 0 for t >= len[b]
  float* gates;         // [T,B,4H] activated gates (training) or nullptr
  float* cs;            // [T,B,H]  c_t (training) or nullptr
  unsigned* barrier;    // one zeroed counter
  int T, B, H;
};

struct RecBwdArgs {
  const float* dout;    // [T,B,H]  dL/d(out_t)
  float* gates;         // [T,B,4H] in: activated gates, out: d(pre-activation gates)
  const float* cs;      // [T,B,H]
  const float* c0;      // [B,H] or nullptr
  const float* Wh;      // [H,4H]
  const int* len;       // [B]
  unsigned* barrier;
  int T, B, H;
};

// Supported: H <= kRecMaxH, B <= kRecMaxB (RS_ERR_UNSUPPORTED otherwise).
constexpr int kRecMaxH = 1152;
constexpr int kRecMaxB = 256;

int lstm_rec_forward(const RecFwdArgs& a, cudaStream_t st);
int lstm_rec_backward(const RecBwdArgs& a, cudaStream_t st);

}  // namespace rs
