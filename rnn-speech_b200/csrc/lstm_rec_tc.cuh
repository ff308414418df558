// Tensor-core (tcgen05) persistent recurrent LSTM kernels.
#pragma once
#include "common.cuh"
#include <cuda_bf16.h>

namespace rs {

// Layouts shared by the tensor-core path
//   activations ("planes"): x ~= hi + lo, two bf16 arrays [rows][H] row-major
//   hplanes: [(T+1)*B][H]; slot 0 = carried-in h, slot t+1 = h_t (0 for t >= len[b])
//   gx ("rec layout"): fp32 [T][H/U][4U][Bpad], row r = (unit % U)*4 + gate  (GEMM_OUT_REC)
//   wrec planes: bf16 [4H][H]: row (unit/U)*4U + (unit%U)*4 + gate, col k  = Wh[k][gate*H + unit]
struct RecTcGeom {
  int H, B, Bpad, U, nslice;      // Bpad = B rounded up to 16 (<= 64); nslice = H / U CTAs
  int stages;                     // h ring depth
  size_t smem_bytes;
};

// false if the shape is outside the tensor-core path (caller falls back to the FFMA kernels)
bool rec_tc_geometry(int H, int B, RecTcGeom* g);

struct RecTcFwdArgs {
  const float* gx;                 // rec layout
  const __nv_bfloat16* wrec_hi;    // [4H][H]
  const __nv_bfloat16* wrec_lo;
  __nv_bfloat16* h_hi;             // [(T+1)*B][H]
  __nv_bfloat16* h_lo;
  const int* len;                  // [B]
  const float* c0;                 // [B,H] fp32
  float* cT;                       // [B,H] or nullptr
  float* hT;                       // [B,H] or nullptr
  const float* h0;                 // [B,H] fp32 (for frozen rows' final state)
  float* gates;                    // [T,B,4H] fp32 activated gates (training) or nullptr
  float* cs;                       // [T,B,H] fp32 (training) or nullptr
  unsigned* barrier;
  int T;
  unsigned long long* dbg;         // optional [T][8] globaltimer stamps of CTA 0 (nullptr = off)
};

int lstm_rec_tc_forward(const RecTcGeom& g, const RecTcFwdArgs& a, cudaStream_t st);

// Backward: one layer of BPTT.  dgates are written as bf16 hi/lo planes [T*B][4H]
// (gate order i,j,f,o, pre-activation gradients); the dh recurrence
// dh_{t-1} = dgates_t @ Wh^T runs on tcgen05 with the CTA's 16 rows of Wh resident.
struct RecTcBwdGeom {
  int H, B, Bpad, nslice, stages;
  size_t smem_bytes;
};
bool rec_tc_bwd_geometry(int H, int B, RecTcBwdGeom* g);

struct RecTcBwdArgs {
  const float* dout;               // [T,B,H]  dL/d(out_t)
  const float* gates;              // [T,B,4H] activated gates saved by forward
  const float* cs;                 // [T,B,H]
  const float* c0;                 // [B,H]
  const __nv_bfloat16* wh_hi;      // [H][4H]  Wh as stored (row = hidden unit k)
  __nv_bfloat16* dg_hi;            // [T*B][4H]
  __nv_bfloat16* dg_lo;            // [T*B][4H]
  const int* len;
  unsigned* barrier;
  int T;
  unsigned long long* dbg;         // optional [T][8] globaltimer stamps of CTA 0
};
int lstm_rec_tc_backward(const RecTcBwdGeom& g, const RecTcBwdArgs& a, cudaStream_t st);

// Wh (rows H..2H-1 of the TF kernel [2H,4H]) -> wrec planes
int pack_wrec(const float* kernel, int H, int U, __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t st);

// defined in gemm_tc.cu
int tmap_2d_bf16(void* map64, const __nv_bfloat16* base, int rows, int cols, int ld, int box_rows);

}  // namespace rs
