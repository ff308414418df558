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
  // TMEM-resident variant (lstm_rec_ts.cu): U = 16, weights in tensor memory
  int ts;                         // 1: use lstm_rec_ts_forward / _backward (and the blob reserve layout)
  int nkb_t;                      // forward K-blocks resident in TMEM (the rest in shared memory)
  int gkb;                        // forward K-blocks per TMA box
  int ngl;                        // 16-column groups per epilogue thread
};

// false if the shape is outside the tensor-core path (caller falls back to the FFMA kernels)
bool rec_tc_geometry(int H, int B, RecTcGeom* g);

struct RecTcFwdArgs {
  const float* gx;                 // rec layout
  const __nv_bfloat16* wrec_hi;    // [4H][H]
  const __nv_bfloat16* wrec_lo;
  __nv_bfloat16* h_hi;             // [(T+1)*B] rows of h_ld elements (h_ld = H, or 2H for the row-interleaved
  __nv_bfloat16* h_lo;             //  [rows][hi | lo] layout of the TMEM-resident kernels, h_lo = h_hi + H)
  int h_ld;
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
  // Time-chunked launches (TMEM-resident kernels only; 0 / 0 = one launch over all T steps): this launch runs the
  // steps [t0, t0 + T) of a sequence of Ttot steps.  Every per-step array (gx, h planes, reserve blob) is indexed
  // by the absolute step; c0 / h0 hold the state after step t0 - 1 (cT / hT may alias them).
  int t0, Ttot;
  // Fused hop to the next consumer (TMEM-resident kernel only; drop_hi == nullptr: off).  The cell output -- 0 past the
  // sequence end -- goes through the cell's output dropout (stream sa) and the next cell's input dropout (stream sb)
  // (DropoutWrapper, /root/reference/models/AcousticModel.py:232-233; the mask of element (t, b, h) is
  // dropout_keep(key, stream, (t*B + b)*H + h, thr), as oracle/model.py::dropout_mask) and is written as bf16 planes
  // drop_hi / drop_lo [Ttot*B][H] (row t*B + b).  thr == 0xffffffff: that mask is off.
  __nv_bfloat16* drop_hi;
  __nv_bfloat16* drop_lo;
  unsigned long long drop_key;
  unsigned drop_sa, drop_thr_a, drop_sb, drop_thr_b;
  float drop_inv_a, drop_inv_b;
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
  const __nv_bfloat16* wh_lo;      // [H][4H]  low plane of the same (TMEM-resident kernel: bf16x3 dh recurrence);
                                   //          nullptr = one bf16 product (the shared-memory-resident kernel ignores it)
  __nv_bfloat16* dg_hi;            // [T*B][4H]
  __nv_bfloat16* dg_lo;            // [T*B][4H]
  const int* len;
  unsigned* barrier;
  int T;
  unsigned long long* dbg;         // optional [T][8] globaltimer stamps of CTA 0
  // Time-chunked launches (TMEM-resident kernels only): this launch runs the steps t0 + T - 1 down to t0 of Ttot.
  // A launch with t0 + T < Ttot starts from dgates_{t0+T} (already in the planes) and from the cell-state gradient
  // the previous launch left in dc_carry (rec_ts_dc_carry_floats() floats, private layout); every launch with
  // dc_carry != nullptr leaves its own there.
  int t0, Ttot;
  float* dc_carry;
  // Fused dropout backward of the hop above (TMEM-resident kernel only): dout is multiplied by the same masks as it is
  // read (thr == 0xffffffff: off), instead of being rewritten in place by a separate kernel.
  unsigned long long drop_key;
  unsigned drop_sa, drop_thr_a, drop_sb, drop_thr_b;
  float drop_inv_a, drop_inv_b;
  // 1: the caller has already filled this launch's rows of the dgates planes with the 0xFFFF pattern of the validated
  // exchange (the pipelined schedule fills every launch's rows on a side stream at the head of the pass, so that the
  // two 25 MB memsets do not stand between a launch and the event it waited for)
  int prefilled;
};
int lstm_rec_tc_backward(const RecTcBwdGeom& g, const RecTcBwdArgs& a, cudaStream_t st);

// TMEM-resident kernels (lstm_rec_ts.cu).  Same argument structs; `gates` is the private reserve
// blob of rec_ts_blob_floats() floats (forward writes, backward reads), `cs` is unused.
// Geometry: H % 128 == 0, B <= 64; false -> use the shared-memory-resident kernels above.
bool rec_ts_geometry(int H, int B, RecTcGeom* g);
size_t rec_ts_blob_floats(const RecTcGeom& g, int T);
size_t rec_ts_dc_carry_floats(const RecTcGeom& g);
int lstm_rec_ts_forward(const RecTcGeom& g, const RecTcFwdArgs& a, cudaStream_t st);
int lstm_rec_ts_backward(const RecTcGeom& g, const RecTcBwdArgs& a, cudaStream_t st);

// W = one [H,4H] half of the TF kernel [2H,4H] (K for the input half, K + H*4H for Wh) -> wrec planes
int pack_wrec(const float* W, int H, int U, __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t st);

// defined in gemm_tc.cu
int tmap_2d_bf16(void* map64, const __nv_bfloat16* base, int rows, int cols, int ld, int box_rows);
int tmap_store_bf16(void* map64, const __nv_bfloat16* base, int rows, int cols, int ld, int box_rows, int box_cols);
int tmap_store3_bf16(void* map64, const __nv_bfloat16* base, int d0, int d1, int d2, size_t s1_bytes, size_t s2_bytes,
                     int b0, int b1, int b2);
int tmap_stacked_bf16(void* map64, const __nv_bfloat16* base, int rows, int cols, int box_rows, int box_kb);
// same box, planes in two separate [rows][ld] arrays `plane_stride_bytes` apart (lo after hi)
int tmap_stacked2_bf16(void* map64, const __nv_bfloat16* base_hi, size_t plane_stride_bytes, int rows, int cols, int ld,
                       int box_rows, int box_kb);

}  // namespace rs
