// Acoustic model (input dense -> L x LSTM -> output dense) forward / backward
// orchestration behind the rs_am_* C ABI.
//
// Replaces AcousticModel._build_base_rnn (/root/reference/models/AcousticModel.py:189-317)
// and the gradient half of _add_training_on_rnn (:386-401).  See oracle/model.py
// for the restated TF semantics this follows.
#include "common.cuh"
#include "gemm.cuh"
#include "lstm_rec.cuh"
#include <stdlib.h>

#include "lstm_internal.cuh"

namespace rs {
namespace {

// out = in * keep_mask(stream a) / keep_a * keep_mask(stream b) / keep_b
// (either factor may be disabled with thr == 0xffffffff)
__global__ void dropout2_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n, uint64_t key,
                                uint32_t sa, uint32_t thr_a, float inv_a, uint32_t sb, uint32_t thr_b,
                                float inv_b) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = in[i];
    if (thr_a != 0xffffffffu) v = dropout_keep(key, sa, (uint64_t)i, thr_a) ? v * inv_a : 0.f;
    if (thr_b != 0xffffffffu) v = dropout_keep(key, sb, (uint64_t)i, thr_b) ? v * inv_b : 0.f;
    out[i] = v;
  }
}

inline uint32_t thr24(float keep) { return keep >= 1.0f ? 0xffffffffu : (uint32_t)((double)keep * 16777216.0); }

int dropout2(const float* in, float* out, int64_t n, uint64_t seed, int sa, float keep_a, int sb, float keep_b,
             cudaStream_t st) {
  const uint32_t ta = sa >= 0 ? thr24(keep_a) : 0xffffffffu, tb = sb >= 0 ? thr24(keep_b) : 0xffffffffu;
  int grid = (int)((n + 255) / 256);
  const int cap = sm_count() * 16;
  if (grid > cap) grid = cap;
  dropout2_kernel<<<grid, 256, 0, st>>>(in, out, n, splitmix64(seed), (uint32_t)(sa < 0 ? 0 : sa), ta,
                                        1.0f / keep_a, (uint32_t)(sb < 0 ? 0 : sb), tb, 1.0f / keep_b);
  RS_CHECK_LAUNCH();
  return RS_OK;
}

// Buffer plan.  Everything is in floats; TBH = Tmax*B*H.
struct Plan {
  size_t TBH, TB4H, state, TH;
  // reserve (per layer): xin | out | gates | cs ; then top
  size_t res_layer;      // floats per layer
  size_t res_total;      // floats
  // workspace: barrier (256 B) | gx (TB4H) | bufA, bufB, bufC (TBH each) | scratch reserve for inference
  size_t ws_fixed;       // bytes before the inference reserve
  size_t ws_total;       // bytes
};

Plan make_plan(const rs_am* am) {
  Plan p;
  // every carved buffer starts on a 256-byte boundary (vector loads in the kernels)
  p.TBH = align_up((size_t)am->Tmax * am->B * am->H, 64);
  p.TB4H = 4 * p.TBH;
  p.res_layer = p.TBH /*xin*/ + p.TBH /*out*/ + p.TB4H /*gates*/ + p.TBH /*cs*/;
  p.state = align_up((size_t)am->L * 2 * am->B * am->H, 64);
  p.TH = align_up((size_t)am->Tmax * am->H, 64);
  p.res_total = p.res_layer * am->L + p.TBH /*top*/ + p.TBH /*rnn_in*/ + p.state /*initial state copy*/ + p.TH /*bn 1/std*/;
  p.ws_fixed = 256 + (p.TB4H + 3 * p.TBH) * sizeof(float);
  // inference (reserve == NULL): ping-pong activations + the state copy live in the workspace
  p.ws_total = p.ws_fixed + (3 * p.TBH + p.state) * sizeof(float);
  return p;
}

}  // namespace
}  // namespace rs

using namespace rs;

extern "C" int rs_am_create(rs_am** out, int num_layers, int hidden_size, int input_dim, int num_labels,
                            int batch_size, int max_T) {
  RS_REQUIRE(out != nullptr, RS_ERR_INVALID, "rs_am_create: out is NULL");
  RS_REQUIRE(num_layers > 0 && num_layers <= 64, RS_ERR_INVALID, "rs_am_create: num_layers %d outside [1,64]", num_layers);
  RS_REQUIRE(hidden_size > 0 && input_dim > 0 && num_labels > 1 && batch_size > 0 && max_T > 0, RS_ERR_INVALID,
             "rs_am_create: non-positive dimension");
  RS_REQUIRE(hidden_size <= kRecMaxH, RS_ERR_UNSUPPORTED, "rs_am_create: hidden_size %d > %d", hidden_size, kRecMaxH);
  RS_REQUIRE(batch_size <= kRecMaxB, RS_ERR_UNSUPPORTED, "rs_am_create: batch_size %d > %d", batch_size, kRecMaxB);
  rs_am* am = new rs_am();
  am->L = num_layers; am->H = hidden_size; am->F = input_dim; am->C = num_labels;
  am->B = batch_size; am->Tmax = max_T;
  am->normalization = 0;
  int64_t off = 0;
  const int64_t H = hidden_size;
  am->off_input_w = off; off += (int64_t)input_dim * H;
  am->off_input_b = off; off += H;
  for (int l = 0; l < num_layers; ++l) {
    am->off_kernel[l] = off; off += 2 * H * 4 * H;
    am->off_bias[l] = off; off += 4 * H;
  }
  am->off_output_w = off; off += H * num_labels;
  am->off_output_b = off; off += num_labels;
  am->n_params = off;
  am->timing = 0;
  am->dbg_fwd = am->dbg_bwd = nullptr;
  am->streams_ready = 0;
  am->tev_base_ready = 0;
  am->ev_next = 0;
  {
    const char* v = getenv("RS_TC_CHUNK");
    am->chunk = v ? atoi(v) : 128;
    am->chunk_is_default = v ? 0 : 1;
    am->params_version = 0;
    am->packed_version[0] = am->packed_version[1] = 0;
    am->packed_ws[0] = am->packed_ws[1] = nullptr;
    am->packed_params[0] = am->packed_params[1] = nullptr;
    v = getenv("RS_TC_WINDOW");
    am->window = v ? atoi(v) : 2;
    if (am->window < 1) am->window = 1;
  }
  // tensor-core path when the shape fits it (RS_DISABLE_TC=1 forces the FFMA kernels)
  {
    RecTcBwdGeom bg;
    const char* off = getenv("RS_DISABLE_TC");
    const bool allow = !(off && off[0] == '1');
    // first choice: weights resident in tensor memory (lstm_rec_ts.cu); else in shared memory
    am->use_tc = allow && (rec_ts_geometry(hidden_size, batch_size, &am->tc) ||
                           (rec_tc_geometry(hidden_size, batch_size, &am->tc) &&
                            rec_tc_bwd_geometry(hidden_size, batch_size, &bg)));
  }
  for (int d = 0; d < 4; ++d)
    for (int l = 0; l < 64; ++l) am->tev_used[d][l] = 0;
  *out = am;
  return RS_OK;
}

extern "C" void rs_am_destroy(rs_am* am) {
  if (!am) return;
  for (int d = 0; d < 4; ++d)
    for (int l = 0; l < 64; ++l)
      for (cudaEvent_t e : am->tev[d][l]) cudaEventDestroy(e);
  for (cudaEvent_t e : am->evpool) cudaEventDestroy(e);
  if (am->tev_base_ready) { cudaEventDestroy(am->tev_base[0]); cudaEventDestroy(am->tev_base[1]); }
  if (am->streams_ready) {
    for (int l = 0; l < am->L; ++l) cudaStreamDestroy(am->lane[l]);
    cudaStreamDestroy(am->gemm_st);
    cudaStreamDestroy(am->side);
    cudaStreamDestroy(am->tr_st);
  }
  delete am;
}

// Batch normalisation of the stack's input over the batch axis (models/AcousticModel.py:253-259; the reference's
// `normalization` constructor argument).  Set before the first forward; the statistics are per call (no moving
// averages in the reference either).
extern "C" int rs_am_set_normalization(rs_am* am, int enable) {
  RS_REQUIRE(am != nullptr, RS_ERR_INVALID, "rs_am_set_normalization: NULL handle");
  am->normalization = enable ? 1 : 0;
  return RS_OK;
}

// See rs_am::params_version.  0 (the default) = always re-pack.
extern "C" int rs_am_set_params_version(rs_am* am, uint64_t version) {
  RS_REQUIRE(am != nullptr, RS_ERR_INVALID, "rs_am_set_params_version: NULL handle");
  am->params_version = version;
  return RS_OK;
}

extern "C" int rs_am_enable_timing(rs_am* am, int enable) {
  RS_REQUIRE(am != nullptr, RS_ERR_INVALID, "rs_am_enable_timing: NULL handle");
  am->timing = enable ? 1 : 0;
  return RS_OK;
}

// Debug: device buffers ([T][8] uint64 each) that receive %globaltimer stamps of CTA 0 of the
// layer-0 tensor-core recurrent kernels (see tests/gpu_diag.py timeline); NULL disables.
extern "C" int rs_am_set_debug_timeline(rs_am* am, void* fwd_d, void* bwd_d) {
  RS_REQUIRE(am != nullptr, RS_ERR_INVALID, "rs_am_set_debug_timeline: NULL handle");
  am->dbg_fwd = (unsigned long long*)fwd_d;
  am->dbg_bwd = (unsigned long long*)bwd_d;
  return RS_OK;
}

extern "C" int rs_am_recurrent_ms(rs_am* am, int backward, int layer, float* ms) {
  RS_REQUIRE(am && ms && am->timing, RS_ERR_INVALID, "rs_am_recurrent_ms: timing not enabled");
  RS_REQUIRE(layer >= 0 && layer < am->L && (backward == 0 || backward == 1), RS_ERR_INVALID,
             "rs_am_recurrent_ms: bad index");
  const int used = am->tev_used[backward][layer];
  RS_REQUIRE(used >= 2, RS_ERR_INVALID, "rs_am_recurrent_ms: kernel has not run");
  // sum over the (chunk) launches of the last call
  float total = 0.f;
  for (int i = 0; i + 1 < used; i += 2) {
    float part = 0.f;
    RS_CHECK_CUDA(cudaEventSynchronize(am->tev[backward][layer][i + 1]));
    RS_CHECK_CUDA(cudaEventElapsedTime(&part, am->tev[backward][layer][i], am->tev[backward][layer][i + 1]));
    total += part;
  }
  *ms = total;
  return RS_OK;
}

// (start, stop) of every recurrent launch of the last call, in ms after the top of that call; returns the number
// of launches written (<= max_launches) or a negative error.
extern "C" int rs_am_recurrent_trace(rs_am* am, int backward, int layer, float* start_stop_ms, int max_launches) {
  RS_REQUIRE(am && start_stop_ms && am->timing && am->tev_base_ready, RS_ERR_INVALID, "rs_am_recurrent_trace: timing not enabled");
  RS_REQUIRE(layer >= 0 && layer < am->L && backward >= 0 && backward < 4, RS_ERR_INVALID, "rs_am_recurrent_trace: bad index");
  const int used = am->tev_used[backward][layer];
  int n = 0;
  for (int i = 0; i + 1 < used && n < max_launches; i += 2, ++n) {
    RS_CHECK_CUDA(cudaEventSynchronize(am->tev[backward][layer][i + 1]));
    RS_CHECK_CUDA(cudaEventElapsedTime(&start_stop_ms[2 * n], am->tev_base[backward ? 1 : 0], am->tev[backward][layer][i]));
    RS_CHECK_CUDA(cudaEventElapsedTime(&start_stop_ms[2 * n + 1], am->tev_base[backward ? 1 : 0], am->tev[backward][layer][i + 1]));
  }
  return n;
}

extern "C" int64_t rs_am_param_count(const rs_am* am) { return am ? am->n_params : -1; }

extern "C" int64_t rs_am_param_offset(const rs_am* am, int which, int layer) {
  if (!am) return -1;
  switch (which) {
    case 0: return am->off_input_w;
    case 1: return am->off_input_b;
    case 2: return (layer >= 0 && layer < am->L) ? am->off_kernel[layer] : -1;
    case 3: return (layer >= 0 && layer < am->L) ? am->off_bias[layer] : -1;
    case 4: return am->off_output_w;
    case 5: return am->off_output_b;
  }
  return -1;
}

extern "C" size_t rs_am_reserve_bytes(const rs_am* am) {
  if (!am) return 0;
  return am->use_tc ? am_tc_reserve_bytes(am) : make_plan(am).res_total * sizeof(float);
}
extern "C" size_t rs_am_workspace_bytes(const rs_am* am) {
  if (!am) return 0;
  return am->use_tc ? am_tc_workspace_bytes(am) : make_plan(am).ws_total;
}
extern "C" int rs_am_uses_tensor_cores(const rs_am* am) { return am ? am->use_tc : 0; }

namespace {
struct Bufs {
  unsigned* barrier;
  float *gx, *bufA, *bufB, *bufC;
  float* rnn_in;
  float* top;
  float* state0;   // [L,2,B,H] copy of the initial state of this call
  float* bn_istd;  // [T,H] batch-norm 1/std (training) or nullptr
  float *xin[64], *out[64], *gates[64], *cs[64];
};

// Carve the workspace and the reserve (or, for inference, the workspace tail).
Bufs carve(const rs_am* am, const Plan& p, void* reserve, void* ws) {
  Bufs b;
  char* w = (char*)ws;
  b.barrier = (unsigned*)w;
  float* f = (float*)(w + 256);
  b.gx = f; f += p.TB4H;
  b.bufA = f; f += p.TBH;
  b.bufB = f; f += p.TBH;
  b.bufC = f; f += p.TBH;
  if (reserve) {
    float* r = (float*)reserve;
    for (int l = 0; l < am->L; ++l) {
      b.xin[l] = r; r += p.TBH;
      b.out[l] = r; r += p.TBH;
      b.gates[l] = r; r += p.TB4H;
      b.cs[l] = r; r += p.TBH;
    }
    b.top = r; r += p.TBH;
    b.rnn_in = r; r += p.TBH;
    b.state0 = r; r += p.state;
    b.bn_istd = r;
  } else {
    // inference: layer l reads xin from one ping-pong buffer and writes out to the other
    float* t0 = f; float* t1 = f + p.TBH; float* t2 = f + 2 * p.TBH;
    for (int l = 0; l < am->L; ++l) {
      b.xin[l] = (l & 1) ? t1 : t0;
      b.out[l] = (l & 1) ? t0 : t1;
      b.gates[l] = nullptr;
      b.cs[l] = nullptr;
    }
    b.top = t2;
    b.rnn_in = t2;
    b.state0 = f + 3 * p.TBH;
    b.bn_istd = nullptr;
  }
  return b;
}

// Without dropout the hop out[l] -> xin[l+1] (and out[L-1] -> top) is the identity:
// alias the buffers instead of copying.  Forward and backward apply the same rule.
void alias_identity_hops(Bufs& b, int L, bool drop_in, bool drop_out) {
  if (drop_in || drop_out) {
    if (!drop_out) b.top = b.out[L - 1];
    return;
  }
  for (int l = 0; l + 1 < L; ++l) b.xin[l + 1] = b.out[l];
  b.top = b.out[L - 1];
}
}  // namespace

extern "C" int rs_am_forward(rs_am* am, const float* params_d, const float* x_d, const int32_t* len_d, int T,
                             const float* state_in_d, float* state_out_d, float keep_in, float keep_out,
                             uint64_t seed, float* logits_d, void* reserve_d, void* ws_d, size_t ws_bytes,
                             void* stream) {
  RS_REQUIRE(am && params_d && x_d && len_d && logits_d && ws_d, RS_ERR_INVALID, "rs_am_forward: NULL argument");
  RS_REQUIRE(T > 0 && T <= am->Tmax, RS_ERR_INVALID, "rs_am_forward: T=%d outside [1,%d]", T, am->Tmax);
  RS_REQUIRE(keep_in > 0.f && keep_in <= 1.f && keep_out > 0.f && keep_out <= 1.f, RS_ERR_INVALID,
             "rs_am_forward: keep probabilities must be in (0,1]");
  tev_begin(am, 0, (cudaStream_t)stream);
  if (am->use_tc)
    return am_tc_forward(am, params_d, x_d, len_d, T, state_in_d, state_out_d, keep_in, keep_out, seed, logits_d,
                         reserve_d, ws_d, ws_bytes, (cudaStream_t)stream);
  const Plan p = make_plan(am);
  RS_REQUIRE(ws_bytes >= p.ws_total, RS_ERR_WORKSPACE, "rs_am_forward: workspace %zu < %zu", ws_bytes, p.ws_total);
  cudaStream_t st = (cudaStream_t)stream;
  const int L = am->L, H = am->H, F = am->F, C = am->C, B = am->B;
  const int TB = T * B;
  const int64_t nTBH = (int64_t)TB * H;
  Bufs bf = carve(am, p, reserve_d, ws_d);
  int rc;

  // Private copy of the initial state: state_out_d may alias state_in_d, and backward
  // must differentiate against the state this call STARTED from.
  if (state_in_d)
    RS_CHECK_CUDA(cudaMemcpyAsync(bf.state0, state_in_d, (size_t)L * 2 * B * H * sizeof(float),
                                  cudaMemcpyDeviceToDevice, st));
  else
    RS_CHECK_CUDA(cudaMemsetAsync(bf.state0, 0, (size_t)L * 2 * B * H * sizeof(float), st));

  // input dense: rnn_in = x @ w_i + b_i                         (models/AcousticModel.py:247-250)
  const bool drop_in = keep_in < 1.f, drop_out = keep_out < 1.f;
  alias_identity_hops(bf, L, drop_in, drop_out);
  float* rnn_in = drop_in ? bf.rnn_in : bf.xin[0];
  if ((rc = sgemm(0, 0, TB, H, F, x_d, F, params_d + am->off_input_w, H, rnn_in, H,
                  params_d + am->off_input_b, 0, st)) != RS_OK) return rc;
  if (am->normalization)                                        // (models/AcousticModel.py:253-259)
    if ((rc = bn_forward(rnn_in, bf.bn_istd, T, B, H, st)) != RS_OK) return rc;
  if (drop_in)
    if ((rc = dropout2(rnn_in, bf.xin[0], nTBH, seed, 0, keep_in, -1, 1.f, st)) != RS_OK) return rc;

  for (int l = 0; l < L; ++l) {
    const float* K = params_d + am->off_kernel[l];
    const float* bias = params_d + am->off_bias[l];
    // hoisted input half: gx = xin @ K[:H] + b                  (BasicLSTMCell: [x,h] @ kernel + bias)
    if ((rc = sgemm(0, 0, TB, 4 * H, H, bf.xin[l], H, K, 4 * H, bf.gx, 4 * H, bias, 0, st)) != RS_OK) return rc;
    RecFwdArgs a;
    a.gx = bf.gx;
    a.Wh = K + (size_t)H * 4 * H;
    a.len = len_d;
    a.c0 = bf.state0 + ((size_t)l * 2 + 0) * B * H;
    a.h0 = bf.state0 + ((size_t)l * 2 + 1) * B * H;
    a.cT = state_out_d ? state_out_d + ((size_t)l * 2 + 0) * B * H : nullptr;
    a.hT = state_out_d ? state_out_d + ((size_t)l * 2 + 1) * B * H : nullptr;
    a.out = bf.out[l];
    a.gates = bf.gates[l];
    a.cs = bf.cs[l];
    a.barrier = bf.barrier;
    a.T = T; a.B = B; a.H = H;
    { int trc = tev_record(am, 0, l, st); if (trc != RS_OK) return trc; }
    if ((rc = lstm_rec_forward(a, st)) != RS_OK) return rc;
    { int trc = tev_record(am, 0, l, st); if (trc != RS_OK) return trc; }
    // cell-output dropout, then the next cell's input dropout   (DropoutWrapper, :232-233)
    float* next = (l + 1 < L) ? bf.xin[l + 1] : bf.top;
    if (next == bf.out[l]) continue;   // identity hop, aliased
    if (l + 1 < L) {
      if ((rc = dropout2(bf.out[l], next, nTBH, seed, drop_out ? 2 * l + 1 : -1, keep_out,
                         drop_in ? 2 * (l + 1) : -1, keep_in, st)) != RS_OK) return rc;
    } else {
      if ((rc = dropout2(bf.out[l], next, nTBH, seed, drop_out ? 2 * l + 1 : -1, keep_out, -1, 1.f, st)) != RS_OK)
        return rc;
    }
  }
  // output dense                                               (models/AcousticModel.py:308-309)
  return sgemm(0, 0, TB, C, H, bf.top, H, params_d + am->off_output_w, C, logits_d, C,
               params_d + am->off_output_b, 0, st);
}

extern "C" int rs_am_backward(rs_am* am, const float* params_d, const float* x_d, const int32_t* len_d, int T,
                              float keep_in, float keep_out, uint64_t seed, const float* dlogits_d,
                              void* reserve_d, float* grads_d, void* ws_d, size_t ws_bytes, void* stream) {
  RS_REQUIRE(am && params_d && x_d && len_d && dlogits_d && reserve_d && grads_d && ws_d, RS_ERR_INVALID,
             "rs_am_backward: NULL argument");
  RS_REQUIRE(T > 0 && T <= am->Tmax, RS_ERR_INVALID, "rs_am_backward: T=%d outside [1,%d]", T, am->Tmax);
  RS_REQUIRE(keep_in > 0.f && keep_in <= 1.f && keep_out > 0.f && keep_out <= 1.f, RS_ERR_INVALID,
             "rs_am_backward: keep probabilities must be in (0,1]");
  tev_begin(am, 1, (cudaStream_t)stream);
  if (am->use_tc)
    return am_tc_backward(am, params_d, x_d, len_d, T, keep_in, keep_out, seed, dlogits_d, reserve_d, grads_d, ws_d,
                          ws_bytes, (cudaStream_t)stream);
  const Plan p = make_plan(am);
  RS_REQUIRE(ws_bytes >= p.ws_total, RS_ERR_WORKSPACE, "rs_am_backward: workspace %zu < %zu", ws_bytes, p.ws_total);
  cudaStream_t st = (cudaStream_t)stream;
  const int L = am->L, H = am->H, F = am->F, C = am->C, B = am->B;
  const int TB = T * B;
  const int64_t nTBH = (int64_t)TB * H;
  Bufs bf = carve(am, p, reserve_d, ws_d);
  int rc;
  const bool drop_in = keep_in < 1.f, drop_out = keep_out < 1.f;
  alias_identity_hops(bf, L, drop_in, drop_out);

  // output dense
  if ((rc = sgemm(1, 0, H, C, TB, bf.top, H, dlogits_d, C, grads_d + am->off_output_w, C, nullptr, 1, st)) != RS_OK) return rc;
  if ((rc = colsum(dlogits_d, TB, C, C, grads_d + am->off_output_b, 1, st)) != RS_OK) return rc;
  // d(top) = dlogits @ w_o^T
  float* dcur = bf.bufA;
  if ((rc = sgemm(0, 1, TB, H, C, dlogits_d, C, params_d + am->off_output_w, C, dcur, H, nullptr, 0, st)) != RS_OK) return rc;

  for (int l = L - 1; l >= 0; --l) {
    const float* K = params_d + am->off_kernel[l];
    float* gK = grads_d + am->off_kernel[l];
    // through the dropout(s) that sit between out[l] and what consumed it
    float* dout = bf.bufB;
    const bool identity_hop = (l == L - 1) ? !drop_out : (!drop_out && !drop_in);
    if (identity_hop) {
      dout = dcur;
    } else if (l == L - 1) {
      if ((rc = dropout2(dcur, dout, nTBH, seed, drop_out ? 2 * l + 1 : -1, keep_out, -1, 1.f, st)) != RS_OK) return rc;
    } else {
      if ((rc = dropout2(dcur, dout, nTBH, seed, drop_out ? 2 * l + 1 : -1, keep_out,
                         drop_in ? 2 * (l + 1) : -1, keep_in, st)) != RS_OK) return rc;
    }
    RecBwdArgs a;
    a.dout = dout;
    a.gates = bf.gates[l];
    a.cs = bf.cs[l];
    a.c0 = bf.state0 + ((size_t)l * 2 + 0) * B * H;
    a.Wh = K + (size_t)H * 4 * H;
    a.len = len_d;
    a.barrier = bf.barrier;
    a.T = T; a.B = B; a.H = H;
    { int trc = tev_record(am, 1, l, st); if (trc != RS_OK) return trc; }
    if ((rc = lstm_rec_backward(a, st)) != RS_OK) return rc;
    { int trc = tev_record(am, 1, l, st); if (trc != RS_OK) return trc; }
    const float* dg = bf.gates[l];
    // dK[:H] += xin^T @ dgates ; dK[H:] += hprev^T @ dgates ; db += colsum(dgates)
    if ((rc = sgemm(1, 0, H, 4 * H, TB, bf.xin[l], H, dg, 4 * H, gK, 4 * H, nullptr, 1, st)) != RS_OK) return rc;
    if (T > 1)
      if ((rc = sgemm(1, 0, H, 4 * H, (T - 1) * B, bf.out[l], H, dg + (size_t)B * 4 * H, 4 * H,
                      gK + (size_t)H * 4 * H, 4 * H, nullptr, 1, st)) != RS_OK) return rc;
    // t = 0 term of the recurrent half: h_{-1} is the carried-in state
    if ((rc = sgemm(1, 0, H, 4 * H, B, bf.state0 + ((size_t)l * 2 + 1) * B * H, H, dg, 4 * H,
                    gK + (size_t)H * 4 * H, 4 * H, nullptr, 1, st)) != RS_OK) return rc;
    if ((rc = colsum(dg, TB, 4 * H, 4 * H, grads_d + am->off_bias[l], 1, st)) != RS_OK) return rc;
    // dxin = dgates @ K[:H]^T
    if ((rc = sgemm(0, 1, TB, H, 4 * H, dg, 4 * H, K, 4 * H, dcur, H, nullptr, 0, st)) != RS_OK) return rc;
  }
  // through layer 0's input dropout, then the input dense
  float* drnn = dcur;
  if (drop_in) {
    drnn = bf.bufB;
    if ((rc = dropout2(dcur, drnn, nTBH, seed, 0, keep_in, -1, 1.f, st)) != RS_OK) return rc;
  }
  if (am->normalization) {
    // x_hat is what forward left in rnn_in (with input dropout) or in xin[0] (without)
    const float* xhat = drop_in ? bf.rnn_in : bf.xin[0];
    if ((rc = bn_backward(drnn, xhat, bf.bn_istd, T, B, H, st)) != RS_OK) return rc;
  }
  if ((rc = sgemm(1, 0, F, H, TB, x_d, F, drnn, H, grads_d + am->off_input_w, H, nullptr, 1, st)) != RS_OK) return rc;
  return colsum(drnn, TB, H, H, grads_d + am->off_input_b, 1, st);
}
