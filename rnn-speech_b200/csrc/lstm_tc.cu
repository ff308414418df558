// Tensor-core path of the acoustic model: every GEMM on tcgen05 (gemm_tc.cu), the
// recurrences in the persistent tcgen05 kernels (lstm_rec_tc.cu).  Same contract and
// semantics as the FFMA path in lstm.cu (/root/reference/models/AcousticModel.py:189-317,
// :386-401; oracle/model.py).
//
// Activations travel between kernels as "planes" (x ~= hi + lo, two bf16 arrays): that is
// what TMA feeds to the tensor cores, and with the bf16x3 product it carries fp32-grade
// accuracy through the forward pass.  The backward recurrence dh_{t-1} = dgates_t Wh^T
// uses plain bf16 operands (gradient-grade accuracy); all other backward GEMMs are bf16x3.
#include "lstm_internal.cuh"
#include "gemm_tc.cuh"
#include "tc_common.cuh"

namespace rs {
namespace {

typedef __nv_bfloat16 bf16;

struct Bump {
  char* base;
  size_t off;
  explicit Bump(void* p) : base((char*)p), off(0) {}
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 1024);
    T* p = base ? (T*)(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

inline int up8(int x) { return (x + 7) / 8 * 8; }

struct TcBufs {
  // ---- workspace
  unsigned* barrier;
  float* gx;                      // rec layout [T][4H][Bpad]
  bf16 *wx_hi[64], *wx_lo[64];    // K[:H]^T  [4H][H]
  bf16 *wrec_hi[64], *wrec_lo[64];// [4H][H] permuted rows
  bf16 *wi_hi, *wi_lo;            // w_i^T [H][Fp]
  bf16 *wo_hi, *wo_lo;            // w_o^T [C][H]
  // backward-only workspace
  bf16 *wxs_hi[64], *wxs_lo[64];  // K[:H] as stored [H][4H]
  bf16 *whs_hi[64];               // K[H:] as stored [H][4H] (hi)
  bf16 *wos_hi, *wos_lo;          // w_o as stored [H][Cp]
  bf16 *dg_hi[2], *dg_lo[2];      // [T*B][4H], ping-pong over layers (the side stream still reads layer l's)
  bf16 *dgT_hi, *dgT_lo;          // [4H][TBp]
  bf16 *actT_hi, *actT_lo;        // [H][TBp] transposed activations
  bf16 *dl_hi, *dl_lo;            // dlogits planes [T*B][Cp]
  bf16 *dlT_hi, *dlT_lo;          // [C][TBp]
  bf16 *xT_hi, *xT_lo;            // [F][TBp]
  float *dcur, *dtmp;             // [T*B][H]
  // ---- reserve
  bf16 *x_hi, *x_lo;              // [T*B][Fp]
  bf16 *xin_hi[64], *xin_lo[64];  // [T*B][H]
  bf16 *hp_hi[64], *hp_lo[64];    // [(T+1)*B][H]
  float *gates[64], *cs[64];
  bf16 *top_hi, *top_lo;
  float* state0;                  // [L,2,B,H]
};

// One carve function defines the layout for sizing (null bases) and for use.
void carve(const rs_am* am, void* reserve, void* ws, bool training, TcBufs* b, size_t* res_bytes, size_t* ws_bytes) {
  const int L = am->L, H = am->H, F = am->F, C = am->C, B = am->B, T = am->Tmax;
  const size_t TB = (size_t)T * B, TBp = (size_t)up8((int)TB);
  const int Fp = up8(F), Cp = up8(C);
  Bump w(ws);
  b->barrier = w.take<unsigned>(256);
  b->gx = w.take<float>((size_t)T * 4 * H * am->tc.Bpad);
  for (int l = 0; l < L; ++l) {
    b->wx_hi[l] = w.take<bf16>((size_t)4 * H * H); b->wx_lo[l] = w.take<bf16>((size_t)4 * H * H);
    b->wrec_hi[l] = w.take<bf16>((size_t)4 * H * H); b->wrec_lo[l] = w.take<bf16>((size_t)4 * H * H);
    b->wxs_hi[l] = w.take<bf16>((size_t)4 * H * H); b->wxs_lo[l] = w.take<bf16>((size_t)4 * H * H);
    b->whs_hi[l] = w.take<bf16>((size_t)4 * H * H);
  }
  b->wi_hi = w.take<bf16>((size_t)H * Fp); b->wi_lo = w.take<bf16>((size_t)H * Fp);
  b->wo_hi = w.take<bf16>((size_t)C * H); b->wo_lo = w.take<bf16>((size_t)C * H);
  b->wos_hi = w.take<bf16>((size_t)H * Cp); b->wos_lo = w.take<bf16>((size_t)H * Cp);
  for (int i = 0; i < 2; ++i) { b->dg_hi[i] = w.take<bf16>(TB * 4 * H); b->dg_lo[i] = w.take<bf16>(TB * 4 * H); }
  b->dgT_hi = w.take<bf16>((size_t)4 * H * TBp); b->dgT_lo = w.take<bf16>((size_t)4 * H * TBp);
  b->actT_hi = w.take<bf16>((size_t)H * TBp); b->actT_lo = w.take<bf16>((size_t)H * TBp);
  b->dl_hi = w.take<bf16>(TB * Cp); b->dl_lo = w.take<bf16>(TB * Cp);
  b->dlT_hi = w.take<bf16>((size_t)C * TBp); b->dlT_lo = w.take<bf16>((size_t)C * TBp);
  b->xT_hi = w.take<bf16>((size_t)F * TBp); b->xT_lo = w.take<bf16>((size_t)F * TBp);
  b->dcur = w.take<float>(TB * H); b->dtmp = w.take<float>(TB * H);
  // activations: in the reserve when training, behind the workspace otherwise
  Bump r(reserve);
  Bump& act = training ? r : w;
  b->x_hi = act.take<bf16>(TB * Fp); b->x_lo = act.take<bf16>(TB * Fp);
  for (int l = 0; l < L; ++l) {
    b->xin_hi[l] = act.take<bf16>(TB * H); b->xin_lo[l] = act.take<bf16>(TB * H);
    if (am->tc.ts) { b->hp_hi[l] = act.take<bf16>(2 * (TB + B) * H); b->hp_lo[l] = b->hp_hi[l] + H; }   // [rows][hi | lo]
    else { b->hp_hi[l] = act.take<bf16>((TB + B) * H); b->hp_lo[l] = act.take<bf16>((TB + B) * H); }
    if (training && am->tc.ts) { b->gates[l] = act.take<float>(rec_ts_blob_floats(am->tc, T)); b->cs[l] = nullptr; }
    else if (training) { b->gates[l] = act.take<float>(TB * 4 * H); b->cs[l] = act.take<float>(TB * H); }
    else { b->gates[l] = nullptr; b->cs[l] = nullptr; }
  }
  b->top_hi = act.take<bf16>(TB * H); b->top_lo = act.take<bf16>(TB * H);
  b->state0 = act.take<float>((size_t)L * 2 * B * H);
  if (res_bytes) *res_bytes = align_up(r.off, 1024);
  if (ws_bytes) *ws_bytes = align_up(w.off, 1024);
}

// planes <- split(dropout(hi + lo)); in == out allowed
__global__ void dropout_planes_kernel(const bf16* __restrict__ ihi, const bf16* __restrict__ ilo, int cols, int ld_in,
                                      bf16* __restrict__ ohi, bf16* __restrict__ olo, int64_t n, uint64_t key, uint32_t sa,
                                      uint32_t thr_a, float inv_a, uint32_t sb, uint32_t thr_b, float inv_b) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ii = (ld_in == cols) ? i : (i / cols) * ld_in + (i % cols);
    float v = __bfloat162float(ihi[ii]) + __bfloat162float(ilo[ii]);
    if (thr_a != 0xffffffffu) v = dropout_keep(key, sa, (uint64_t)i, thr_a) ? v * inv_a : 0.f;
    if (thr_b != 0xffffffffu) v = dropout_keep(key, sb, (uint64_t)i, thr_b) ? v * inv_b : 0.f;
    bf16 h, l;
    tc::split_bf16(v, h, l);
    ohi[i] = h; olo[i] = l;
  }
}
__global__ void dropout_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n, uint64_t key,
                                   uint32_t sa, uint32_t thr_a, float inv_a, uint32_t sb, uint32_t thr_b, float inv_b) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = in[i];
    if (thr_a != 0xffffffffu) v = dropout_keep(key, sa, (uint64_t)i, thr_a) ? v * inv_a : 0.f;
    if (thr_b != 0xffffffffu) v = dropout_keep(key, sb, (uint64_t)i, thr_b) ? v * inv_b : 0.f;
    out[i] = v;
  }
}
inline uint32_t thr24(float keep) { return keep >= 1.0f ? 0xffffffffu : (uint32_t)((double)keep * 16777216.0); }
inline int ew_grid(int64_t n) {
  int grid = (int)((n + 255) / 256);
  const int cap = sm_count() * 16;
  return grid > cap ? cap : grid;
}
// in: rows of `cols` elements with row stride ld_in; out: contiguous
int dropout_planes(const bf16* ihi, const bf16* ilo, int cols, int ld_in, bf16* ohi, bf16* olo, int64_t n, uint64_t seed,
                   int sa, float keep_a, int sb, float keep_b, cudaStream_t st) {
  const uint32_t ta = sa >= 0 ? thr24(keep_a) : 0xffffffffu, tb = sb >= 0 ? thr24(keep_b) : 0xffffffffu;
  dropout_planes_kernel<<<ew_grid(n), 256, 0, st>>>(ihi, ilo, cols, ld_in, ohi, olo, n, splitmix64(seed), (uint32_t)(sa < 0 ? 0 : sa),
                                                    ta, 1.0f / keep_a, (uint32_t)(sb < 0 ? 0 : sb), tb, 1.0f / keep_b);
  RS_CHECK_LAUNCH();
  return RS_OK;
}
int dropout_f32(const float* in, float* out, int64_t n, uint64_t seed, int sa, float keep_a, int sb, float keep_b,
                cudaStream_t st) {
  const uint32_t ta = sa >= 0 ? thr24(keep_a) : 0xffffffffu, tb = sb >= 0 ? thr24(keep_b) : 0xffffffffu;
  dropout_f32_kernel<<<ew_grid(n), 256, 0, st>>>(in, out, n, splitmix64(seed), (uint32_t)(sa < 0 ? 0 : sa), ta,
                                                 1.0f / keep_a, (uint32_t)(sb < 0 ? 0 : sb), tb, 1.0f / keep_b);
  RS_CHECK_LAUNCH();
  return RS_OK;
}

// [R, C] fp32 (ld_in) -> planes [R, Cp] (ld_out >= C), zero padded columns
__global__ void split_rows_kernel(const float* __restrict__ in, int R, int C, int ld_in, bf16* __restrict__ hi,
                                  bf16* __restrict__ lo, int ld_out) {
  const int64_t n = (int64_t)R * ld_out;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / ld_out), c = (int)(i - (int64_t)r * ld_out);
    bf16 h, l;
    tc::split_bf16(c < C ? in[(size_t)r * ld_in + c] : 0.f, h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
  }
}
// [R, C] fp32 contiguous -> planes with row stride ld_out, touching only the C columns of each row
__global__ void split_rows_strided_kernel(const float* __restrict__ in, int R, int C, bf16* __restrict__ hi,
                                          bf16* __restrict__ lo, int ld_out) {
  const int64_t n = (int64_t)R * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / C), c = (int)(i - (int64_t)r * C);
    bf16 h, l;
    tc::split_bf16(in[i], h, l);
    hi[(size_t)r * ld_out + c] = h;
    lo[(size_t)r * ld_out + c] = l;
  }
}
int split_rows_strided(const float* in, int R, int C, bf16* hi, bf16* lo, int ld_out, cudaStream_t st) {
  split_rows_strided_kernel<<<ew_grid((int64_t)R * C), 256, 0, st>>>(in, R, C, hi, lo, ld_out);
  RS_CHECK_LAUNCH();
  return RS_OK;
}
int split_rows(const float* in, int R, int C, int ld_in, bf16* hi, bf16* lo, int ld_out, cudaStream_t st) {
  split_rows_kernel<<<ew_grid((int64_t)R * ld_out), 256, 0, st>>>(in, R, C, ld_in, hi, lo, ld_out);
  RS_CHECK_LAUNCH();
  return RS_OK;
}

#define RC(x) do { int _rc = (x); if (_rc != RS_OK) return _rc; } while (0)

}  // namespace

size_t am_tc_reserve_bytes(const rs_am* am) {
  TcBufs b; size_t r = 0, w = 0;
  carve(am, nullptr, nullptr, true, &b, &r, &w);
  return r;
}
size_t am_tc_workspace_bytes(const rs_am* am) {
  TcBufs b; size_t r = 0, w_train = 0, w_inf = 0;
  carve(am, nullptr, nullptr, true, &b, &r, &w_train);
  carve(am, nullptr, nullptr, false, &b, &r, &w_inf);
  return w_train > w_inf ? w_train : w_inf;
}

int am_tc_forward(rs_am* am, const float* params_d, const float* x_d, const int32_t* len_d, int T,
                  const float* state_in_d, float* state_out_d, float keep_in, float keep_out, uint64_t seed,
                  float* logits_d, void* reserve_d, void* ws_d, size_t ws_bytes, cudaStream_t st) {
  const int L = am->L, H = am->H, F = am->F, C = am->C, B = am->B;
  const int TB = T * B, Fp = up8(F);
  const int64_t nTBH = (int64_t)TB * H;
  const bool training = reserve_d != nullptr;
  TcBufs bf;
  size_t need_ws = 0;
  carve(am, reserve_d, ws_d, training, &bf, nullptr, &need_ws);
  RS_REQUIRE(ws_bytes >= need_ws, RS_ERR_WORKSPACE, "rs_am_forward: workspace %zu < %zu", ws_bytes, need_ws);
  const bool drop_in = keep_in < 1.f, drop_out = keep_out < 1.f;
  const size_t state_n = (size_t)L * 2 * B * H;
  if (state_in_d) RS_CHECK_CUDA(cudaMemcpyAsync(bf.state0, state_in_d, state_n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  else RS_CHECK_CUDA(cudaMemsetAsync(bf.state0, 0, state_n * sizeof(float), st));

  // weight planes (the parameters change every step: repack; 57 MB read at cfg-2)
  RC(split_planes_transposed(params_d + am->off_input_w, F, H, H, bf.wi_hi, bf.wi_lo, Fp, st));      // w_i^T [H][Fp]
  RC(split_planes_transposed(params_d + am->off_output_w, H, C, C, bf.wo_hi, bf.wo_lo, H, st));      // w_o^T [C][H]
  for (int l = 0; l < L; ++l) {
    const float* K = params_d + am->off_kernel[l];
    RC(split_planes_transposed(K, H, 4 * H, 4 * H, bf.wx_hi[l], bf.wx_lo[l], H, st));               // K[:H]^T [4H][H]
    RC(pack_wrec(K, H, am->tc.U, bf.wrec_hi[l], bf.wrec_lo[l], st));
  }
  // input dense -> xin[0] planes                                 (models/AcousticModel.py:247-250)
  RC(split_rows(x_d, TB, F, F, bf.x_hi, bf.x_lo, Fp, st));
  {
    SplitMat A{bf.x_hi, bf.x_lo, TB, F, Fp}, Bm{bf.wi_hi, bf.wi_lo, H, F, Fp};
    GemmTcOut o{};
    o.mode = GEMM_OUT_SPLIT; o.Chi = bf.xin_hi[0]; o.Clo = bf.xin_lo[0]; o.ldc = H; o.bias = params_d + am->off_input_b;
    RC(gemm_tc_nt(A, Bm, TB, H, F, 3, o, st));
    if (drop_in) RC(dropout_planes(bf.xin_hi[0], bf.xin_lo[0], H, H, bf.xin_hi[0], bf.xin_lo[0], nTBH, seed, 0, keep_in, -1, 1.f, st));
  }
  const int hld = am->tc.ts ? 2 * H : H;            // row stride of the h planes
  const bf16 *cur_hi = bf.xin_hi[0], *cur_lo = bf.xin_lo[0];
  int cur_ld = H;
  for (int l = 0; l < L; ++l) {
    // hoisted input half: gx = xin @ K[:H] + b, written in the recurrent kernel's layout
    {
      SplitMat A{cur_hi, cur_lo, TB, H, cur_ld}, Bm{bf.wx_hi[l], bf.wx_lo[l], 4 * H, H, H};
      GemmTcOut o{};
      o.mode = GEMM_OUT_REC; o.C = bf.gx; o.bias = params_d + am->off_bias[l];
      o.recB = B; o.recBpad = am->tc.Bpad; o.recH = H; o.recU = am->tc.U;
      RC(gemm_tc_nt(A, Bm, TB, 4 * H, H, 3, o, st));
    }
    // carried-in h -> slot 0 of the h planes
    const float* c0 = bf.state0 + ((size_t)l * 2 + 0) * B * H;
    const float* h0 = bf.state0 + ((size_t)l * 2 + 1) * B * H;
    RC(split_rows_strided(h0, B, H, bf.hp_hi[l], bf.hp_lo[l], hld, st));
    RecTcFwdArgs a;
    a.gx = bf.gx; a.wrec_hi = bf.wrec_hi[l]; a.wrec_lo = bf.wrec_lo[l];
    a.h_hi = bf.hp_hi[l]; a.h_lo = bf.hp_lo[l]; a.h_ld = hld; a.len = len_d; a.c0 = c0; a.h0 = h0;
    a.cT = state_out_d ? state_out_d + ((size_t)l * 2 + 0) * B * H : nullptr;
    a.hT = state_out_d ? state_out_d + ((size_t)l * 2 + 1) * B * H : nullptr;
    a.gates = bf.gates[l]; a.cs = bf.cs[l]; a.barrier = bf.barrier; a.T = T;
    a.dbg = (l == 0) ? am->dbg_fwd : nullptr;
    if (am->timing) RS_CHECK_CUDA(cudaEventRecord(am->ev[0][l][0], st));
    if (am->tc.ts) RC(lstm_rec_ts_forward(am->tc, a, st));
    else RC(lstm_rec_tc_forward(am->tc, a, st));
    if (am->timing) { RS_CHECK_CUDA(cudaEventRecord(am->ev[0][l][1], st)); am->ev_valid[0][l] = 1; }
    // the hop to the next consumer: identity (alias slots 1..T) or dropout(s)
    const bf16 *o_hi = bf.hp_hi[l] + (size_t)B * hld, *o_lo = bf.hp_lo[l] + (size_t)B * hld;
    const bool last = l + 1 == L;
    const bool hop_drop = last ? drop_out : (drop_out || drop_in);
    if (!hop_drop) {
      cur_hi = o_hi; cur_lo = o_lo; cur_ld = hld;
    } else {
      bf16 *d_hi = last ? bf.top_hi : bf.xin_hi[l + 1], *d_lo = last ? bf.top_lo : bf.xin_lo[l + 1];
      RC(dropout_planes(o_hi, o_lo, H, hld, d_hi, d_lo, nTBH, seed, drop_out ? 2 * l + 1 : -1, keep_out,
                        (!last && drop_in) ? 2 * (l + 1) : -1, keep_in, st));
      cur_hi = d_hi; cur_lo = d_lo; cur_ld = H;
    }
  }
  // output dense                                               (models/AcousticModel.py:308-309)
  SplitMat A{cur_hi, cur_lo, TB, H, cur_ld}, Bm{bf.wo_hi, bf.wo_lo, C, H, H};
  GemmTcOut o{};
  o.mode = GEMM_OUT_F32; o.C = logits_d; o.ldc = C; o.bias = params_d + am->off_output_b;
  return gemm_tc_nt(A, Bm, TB, C, H, 3, o, st);
}

int am_tc_backward(rs_am* am, const float* params_d, const float* x_d, const int32_t* len_d, int T, float keep_in,
                   float keep_out, uint64_t seed, const float* dlogits_d, void* reserve_d, float* grads_d,
                   void* ws_d, size_t ws_bytes, cudaStream_t st) {
  (void)x_d;
  const int L = am->L, H = am->H, F = am->F, C = am->C, B = am->B;
  const int TB = T * B, TBp = up8(TB), Fp = up8(F), Cp = up8(C);
  const int64_t nTBH = (int64_t)TB * H;
  TcBufs bf;
  size_t need_ws = 0;
  carve(am, reserve_d, ws_d, true, &bf, nullptr, &need_ws);
  RS_REQUIRE(ws_bytes >= need_ws, RS_ERR_WORKSPACE, "rs_am_backward: workspace %zu < %zu", ws_bytes, need_ws);
  RecTcBwdGeom bg;
  if (am->tc.ts) { bg.H = H; bg.B = B; bg.Bpad = am->tc.Bpad; bg.nslice = am->tc.nslice; bg.stages = 1; bg.smem_bytes = 0; }
  else RS_REQUIRE(rec_tc_bwd_geometry(H, B, &bg), RS_ERR_UNSUPPORTED, "rs_am_backward: shape outside the tensor-core path");
  const bool drop_in = keep_in < 1.f, drop_out = keep_out < 1.f;

  // weights as stored (K-major for the "multiply by W^T" GEMMs)
  RC(split_rows(params_d + am->off_output_w, H, C, C, bf.wos_hi, bf.wos_lo, Cp, st));
  for (int l = 0; l < L; ++l) {
    const float* K = params_d + am->off_kernel[l];
    RC(split_planes(K, bf.wxs_hi[l], bf.wxs_lo[l], (int64_t)H * 4 * H, st));
    RC(split_planes(K + (size_t)H * 4 * H, bf.whs_hi[l], nullptr, (int64_t)H * 4 * H, st));
  }
  // where forward left each layer's input / the top activations
  const int hld = am->tc.ts ? 2 * H : H;
  const bf16 *xin_hi[64], *xin_lo[64], *top_hi, *top_lo;
  int xin_ld[64], top_ld = H;
  xin_hi[0] = bf.xin_hi[0]; xin_lo[0] = bf.xin_lo[0]; xin_ld[0] = H;
  for (int l = 0; l < L; ++l) {
    const bool last = l + 1 == L;
    const bool hop_drop = last ? drop_out : (drop_out || drop_in);
    const bf16 *o_hi = bf.hp_hi[l] + (size_t)B * hld, *o_lo = bf.hp_lo[l] + (size_t)B * hld;
    const bf16* n_hi = hop_drop ? (last ? bf.top_hi : bf.xin_hi[l + 1]) : o_hi;
    const bf16* n_lo = hop_drop ? (last ? bf.top_lo : bf.xin_lo[l + 1]) : o_lo;
    const int n_ld = hop_drop ? H : hld;
    if (last) { top_hi = n_hi; top_lo = n_lo; top_ld = n_ld; } else { xin_hi[l + 1] = n_hi; xin_lo[l + 1] = n_lo; xin_ld[l + 1] = n_ld; }
  }

  // ---- output dense: dW_o += top^T dlogits, db_o += colsum, dtop = dlogits w_o^T
  RC(split_rows(dlogits_d, TB, C, C, bf.dl_hi, bf.dl_lo, Cp, st));
  RC(split_planes_transposed(dlogits_d, TB, C, C, bf.dlT_hi, bf.dlT_lo, TBp, st));                  // [C][TBp]
  RC(transpose_bf16(top_hi, TB, H, top_ld, bf.actT_hi, TBp, st));
  RC(transpose_bf16(top_lo, TB, H, top_ld, bf.actT_lo, TBp, st));
  {
    SplitMat A{bf.actT_hi, bf.actT_lo, H, TB, TBp}, Bm{bf.dlT_hi, bf.dlT_lo, C, TB, TBp};
    GemmTcOut o{};
    o.mode = GEMM_OUT_F32; o.C = grads_d + am->off_output_w; o.ldc = C; o.accumulate = 1;
    RC(gemm_tc_nt(A, Bm, H, C, TB, 3, o, st));
    RC(rowsum_planes(bf.dlT_hi, bf.dlT_lo, C, TB, TBp, grads_d + am->off_output_b, 1, st));
  }
  {
    SplitMat A{bf.dl_hi, bf.dl_lo, TB, C, Cp}, Bm{bf.wos_hi, bf.wos_lo, H, C, Cp};
    GemmTcOut o{};
    o.mode = GEMM_OUT_F32; o.C = bf.dcur; o.ldc = H;
    RC(gemm_tc_nt(A, Bm, TB, H, C, 3, o, st));
  }
  // Weight-gradient work (transposes, dK GEMMs, bias sums) is off the critical path: it runs on
  // a side stream, on the SMs the 48-CTA recurrent kernel of the next layer leaves idle.
  if (!am->side_ready) {
    RS_CHECK_CUDA(cudaStreamCreateWithFlags(&am->side, cudaStreamNonBlocking));
    for (int l = 0; l < L; ++l) {
      RS_CHECK_CUDA(cudaEventCreateWithFlags(&am->ev_rec[l], cudaEventDisableTiming));
      RS_CHECK_CUDA(cudaEventCreateWithFlags(&am->ev_side[l], cudaEventDisableTiming));
    }
    RS_CHECK_CUDA(cudaEventCreateWithFlags(&am->ev_fork, cudaEventDisableTiming));
    am->side_ready = 1;
  }
  cudaStream_t side = am->side;
  const int side_ctas = sm_count() - bg.nslice > 16 ? sm_count() - bg.nslice : 16;
  RS_CHECK_CUDA(cudaEventRecord(am->ev_fork, st));
  RS_CHECK_CUDA(cudaStreamWaitEvent(side, am->ev_fork, 0));
  for (int l = L - 1; l >= 0; --l) {
    const bool last = l + 1 == L;
    const bool hop_drop = last ? drop_out : (drop_out || drop_in);
    const float* dout = bf.dcur;
    if (hop_drop) {
      RC(dropout_f32(bf.dcur, bf.dtmp, nTBH, seed, drop_out ? 2 * l + 1 : -1, keep_out,
                     (!last && drop_in) ? 2 * (l + 1) : -1, keep_in, st));
      dout = bf.dtmp;
    }
    const int set = l & 1;
    // this dg set was last read by the side work of layer l+2
    if (l + 2 < L) RS_CHECK_CUDA(cudaStreamWaitEvent(st, am->ev_side[l + 2], 0));
    RecTcBwdArgs a;
    a.dout = dout; a.gates = bf.gates[l]; a.cs = bf.cs[l];
    a.c0 = bf.state0 + ((size_t)l * 2 + 0) * B * H;
    a.wh_hi = bf.whs_hi[l]; a.dg_hi = bf.dg_hi[set]; a.dg_lo = bf.dg_lo[set]; a.len = len_d; a.barrier = bf.barrier; a.T = T;
    a.dbg = (l == 0) ? am->dbg_bwd : nullptr;
    if (am->timing) RS_CHECK_CUDA(cudaEventRecord(am->ev[1][l][0], st));
    if (am->tc.ts) RC(lstm_rec_ts_backward(am->tc, a, st));
    else RC(lstm_rec_tc_backward(bg, a, st));
    if (am->timing) { RS_CHECK_CUDA(cudaEventRecord(am->ev[1][l][1], st)); am->ev_valid[1][l] = 1; }
    RS_CHECK_CUDA(cudaEventRecord(am->ev_rec[l], st));
    // ---- critical path: dxin = dg @ K[:H]^T feeds the next layer's recurrence
    {
      SplitMat A{bf.dg_hi[set], bf.dg_lo[set], TB, 4 * H, 4 * H}, Bm{bf.wxs_hi[l], bf.wxs_lo[l], H, 4 * H, 4 * H};
      GemmTcOut o{};
      o.mode = GEMM_OUT_F32; o.C = bf.dcur; o.ldc = H;
      RC(gemm_tc_nt(A, Bm, TB, H, 4 * H, 3, o, st));
    }
    // ---- side stream: dK[:H] += xin^T dg ; dK[H:] += hprev^T dg ; db += colsum(dg)
    RS_CHECK_CUDA(cudaStreamWaitEvent(side, am->ev_rec[l], 0));
    RC(transpose_bf16(bf.dg_hi[set], TB, 4 * H, 4 * H, bf.dgT_hi, TBp, side));
    RC(transpose_bf16(bf.dg_lo[set], TB, 4 * H, 4 * H, bf.dgT_lo, TBp, side));
    SplitMat G{bf.dgT_hi, bf.dgT_lo, 4 * H, TB, TBp};
    float* gK = grads_d + am->off_kernel[l];
    {
      RC(transpose_bf16(xin_hi[l], TB, H, xin_ld[l], bf.actT_hi, TBp, side));
      RC(transpose_bf16(xin_lo[l], TB, H, xin_ld[l], bf.actT_lo, TBp, side));
      SplitMat A{bf.actT_hi, bf.actT_lo, H, TB, TBp};
      GemmTcOut o{};
      o.mode = GEMM_OUT_F32; o.C = gK; o.ldc = 4 * H; o.accumulate = 1; o.max_ctas = side_ctas;
      RC(gemm_tc_nt(A, G, H, 4 * H, TB, 3, o, side));
    }
    {
      RC(transpose_bf16(bf.hp_hi[l], TB, H, hld, bf.actT_hi, TBp, side));     // slots 0..T-1 = h_{t-1}
      RC(transpose_bf16(bf.hp_lo[l], TB, H, hld, bf.actT_lo, TBp, side));
      SplitMat A{bf.actT_hi, bf.actT_lo, H, TB, TBp};
      GemmTcOut o{};
      o.mode = GEMM_OUT_F32; o.C = gK + (size_t)H * 4 * H; o.ldc = 4 * H; o.accumulate = 1; o.max_ctas = side_ctas;
      RC(gemm_tc_nt(A, G, H, 4 * H, TB, 3, o, side));
    }
    RC(rowsum_planes(bf.dgT_hi, bf.dgT_lo, 4 * H, TB, TBp, grads_d + am->off_bias[l], 1, side));
    RS_CHECK_CUDA(cudaEventRecord(am->ev_side[l], side));
  }
  // join: the input-dense gradient below reuses actT, and the caller's stream owns grads_d afterwards
  for (int l = 0; l < L && l < 2; ++l) RS_CHECK_CUDA(cudaStreamWaitEvent(st, am->ev_side[l], 0));
  // through layer 0's input dropout, then the input dense: dw_i += x^T drnn, db_i += colsum
  const float* drnn = bf.dcur;
  if (drop_in) {
    RC(dropout_f32(bf.dcur, bf.dtmp, nTBH, seed, 0, keep_in, -1, 1.f, st));
    drnn = bf.dtmp;
  }
  RC(split_planes_transposed(drnn, TB, H, H, bf.actT_hi, bf.actT_lo, TBp, st));                     // [H][TBp]
  RC(transpose_bf16(bf.x_hi, TB, F, Fp, bf.xT_hi, TBp, st));
  RC(transpose_bf16(bf.x_lo, TB, F, Fp, bf.xT_lo, TBp, st));
  SplitMat A{bf.xT_hi, bf.xT_lo, F, TB, TBp}, Bm{bf.actT_hi, bf.actT_lo, H, TB, TBp};
  GemmTcOut o{};
  o.mode = GEMM_OUT_F32; o.C = grads_d + am->off_input_w; o.ldc = H; o.accumulate = 1;
  RC(gemm_tc_nt(A, Bm, F, H, TB, 3, o, st));
  return rowsum_planes(bf.actT_hi, bf.actT_lo, H, TB, TBp, grads_d + am->off_input_b, 1, st);
}

}  // namespace rs
