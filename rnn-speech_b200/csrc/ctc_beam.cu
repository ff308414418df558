// CTC beam search decoder for sm_100a.
//
// Replaces decoded, _ = tf.nn.ctc_beam_search_decoder(logits, input_seq_lengths) -- the reference's
// `prediction` (/root/reference/models/AcousticModel.py:312-314; defaults beam_width = 100,
// top_paths = 1, merge_repeated = True), which also feeds its error rate (:370).  The algorithm is
// TensorFlow's (core/util/ctc/ctc_beam_search.h: CTCBeamSearchDecoder::Step / TopPaths,
// BeamEntry::LabelSeq) with the default scorer; oracle/ctc.py::beam_search_decode restates it rule
// by rule and is what the tests compare with.  TF is absent from the reference tree: parity unpinned
// upstream.
//
// One CTA per utterance, the beam (<= 128 entries) in shared memory.  Per time step:
//   1. class scores = logits - max (- log-sum-exp), one warp;
//   2. every entry: old <- new; label-ended mass += parent's mass if the parent is still in the beam
//      (parents are found through a 256-slot hash table of the entries' prefix hashes -- an entry's
//      identity is the 64-bit hash of its label sequence, so a prefix that leaves the beam and comes
//      back is the same node, as in TF's prefix tree); blank-ended mass; total;
//   3. extensions (entry x label) that do not already exist as entries are scored; only those above
//      the lowest updated total can enter a full beam (exactly TF's is_candidate test), they are
//      compacted into a candidate list behind the updated entries;
//   4. bitonic sort of the list by (total desc, incumbents first, slot, label) -- usually a few
//      hundred elements, 8192 at worst (near-uniform scores) -- and the best beam_width become the new
//      beam; (previous slot, label) of every new entry goes to a [T][W] history in the workspace;
//   5. after the last frame the history is walked back from slot 0 (the best total).
#include "common.cuh"

namespace rs {
namespace {

constexpr int kMaxW = 128;          // beam entries
constexpr int kMaxC = 128;          // classes
constexpr int kHash = 256;          // hash-table slots (>= 2 * kMaxW)
constexpr int kThreads = 256;
constexpr int kRankMax = 256;       // candidate lists up to this size are ordered by counting ranks (one candidate per
                                    // thread, <= 256 comparisons each); longer lists go through the bitonic sort, whose
                                    // cost grows with n log^2 n instead of n^2 / 256 (round 1 ranked up to 1024: with the
                                    // flat posteriors of an untrained model that was most of the kernel's 24 ms)
constexpr int kMaxCand = 8192;      // >= kMaxW + kMaxW * (kMaxC - 1) is not needed: W * (C - 1) + W <= 8192 is checked
constexpr float kNegInf = -INFINITY;

__device__ __forceinline__ float lse2f(float a, float b) {      // TF LogSumExp
  if (a == kNegInf) return b;
  if (b == kNegInf) return a;
  return fmaxf(a, b) + log1pf(expf(-fabsf(a - b)));
}
__device__ __forceinline__ unsigned long long mix64(unsigned long long h, int label) {
  unsigned long long z = (h ^ (unsigned long long)(label + 1)) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return z ? z : 1ull;               // 0 marks an empty hash slot
}

struct Beam {                        // structure of arrays, one set per buffer
  unsigned long long node[kMaxW], parent[kMaxW];
  int label[kMaxW];
  float tot[kMaxW], blk[kMaxW], lab[kMaxW];
};

// (key desc, id asc): ids < kMaxW are updated entries (incumbents), id = kMaxW + slot * (C-1) + label otherwise
__device__ __forceinline__ bool before(float ka, int ia, float kb, int ib) {
  return (ka > kb) || (ka == kb && ia < ib);
}

// kGroups utterances per CTA (groups of kThreads threads, each with its own state and its own named barrier).  The
// kernel is latency-bound -- ~12 dependent phases per frame, 8 warps, ncu: issue slots 19 % busy (profiles/r02_ncu_*) --
// but two groups per SM were measured SLOWER (20.0 vs 16.7 ms per 32 x 998 frames), and what the training step with the
// error rate waits for is the decoder's duration, not the SMs it holds (24.7 vs 24.2 ms per step): one group.
constexpr int kGroups = 1;

struct BeamShared {
  Beam beam[2];
  float old_tot[kMaxW], old_blk[kMaxW], old_lab[kMaxW];
  unsigned long long hkey[kHash];
  int hslot[kHash];
  unsigned childmask[kMaxW][4];
  float in[kMaxC];
  int s_count;
  float s_lb;
  float vals[2 * kMaxW];
};

__device__ __forceinline__ void group_sync(int grp) {
  asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(kThreads) : "memory");
}

__global__ void __launch_bounds__(kGroups * kThreads)
ctc_beam_kernel(const float* __restrict__ logits, const int* __restrict__ len, int T, int B, int C, int W,
                int merge_repeated, int normalize, unsigned char* __restrict__ hist, int* __restrict__ out,
                int* __restrict__ out_len, float* __restrict__ out_score) {
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ BeamShared shared_state[kGroups];
  const int grp = threadIdx.x / kThreads;
  const int b = blockIdx.x * kGroups + grp, tid = threadIdx.x - grp * kThreads, lane = tid & 31, warp = tid >> 5;
  if (b >= B) return;                                          // (a whole group leaves: its barrier is its own)
  BeamShared& S = shared_state[grp];
  float* ckey = reinterpret_cast<float*>(dyn) + (size_t)grp * 2 * kMaxCand;      // [kMaxCand] per group
  int* cid = reinterpret_cast<int*>(ckey + kMaxCand);                            // [kMaxCand]
  Beam (&beam)[2] = S.beam;
  float (&old_tot)[kMaxW] = S.old_tot; float (&old_blk)[kMaxW] = S.old_blk; float (&old_lab)[kMaxW] = S.old_lab;
  unsigned long long (&hkey)[kHash] = S.hkey;
  int (&hslot)[kHash] = S.hslot;
  unsigned (&childmask)[kMaxW][4] = S.childmask;
  float (&in)[kMaxC] = S.in;
  int& s_count = S.s_count;
  float& s_lb = S.s_lb;
  float (&vals)[2 * kMaxW] = S.vals;
  const int L = min(len[b], T);
  const int blank = C - 1, nlab = C - 1;
  unsigned char* hprev = hist + (size_t)b * 2 * T * kMaxW;     // [T][kMaxW] previous slot (255 = none)
  unsigned char* hlab = hprev + (size_t)T * kMaxW;             // [T][kMaxW] appended label (255 = none)

  int cur = 0, n = 1;
  if (tid == 0) {
    Beam& r = beam[0];
    r.node[0] = 0x243F6A8885A308D3ull; r.parent[0] = 0ull; r.label[0] = -1;
    r.tot[0] = 0.f; r.blk[0] = 0.f; r.lab[0] = kNegInf;
  }
  group_sync(grp);

  for (int t = 0; t < L; ++t) {
    Beam& bm = beam[cur];
    Beam& nx = beam[cur ^ 1];
    // ---- 1. class scores
    if (warp == 0) {
      const float* row = logits + ((size_t)t * B + b) * C;
      float m = kNegInf;
      for (int k = lane; k < C; k += 32) m = fmaxf(m, row[k]);
      m = warp_max(m);
      float s = 0.f;
      if (normalize) {
        for (int k = lane; k < C; k += 32) s += expf(row[k] - m);
        s = logf(warp_sum(s));
      }
      for (int k = lane; k < C; k += 32) in[k] = (row[k] - m) - s;
    }
    // hash table of the entries' prefixes; old <- new
    for (int i = tid; i < kHash; i += kThreads) hkey[i] = 0ull;
    for (int i = tid; i < kMaxW * 4; i += kThreads) (&childmask[0][0])[i] = 0u;
    if (tid < n) { old_tot[tid] = bm.tot[tid]; old_blk[tid] = bm.blk[tid]; old_lab[tid] = bm.lab[tid]; }
    if (tid == 0) s_count = 0;
    group_sync(grp);
    if (tid < n) {
      unsigned h = (unsigned)(bm.node[tid] >> 17) & (kHash - 1);
      while (true) {
        const unsigned long long prev = atomicCAS(&hkey[h], 0ull, bm.node[tid]);
        if (prev == 0ull) { hslot[h] = tid; break; }
        h = (h + 1) & (kHash - 1);
      }
    }
    group_sync(grp);
    // ---- 2. update the entries
    float my_tot = kNegInf;
    if (tid < n) {
      const int lab = bm.label[tid];
      float nl = old_lab[tid];
      if (lab >= 0) {
        int ps = -1;
        const unsigned long long pk = bm.parent[tid];
        unsigned h = (unsigned)(pk >> 17) & (kHash - 1);
        while (hkey[h] != 0ull) {
          if (hkey[h] == pk) { ps = hslot[h]; break; }
          h = (h + 1) & (kHash - 1);
        }
        if (ps >= 0) {
          const float previous = (lab == bm.label[ps]) ? old_blk[ps] : old_tot[ps];
          nl = lse2f(nl, previous);
          atomicOr(&childmask[ps][lab >> 5], 1u << (lab & 31));
        }
        nl += in[lab];
      }
      const float nb = old_tot[tid] + in[blank];
      float nt = lse2f(nb, nl);
      if (!(nt == nt)) nt = kNegInf;      // NaN logits (a diverged model): keep the order total, the kernel memory-safe
      bm.lab[tid] = nl; bm.blk[tid] = nb; bm.tot[tid] = nt;
      ckey[tid] = nt; cid[tid] = tid;
      my_tot = nt;
    }
    // ---- 2b. a bound on what can enter the beam.  The beam_width-th best of (updated totals + each entry's
    // best new extension) is reached by at least beam_width candidates, so nothing below it can be kept.  (TF's
    // own test -- a new leaf must beat the current bottom of a full beam -- prunes less and keeps the same set.)
    {
      const int slot = tid >> 1, half = tid & 1;
      float best = kNegInf;
      if (slot < n) {
        const int lab = bm.label[slot];
        const float pt = old_tot[slot], pb = old_blk[slot];
        for (int c = half; c < nlab; c += 2) {
          if (childmask[slot][c >> 5] & (1u << (c & 31))) continue;
          best = fmaxf(best, in[c] + ((c == lab) ? pb : pt));
        }
      }
      best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, 1));
      if (half == 0) vals[kMaxW + slot] = (slot < n) ? best : kNegInf;
      if (tid < kMaxW) vals[tid] = (tid < n) ? my_tot : kNegInf;
    }
    group_sync(grp);
    {
      // rank of vals[tid] among the 2 * kMaxW values (descending, index breaks ties): one thread finds the bound
      const float v = vals[tid];
      int rank = 0;
      for (int j = 0; j < 2 * kMaxW; ++j) {
        const float u = vals[j];
        rank += (u > v || (u == v && j < tid)) ? 1 : 0;
      }
      if (tid == 0) s_lb = kNegInf;
      group_sync(grp);
      if (rank == W - 1) s_lb = v;                        // -inf when fewer than W finite values exist
    }
    group_sync(grp);
    // ---- 3. extensions
    const float lbv = s_lb;
    const int next = n * nlab;
    for (int idx = tid; idx < next; idx += kThreads) {
      const int slot = idx / nlab, c = idx - slot * nlab;
      if (childmask[slot][c >> 5] & (1u << (c & 31))) continue;       // exists as an active entry: handled in 2.
      const float previous = (c == bm.label[slot]) ? old_blk[slot] : old_tot[slot];
      const float tot = in[c] + previous;
      if (tot > kNegInf && tot >= lbv) {
        // append, one shared-memory atomic per warp and iteration (the lanes that pass agree on their positions by vote)
        const unsigned act = __activemask();
        const int leader = __ffs(act) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(&s_count, __popc(act));
        base = __shfl_sync(act, base, leader);
        const int pos = n + base + __popc(act & ((1u << lane) - 1u));
        ckey[pos] = tot;
        cid[pos] = kMaxW + idx;
      }
    }
    group_sync(grp);
    const int count = n + s_count;
    const int newn = min(W, count);
    const float* skey = ckey;
    const int* sid = cid;
    if (count <= kRankMax) {
      // ---- 4a. few candidates (the usual case): rank by counting, the best newn land in order in the upper half
      float* okey = ckey + kMaxCand / 2;
      int* oid = cid + kMaxCand / 2;
      for (int i = tid; i < count; i += kThreads) {
        const float v = ckey[i];
        const int id = cid[i];
        int rank = 0;
        for (int j = 0; j < count; ++j) rank += before(ckey[j], cid[j], v, id) ? 1 : 0;
        if (rank < newn) { okey[rank] = v; oid[rank] = id; }
      }
      skey = okey; sid = oid;
      group_sync(grp);
    } else {
      // ---- 4b. bitonic sort, best first
      int p2 = 32;
      while (p2 < count) p2 <<= 1;
      for (int i = count + tid; i < p2; i += kThreads) { ckey[i] = kNegInf; cid[i] = 0x7fffffff; }
      group_sync(grp);
      for (int k = 2; k <= p2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int i = tid; i < p2; i += kThreads) {
            const int ixj = i ^ j;
            if (ixj > i) {
              const float ka = ckey[i], kb = ckey[ixj];
              const int ia = cid[i], ib = cid[ixj];
              const bool up = (i & k) == 0;                 // this block ends up best-first
              const bool swap = up ? before(kb, ib, ka, ia) : before(ka, ia, kb, ib);
              if (swap) { ckey[i] = kb; ckey[ixj] = ka; cid[i] = ib; cid[ixj] = ia; }
            }
          }
          group_sync(grp);
        }
      }
    }
    if (tid < newn) {
      const int id = sid[tid];
      if (id < kMaxW) {                                   // an updated entry stays
        nx.node[tid] = bm.node[id]; nx.parent[tid] = bm.parent[id]; nx.label[tid] = bm.label[id];
        nx.tot[tid] = bm.tot[id]; nx.blk[tid] = bm.blk[id]; nx.lab[tid] = bm.lab[id];
        hprev[(size_t)t * kMaxW + tid] = (unsigned char)id;
        hlab[(size_t)t * kMaxW + tid] = 255;
      } else {                                            // a new extension enters
        const int idx = id - kMaxW, slot = idx / nlab, c = idx - slot * nlab;
        nx.node[tid] = mix64(bm.node[slot], c); nx.parent[tid] = bm.node[slot]; nx.label[tid] = c;
        nx.tot[tid] = skey[tid]; nx.blk[tid] = kNegInf; nx.lab[tid] = skey[tid];
        hprev[(size_t)t * kMaxW + tid] = (unsigned char)slot;
        hlab[(size_t)t * kMaxW + tid] = (unsigned char)c;
      }
    }
    n = newn;
    cur ^= 1;
    group_sync(grp);
  }

  // ---- 5. walk the history back from the best entry (slot 0), then LabelSeq(merge_repeated)
  if (tid == 0) {
    int* seq = out + (size_t)b * T;
    int m = 0, s = 0;
    for (int t = L - 1; t >= 0; --t) {
      const int lab = hlab[(size_t)t * kMaxW + s];
      if (lab != 255) seq[m++] = lab;                    // reversed order for now
      s = hprev[(size_t)t * kMaxW + s];
    }
    for (int i = 0; i < m / 2; ++i) { const int x = seq[i]; seq[i] = seq[m - 1 - i]; seq[m - 1 - i] = x; }
    int w = 0;
    for (int i = 0; i < m; ++i)
      if (!merge_repeated || i == 0 || seq[i] != seq[i - 1]) seq[w++] = seq[i];
    for (int i = w; i < T; ++i) seq[i] = -1;
    out_len[b] = w;
    if (out_score) out_score[b] = (L > 0) ? beam[cur].tot[0] : 0.f;
  }
}

}  // namespace
}  // namespace rs

using namespace rs;

extern "C" size_t rs_ctc_beam_workspace_bytes(int T, int B) {
  if (T <= 0 || B <= 0) return 0;
  return (size_t)B * 2 * T * kMaxW;
}

extern "C" int rs_ctc_beam_search(const float* logits_d, const int32_t* len_d, int T, int B, int C, int beam_width,
                                  int merge_repeated, int normalize, int32_t* out_d, int32_t* out_len_d,
                                  float* out_score_d, void* ws_d, size_t ws_bytes, void* stream) {
  RS_REQUIRE(logits_d && len_d && out_d && out_len_d && ws_d, RS_ERR_INVALID, "rs_ctc_beam_search: NULL argument");
  RS_REQUIRE(T > 0 && B > 0 && C > 1, RS_ERR_INVALID, "rs_ctc_beam_search: bad shape T=%d B=%d C=%d", T, B, C);
  RS_REQUIRE(C <= kMaxC, RS_ERR_UNSUPPORTED, "rs_ctc_beam_search: %d classes > %d", C, kMaxC);
  RS_REQUIRE(beam_width >= 1 && beam_width <= kMaxW, RS_ERR_UNSUPPORTED, "rs_ctc_beam_search: beam_width %d outside [1,%d]",
             beam_width, kMaxW);
  RS_REQUIRE(beam_width + beam_width * (C - 1) + 8 <= kMaxCand, RS_ERR_UNSUPPORTED,
             "rs_ctc_beam_search: beam_width * classes = %d candidates > %d", beam_width * C, kMaxCand);
  RS_REQUIRE(ws_bytes >= rs_ctc_beam_workspace_bytes(T, B), RS_ERR_WORKSPACE, "rs_ctc_beam_search: workspace %zu < %zu",
             ws_bytes, rs_ctc_beam_workspace_bytes(T, B));
  // The decoder runs on a side stream under the backward pass.  Its CTAs are compute loops: sharing an SM with a CTA of
  // the latency-critical recurrent kernels slows every recurrent step (measured: the training step went from 26 to 65 ms
  // when they co-resided), so the candidate list is padded to a size that gives the CTA its SM to itself.
  // (padded to 120 KB: more than a recurrent CTA leaves free on its SM)
  const size_t list_bytes = (size_t)kGroups * kMaxCand * (sizeof(float) + sizeof(int));
  const size_t smem = list_bytes > 120 * 1024 ? list_bytes : 120 * 1024;
  static bool attr_done[kMaxDevices] = {};
  const int dev = device_slot();
  if (!attr_done[dev]) {
    RS_CHECK_CUDA(cudaFuncSetAttribute(ctc_beam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done[dev] = true;
  }
  ctc_beam_kernel<<<cdiv(B, kGroups), kGroups * kThreads, smem, (cudaStream_t)stream>>>(logits_d, len_d, T, B, C, beam_width, merge_repeated ? 1 : 0,
                                                               normalize ? 1 : 0, (unsigned char*)ws_d, out_d, out_len_d,
                                                               out_score_d);
  RS_CHECK_LAUNCH();
  return RS_OK;
}
