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
//   3. extensions (entry x label) that do not already exist as entries are scored; a bound -- the
//      beam_width-th best of 256 distinct candidates, one per thread -- tells which of them can still
//      enter the beam; those are compacted into a candidate list behind the updated entries;
//   4. the list is ordered by (total desc, incumbents first, slot, label) -- rank by counting up to 256
//      elements (the usual case, also with flat posteriors), bitonic sort above -- and the best beam_width
//      become the new beam; (previous slot, label) of every new entry goes to a [T][W] history in the workspace;
//   5. after the last frame the history is walked back from slot 0 (the best total).
#include "common.cuh"
#include <stdio.h>
#include <stdlib.h>

namespace rs {
namespace {

constexpr int kMaxW = 128;          // beam entries
constexpr int kMaxC = 128;          // classes
constexpr int kHash = 256;          // hash-table slots (>= 2 * kMaxW)
constexpr int kThreads = 1024;      // 32 warps: the frame loop is a chain of ~10 short phases, each a loop over the beam / the
                                    // candidates, so more threads mean shorter loops (256 threads: 13 us per frame, most of it in loops of 39 - 256 iterations)
constexpr int kRankMax = 1024;      // candidate lists up to this size are ordered by counting ranks (kThreads / count lanes
                                    // per candidate: count^2 / kThreads comparisons per thread; the bound of step 2b keeps
                                    // the list at 200 - 350 elements); longer lists go through the bitonic sort
constexpr int kMaxCand = 8192;      // >= kMaxW + kMaxW * (kMaxC - 1) is not needed: W * (C - 1) + W <= 8192 is checked
constexpr float kNegInf = -INFINITY;
constexpr int kExt = kMaxCand / kThreads;       // extensions per thread: the host checks beam_width * classes + 8 <= kMaxCand
static_assert(kThreads == 4 * 2 * kMaxW && kThreads == kRankMax, "four lanes share one bound value; at least one lane per ranked candidate");

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

// Ordering by (score desc, id asc) as ONE unsigned 64-bit comparison: the order-preserving integer image of the score in
// the high word, the complement of the id in the low word (a larger composite = an earlier rank; ids are distinct, so
// composites are).  -0 is folded onto +0 first, as the float comparison does.
__device__ __forceinline__ unsigned long long composite(float score, unsigned id) {
  const unsigned u = __float_as_uint(score + 0.0f);
  const unsigned o = u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
  return ((unsigned long long)o << 32) | (unsigned long long)(0xffffffffu - id);
}
// 32 composites, one per lane, sorted best (largest) first across the lanes: bitonic network through shuffles
__device__ __forceinline__ unsigned long long warp_sort_desc(unsigned long long v, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, j);
      const bool first = (lane & j) == 0;                 // this lane is the earlier position of the pair
      const bool desc = (lane & k) == 0;                  // this block ends up best-first (k == 32: the whole warp)
      const bool take_max = first == desc;
      v = take_max ? (v > o ? v : o) : (v < o ? v : o);
    }
  }
  return v;
}
// number of composites > x in runs[0 .. nruns*32) (each run of 32 sorted best-first): a 5-step binary search per run
__device__ __forceinline__ int count_greater(const unsigned long long* runs, int nruns, unsigned long long x) {
  int total = 0;
  for (int r = 0; r < nruns; ++r) {
    const unsigned long long* run = runs + r * 32;
    int lo = 0;                                           // number of elements of the run known to be > x
#pragma unroll
    for (int step = 16; step > 0; step >>= 1)
      if (run[lo + step - 1] > x) lo += step;
    total += lo + ((lo < 32 && run[lo] > x) ? 1 : 0);
  }
  return total;
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
// kernel is latency-bound -- ~12 dependent phases per frame, ncu with 8 warps: issue slots 19 % busy (profiles/r02_ncu_*) --
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
                int* __restrict__ out_len, float* __restrict__ out_score, int prof) {
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ BeamShared shared_state[kGroups];
  const int grp = threadIdx.x / kThreads;
  const int b = blockIdx.x * kGroups + grp, tid = threadIdx.x - grp * kThreads, lane = tid & 31, warp = tid >> 5;
  if (b >= B) return;                                          // (a whole group leaves: its barrier is its own)
  BeamShared& S = shared_state[grp];
  float* ckey = reinterpret_cast<float*>(dyn) + (size_t)grp * 2 * kMaxCand;      // [kMaxCand] per group
  int* cid = reinterpret_cast<int*>(ckey + kMaxCand);                            // [kMaxCand]
  unsigned long long* runs = reinterpret_cast<unsigned long long*>(dyn + (size_t)kGroups * kMaxCand * 8) + (size_t)grp * kRankMax;
                                                                                 // [kRankMax] composites, sorted in runs of 32
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
  const float inv_nlab = 1.0f / (float)nlab;
  unsigned char* hprev = hist + (size_t)b * 2 * T * kMaxW;     // [T][kMaxW] previous slot (255 = none)
  unsigned char* hlab = hprev + (size_t)T * kMaxW;             // [T][kMaxW] appended label (255 = none)

  // RS_BEAM_PROF=1: thread 0 of CTA 0 sums the SM clock over the phases of a frame and prints them at the end
  long long pc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, pt0 = 0, psum_cand = 0;
  const bool profiling = prof && blockIdx.x == 0 && tid == 0;
#define BEAM_PH(k) do { if (profiling) { const long long now_ = clock64(); pc[k] += now_ - pt0; pt0 = now_; } } while (0)
  int cur = 0, n = 1;
  if (tid == 0) {
    Beam& r = beam[0];
    r.node[0] = 0x243F6A8885A308D3ull; r.parent[0] = 0ull; r.label[0] = -1;
    r.tot[0] = 0.f; r.blk[0] = 0.f; r.lab[0] = kNegInf;
  }
  group_sync(grp);

  float rowv[kMaxC / 32] = {};                         // warp 0: the lane's logits of the next frame
  if (warp == 0 && L > 0) {
    const float* row = logits + (size_t)b * C;
#pragma unroll
    for (int k = 0; k < kMaxC / 32; ++k) if (lane + 32 * k < C) rowv[k] = row[lane + 32 * k];
  }
  for (int t = 0; t < L; ++t) {
    Beam& bm = beam[cur];
    Beam& nx = beam[cur ^ 1];
    if (profiling) pt0 = clock64();
    // ---- 1. class scores
    if (warp == 0) {
      // (the frame's logits were fetched during the previous frame: a DRAM round trip off the chain of phases)
      float m = kNegInf;
#pragma unroll
      for (int k = 0; k < kMaxC / 32; ++k) if (lane + 32 * k < C) m = fmaxf(m, rowv[k]);
      m = warp_max(m);
      float s = 0.f;
      if (normalize) {
#pragma unroll
        for (int k = 0; k < kMaxC / 32; ++k) if (lane + 32 * k < C) s += expf(rowv[k] - m);
        s = logf(warp_sum(s));
      }
#pragma unroll
      for (int k = 0; k < kMaxC / 32; ++k) if (lane + 32 * k < C) in[lane + 32 * k] = (rowv[k] - m) - s;
      if (t + 1 < L) {
        const float* row = logits + ((size_t)(t + 1) * B + b) * C;
#pragma unroll
        for (int k = 0; k < kMaxC / 32; ++k) if (lane + 32 * k < C) rowv[k] = row[lane + 32 * k];
      }
    }
    // hash table of the entries' prefixes; old <- new
    for (int i = tid; i < kHash; i += kThreads) hkey[i] = 0ull;
    for (int i = tid; i < kMaxW * 4; i += kThreads) (&childmask[0][0])[i] = 0u;
    if (tid < n) { old_tot[tid] = bm.tot[tid]; old_blk[tid] = bm.blk[tid]; old_lab[tid] = bm.lab[tid]; }
    if (tid == 0) s_count = 0;
    group_sync(grp);
    BEAM_PH(0);
    if (tid < n) {
      unsigned h = (unsigned)(bm.node[tid] >> 17) & (kHash - 1);
      while (true) {
        const unsigned long long prev = atomicCAS(&hkey[h], 0ull, bm.node[tid]);
        if (prev == 0ull) { hslot[h] = tid; break; }
        h = (h + 1) & (kHash - 1);
      }
    }
    group_sync(grp);
    BEAM_PH(1);
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
    // ---- 2b. a bound on what can enter the beam.  The candidates -- the updated entries and the new extensions (slot, c),
    // flat index idx = slot * nlab + c -- are dealt out to 256 groups by idx mod 256 and the best of every group is taken:
    // 256 DISTINCT candidates, so the beam_width-th best of them is reached by at least beam_width candidates and nothing
    // below it can be kept.  Dealt out like this the groups are alike (each gets ~31 extensions across all slots and
    // classes), and a maximum of 31 is a top-3 % value: ~100 - 170 of the 7900 extensions pass, with peaky and with flat
    // posteriors, so the list stays within the rank-by-counting size.  (Round 2's first bound -- the updated totals and
    // each entry's best extension -- let thousands through with flat posteriors, and a partition into slots x classes
    // blocks 300 - 700; sorting them was most of the kernel's time.  TF's own test -- a new leaf must beat the current
    // bottom of a full beam -- prunes less and keeps the same set.)  The child masks of step 2 must be complete first.
    group_sync(grp);
    BEAM_PH(2);
    // this thread's share of the extensions: idx = (tid >> 2) + 256 * ((tid & 3) + 4 k), k < kExt; their scores stay in
    // registers for step 3 (-inf: no such extension, or it exists as an entry)
    float ext[kExt];
    {
      float best = my_tot;                                // (threads 0..n-1: their entry's updated total)
      const int next = n * nlab;
#pragma unroll
      for (int k = 0; k < kExt; ++k) {
        const int idx = (tid >> 2) + 256 * (tid & 3) + kThreads * k;
        float sc = kNegInf;
        if (idx < next) {
          const int slot = __float2int_rz(((float)idx + 0.5f) * inv_nlab), c = idx - slot * nlab;       // idx / nlab, exact below 2^20
          if (!(childmask[slot][c >> 5] & (1u << (c & 31)))) sc = in[c] + ((c == bm.label[slot]) ? old_blk[slot] : old_tot[slot]);
        }
        ext[k] = sc;
        best = fmaxf(best, sc);
      }
      // four neighbouring lanes share one of the 2 * kMaxW = kThreads / 4 values (their shares are disjoint: still
      // distinct candidates)
      best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, 1));
      best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, 2));
      if ((tid & 3) == 0) vals[tid >> 2] = best;
    }
    group_sync(grp);
    BEAM_PH(3);
    // the beam_width-th best of the 256 values: eight warps sort 32 each, then every value finds its rank by binary
    // searches in the eight runs (256 x 256 comparisons by counting took 8.5 of the frame's 30 thousand cycles)
    unsigned long long vcomp = 0ull;
    if (tid < 2 * kMaxW) {
      vcomp = composite(vals[tid], (unsigned)tid);
      runs[tid] = warp_sort_desc(vcomp, lane);
    }
    if (tid == 0) s_lb = kNegInf;
    group_sync(grp);
    if (tid < 2 * kMaxW && count_greater(runs, 2 * kMaxW / 32, vcomp) == W - 1) s_lb = vals[tid];      // (-inf when fewer than W finite values exist)
    group_sync(grp);
    BEAM_PH(4);
    // ---- 3. extensions that can still enter the beam go to the list behind the updated entries: positions from a
    // warp-level scan of the per-thread counts and ONE shared-memory atomic per warp
    const float lbv = s_lb;
    {
      int mine = 0;
#pragma unroll
      for (int k = 0; k < kExt; ++k) mine += (ext[k] > kNegInf && ext[k] >= lbv) ? 1 : 0;
      int incl = mine;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
      }
      int base = 0;
      if (lane == 31 && incl > 0) base = atomicAdd(&s_count, incl);
      base = __shfl_sync(0xffffffffu, base, 31);
      int pos = n + base + incl - mine;
#pragma unroll
      for (int k = 0; k < kExt; ++k)
        if (ext[k] > kNegInf && ext[k] >= lbv) {
          ckey[pos] = ext[k];
          cid[pos] = kMaxW + (tid >> 2) + 256 * (tid & 3) + kThreads * k;
          ++pos;
        }
    }
    group_sync(grp);
    BEAM_PH(5);
    const int count = n + s_count;
    if (profiling) psum_cand += count;
    const int newn = min(W, count);
    const float* skey = ckey;
    const int* sid = cid;
    if (count <= kRankMax) {
      // ---- 4a. the usual case: a thread per candidate; warps sort their 32, every candidate finds its rank in the
      // sorted runs, the best newn land in order in the upper half of the list
      float* okey = ckey + kMaxCand / 2;
      int* oid = cid + kMaxCand / 2;
      const int nruns = (count + 31) >> 5;
      const bool mine = tid < count;
      const float v = mine ? ckey[tid] : kNegInf;
      const int id = mine ? cid[tid] : 0x7fffffff;
      const unsigned long long comp = mine ? composite(v, (unsigned)id) : 0ull;          // (0 sorts behind every candidate)
      if (warp < nruns) runs[tid] = warp_sort_desc(comp, lane);
      group_sync(grp);
      if (mine) {
        const int rank = count_greater(runs, nruns, comp);
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
    BEAM_PH(6);
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
    BEAM_PH(7);
  }
  if (profiling && L > 0)
    printf("ctc_beam_kernel CTA 0: %d frames; cycles per frame: scores+clear %lld, hash insert %lld, update %lld, extension scores %lld, "
           "bound %lld, list %lld, order %lld, new beam %lld; mean list %lld\n", L, pc[0] / L, pc[1] / L, pc[2] / L, pc[3] / L,
           pc[4] / L, pc[5] / L, pc[6] / L, pc[7] / L, psum_cand / L);
#undef BEAM_PH

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
  const size_t list_bytes = (size_t)kGroups * (kMaxCand * (sizeof(float) + sizeof(int)) + kRankMax * sizeof(unsigned long long));
  const size_t smem = list_bytes > 120 * 1024 ? list_bytes : 120 * 1024;
  static bool attr_done[kMaxDevices] = {};
  const int dev = device_slot();
  if (!attr_done[dev]) {
    RS_CHECK_CUDA(cudaFuncSetAttribute(ctc_beam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done[dev] = true;
  }
  const char* prof_env = getenv("RS_BEAM_PROF");
  const int prof = (prof_env && prof_env[0] == '1') ? 1 : 0;
  ctc_beam_kernel<<<cdiv(B, kGroups), kGroups * kThreads, smem, (cudaStream_t)stream>>>(logits_d, len_d, T, B, C, beam_width, merge_repeated ? 1 : 0,
                                                               normalize ? 1 : 0, (unsigned char*)ws_d, out_d, out_len_d,
                                                               out_score_d, prof);
  RS_CHECK_LAUNCH();
  return RS_OK;
}
