// Batched Viterbi beam search over kana lattices (decoder/decoder.py:164-241,
// decoder/decoder_dynamic.py:53-194), many independent sentences in lock-step.
//
// Everything structural is known from the lattice alone: the number of paths kept in frame t of a
// sentence is min(beam, sum over nodes ending at t of the count kept at the node's start frame), so
// the host plans every beam slot, LM row and candidate offset up front and the device never
// synchronises with the host inside the frame loop.  Beam slots are laid out frame-major (all
// sentences' frame t are contiguous) so one LM step is one GEMM over a contiguous row range.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <numeric>
#include <string>
#include <thread>
#include <unordered_map>

#include <cooperative_groups.h>

#include "jlm_beam.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// device kernels
// ------------------------------------------------------------------------------------------------
__global__ void k_init_frame0(BeamDev d, int S) {
  pdl_enter();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= S) return;
  const int fid = (int)d.fbase[p];
  const int64_t s0 = d.slot0[fid];
  const int node = d.frame_lo[fid];
  d.slot_score[s0] = 0.0;
  d.slot_lse[s0] = 0.0;
  d.slot_parent[s0] = -1;
  d.slot_node[s0] = node;
  d.slot_word[s0] = d.node_word[node];
  if (d.slot_cumy) d.slot_cumy[s0] = 0.0;
  d.guard_gap[p] = INFINITY;
  d.guard_flag[p] = 0;
  if (p == 0) *d.guard_n = 0;
}

__global__ void k_build_items(BeamDev d) {
  pdl_enter();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.n_items) return;
  const int n = d.start_items[i];
  const int pf = d.node_pfid[n];
  ScoreItem it;
  it.word = d.node_word[n];
  it.rows = d.bc[pf];
  it.node = n;
  it.pad = d.node_span ? d.node_span[n] : 0;
  it.ps0 = d.slot0[pf];
  it.cpos = d.cand_pos[n];
  d.items[i] = it;
}

// Score: Path.append_node (decoder.py:43-49) for every lattice node that STARTS at the frame that
// was just stepped.  One warp per node: the node's word row of the output block is read once and
// dotted (float64 accumulation) with the stage-1 projection row of each kept path of the start frame;
// the result lands in the node's slice of its END frame's candidate list, so the prune kernel of
// that later frame scans one contiguous, coalesced array.
//   static : cand_val = parent score + (LSE(parent row) - y[word])      (-y when self-normalised)
//   dynamic: cand_val = y[word]; scores depend on the end frame's vocabulary and are formed at prune time
constexpr int SC_RC = 8;       // rows accumulated per pass
constexpr int SC_WARPS = 8;    // warps (= nodes) per CTA
constexpr int SC_MAXPASS = 2;  // row passes whose parent scores are prefetched (beam <= 16); wider beams load late

__device__ __forceinline__ void load4(const float* p, double (&t)[4]) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
}
__device__ __forceinline__ void load4(const double* p, double (&t)[4]) {
  const double2 a = reinterpret_cast<const double2*>(p)[0], b = reinterpret_cast<const double2*>(p)[1];
  t[0] = a.x; t[1] = a.y; t[2] = b.x; t[3] = b.y;
}

// One warp scores work item `item` of the step (k_score_nodes: warp-per-item kernel; k_single_f64: its scoring phase).
// lse_rows: the log-sum-exp of the start frame's rows by rank when they are not in d.slot_lse yet (k_single_f64 keeps
// them in shared memory), else nullptr.
template <typename TT, bool DYN>
__device__ __forceinline__ void score_item(const SegTable& seg, const TT* T, int64_t ldt, const BeamDev& d,
                                           const float* __restrict__ b2, int64_t item0, int item, int64_t row0, int use_lse,
                                           int defer, const double* lse_rows, int t_step = 0, int tstride = 0,
                                           const TT* tile0 = nullptr, int64_t tile_ps0 = -1, int tile_seg0 = -1,
                                           const TT* tile1 = nullptr, int64_t tile_ps1 = -1, int tile_seg1 = -1) {
  const int lane = threadIdx.x & 31;
  const int4* ip = reinterpret_cast<const int4*>(d.items + item0 + item);
  const int4 i0 = __ldg(ip), i1 = __ldg(ip + 1);
  const int w = i0.x;
  const int rows = i0.y;
  const int64_t ps0 = ((int64_t)(uint32_t)i1.x) | ((int64_t)i1.y << 32);
  const int64_t cpos = ((int64_t)(uint32_t)i1.z) | ((int64_t)i1.w << 32);
  // the parent rows' (score, LSE) this lane will combine with its dot products: fetched now, used last
  const int my_r = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  double pre[SC_MAXPASS];
#pragma unroll
  for (int q = 0; q < SC_MAXPASS; ++q) {
    pre[q] = 0.0;
    const int r = q * SC_RC + my_r;
    if (!DYN && !defer && (lane & 3) == 0 && r < rows)
      pre[q] = d.slot_score[ps0 + r] + (use_lse ? (lse_rows ? lse_rows[r] : d.slot_lse[ps0 + r]) : 0.0);
  }
  int s = 0;
#pragma unroll
  for (int i = 1; i < JLM_MAX_SEGMENTS; ++i)
    if (i < seg.n && w >= seg.start[i]) s = i;
  const int kpad = seg.kpad[s];
  const float* wrow = seg.W[s] + (int64_t)(w - seg.start[s]) * kpad;
  const TT* trow = T + (ps0 - row0) * ldt + seg.koff[s];
  // the start frame's stage-1 rows staged in shared memory by the CTA (k_score_nodes): [rows][kpad]
  if (tile0 && ps0 == tile_ps0 && s == tile_seg0) {
    trow = tile0;
    ldt = kpad;
  } else if (tile1 && ps0 == tile_ps1 && s == tile_seg1) {
    trow = tile1;
    ldt = kpad;
  }
  const double bias = (double)b2[w];
  for (int r0 = 0; r0 < rows; r0 += SC_RC) {
    double acc[SC_RC];
#pragma unroll
    for (int r = 0; r < SC_RC; ++r) acc[r] = 0.0;
    for (int k = lane * 4; k < kpad; k += 128) {
      const float4 wv = *reinterpret_cast<const float4*>(wrow + k);
      const double w0 = (double)wv.x, w1 = (double)wv.y, w2 = (double)wv.z, w3 = (double)wv.w;
#pragma unroll
      for (int r = 0; r < SC_RC; ++r) {
        if (r0 + r < rows) {
          double t[4];
          load4(trow + (int64_t)(r0 + r) * ldt + k, t);
          acc[r] = fma(t[0], w0, acc[r]);
          acc[r] = fma(t[1], w1, acc[r]);
          acc[r] = fma(t[2], w2, acc[r]);
          acc[r] = fma(t[3], w3, acc[r]);
        }
      }
    }
    // reduce-scatter over the warp: halve the rows a lane carries at xor 16 / 8 / 4, then finish the
    // one remaining row over xor 2 / 1.  Lane l ends with row (l >> 2) & 7 (bits 4,3,2 -> row bits 2,1,0).
    static_assert(SC_RC == 8, "the reduce-scatter below is written for 8 rows");
    const unsigned FULL = 0xffffffffu;
    double v4[4], v2[2], v1;
    {
      const bool up = (lane & 16) != 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double send = up ? acc[i] : acc[i + 4], keep = up ? acc[i + 4] : acc[i];
        v4[i] = keep + __shfl_xor_sync(FULL, send, 16);
      }
    }
    {
      const bool up = (lane & 8) != 0;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const double send = up ? v4[i] : v4[i + 2], keep = up ? v4[i + 2] : v4[i];
        v2[i] = keep + __shfl_xor_sync(FULL, send, 8);
      }
    }
    {
      const bool up = (lane & 4) != 0;
      const double send = up ? v2[0] : v2[1], keep = up ? v2[1] : v2[0];
      v1 = keep + __shfl_xor_sync(FULL, send, 4);
    }
    v1 += __shfl_xor_sync(FULL, v1, 2);
    v1 += __shfl_xor_sync(FULL, v1, 1);
    const int r = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    if ((lane & 3) == 0 && r0 + r < rows) {
      const double y = v1 + bias;
      const int64_t ps = ps0 + r0 + r;
      const int q = r0 / SC_RC;
      static_assert(SC_MAXPASS == 2, "select below is written for two prefetched passes");
      if (DYN) {
        // _fix_neg_log (decoder_dynamic.py:150-175): the candidate is ranked at its END frame by (sum of the path's
        // LSEs under that frame's vocabulary) - (sum of its logits).  Both sums of the parent path are final once
        // k_dyn_prefix_lse of this step has run, so the score is formed here - the prune kernel then reads ONE
        // coalesced array instead of three dependent random loads per candidate (cfg 4: 45 -> 2x us per frame).
        d.cand_val[cpos + r0 + r] = y;
        d.cand_parent[cpos + r0 + r] = (int32_t)ps;
        // (fetching the two sums at the start of the kernel like the static path's `pre` was measured slower: 69 vs 61 us)
        d.cand_score[cpos + r0 + r] =
            (use_lse ? d.dyn_chain[ps * tstride + t_step + i0.w] : 0.0) - (d.slot_cumy[ps] + y);
      } else if (defer) {
        // the parent rows' LSE is still being computed (output GEMM on the main stream): leave -y, k_add_base
        // finishes the score with the same association, score + (LSE - y)
        d.cand_val[cpos + r0 + r] = -y;
      } else {
        double base_score;
        if (q < SC_MAXPASS) base_score = q == 0 ? pre[0] : pre[1];
        else base_score = d.slot_score[ps] + (use_lse ? (lse_rows ? lse_rows[r0 + r] : d.slot_lse[ps]) : 0.0);
        d.cand_val[cpos + r0 + r] = base_score - y;
      }
    }
  }
}

// The eight items of a CTA are consecutive in (sentence, node) order, so most of them extend the SAME start frame:
// its stage-1 rows (beam x K floats, read once per node before: 66 MB per frame at cfg 2, the kernel's bandwidth)
// are staged in shared memory once per CTA - for the first and for the last item's sentence - and the warps whose
// item belongs to one of the two read them from there.  tile_elems = capacity of one tile in TT elements (0: off).
// TILES = false is the plain warp-per-item kernel (a separate instantiation: with the staging code in the same kernel the
// untiled path ran 24 -> 33 us per frame at cfg 2).
template <typename TT, bool DYN, bool TILES>
__global__ void __launch_bounds__(SC_WARPS * 32)
k_score_nodes(SegTable seg, const TT* __restrict__ T, int64_t ldt, BeamDev d, const float* __restrict__ b2,
              int64_t item0, int n_items, int64_t row0, int use_lse, int defer, int t_step, int tstride, int tile_elems) {
  pdl_enter();
  const int item = blockIdx.x * SC_WARPS + (threadIdx.x >> 5);
  if constexpr (!TILES) {
    if (item >= n_items) return;
    score_item<TT, DYN>(seg, T, ldt, d, b2, item0, item, row0, use_lse, defer, nullptr, t_step, tstride);
    return;
  } else {
  extern __shared__ float4 sc_dyn[];
  __shared__ long long t_ps[2];
  __shared__ int t_seg[2], t_rows[2];
  const TT* tiles = reinterpret_cast<const TT*>(sc_dyn);
  {
    const int first = blockIdx.x * SC_WARPS, last = min(n_items, first + SC_WARPS) - 1;
    if (threadIdx.x < 2) {
      const ScoreItem it = d.items[item0 + (threadIdx.x == 0 ? first : last)];
      int sg = 0;
#pragma unroll
      for (int i = 1; i < JLM_MAX_SEGMENTS; ++i)
        if (i < seg.n && it.word >= seg.start[i]) sg = i;
      t_ps[threadIdx.x] = it.rows * seg.kpad[sg] <= tile_elems ? (long long)it.ps0 : -1;
      t_seg[threadIdx.x] = sg;
      t_rows[threadIdx.x] = it.rows;
    }
    __syncthreads();
    if (threadIdx.x == 0 && t_ps[1] == t_ps[0] && t_seg[1] == t_seg[0]) t_ps[1] = -1;
    __syncthreads();
    TT* wt = reinterpret_cast<TT*>(sc_dyn);
    constexpr int VEC = 16 / (int)sizeof(TT);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (t_ps[i] < 0) continue;
      const int kp = seg.kpad[t_seg[i]], n = t_rows[i] * kp;
      const TT* src = T + ((int64_t)t_ps[i] - row0) * ldt + seg.koff[t_seg[i]];
      for (int e = threadIdx.x * VEC; e < n; e += SC_WARPS * 32 * VEC) {
        const int r = e / kp, k = e - r * kp;
        *reinterpret_cast<float4*>(wt + (size_t)i * tile_elems + e) = *reinterpret_cast<const float4*>(src + (int64_t)r * ldt + k);
      }
    }
    __syncthreads();
  }
  if (item >= n_items) return;
  score_item<TT, DYN>(seg, T, ldt, d, b2, item0, item, row0, use_lse, defer, nullptr, t_step, tstride,
                      t_ps[0] >= 0 ? tiles : nullptr, t_ps[0], t_seg[0],
                      t_ps[1] >= 0 ? tiles + tile_elems : nullptr, t_ps[1], t_seg[1]);
  }
}

// Second half of a deferred k_score_nodes: cand_val holds -y; add the parent path's score and LSE.
__global__ void k_add_base(BeamDev d, int64_t item0, int n_items, int W) {
  pdl_enter();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int item = (int)(i / W), r = (int)(i % W);
  if (item >= n_items) return;
  const ScoreItem* it = d.items + item0 + item;
  if (r >= it->rows) return;
  const int64_t ps = it->ps0 + r;
  double* v = d.cand_val + it->cpos + r;
  *v = d.slot_score[ps] + (d.slot_lse[ps] + *v);
}

// Prune: one warp per sentence keeps the beam_width best candidates of the frame under the
// reference's stable sort (decoder.py:227-229).  The frame's candidates are one contiguous array in
// the reference's enumeration order (node order, then parent rank), read 128 per iteration with
// coalesced loads; a warp ballot picks the few that beat the current k-th score and a candidate only
// displaces kept entries that are strictly worse, so equal scores keep the earlier ordinal.  The kept
// list lives in registers, sorted, entry e at (lane e%32, register e/32), as (score, ordinal); the
// (node, parent) of the survivors is recovered at the end by a binary search over the frame's nodes.
// Serial selection: the W best (score, ordinal) pairs in a register-resident sorted list, entry e at
// (lane e % 32, register e / 32); candidates are read 128 per iteration, a ballot picks the few that beat the
// current W-th score and each is inserted with a shuffle shift.  Stable: a candidate only displaces entries
// that are strictly worse.  One warp.
template <int L, class ValueFn>
__device__ __forceinline__ void serial_select(ValueFn value, int nc, int W, int lane, double (&es)[L], int (&ec)[L]) {
  const unsigned FULL = 0xffffffffu;
  const int kl = (W - 1) >> 5, klane = (W - 1) & 31;
  double kth = INFINITY;

  constexpr int U = 4;   // 32-candidate groups in flight per iteration
  for (int base = 0; base < nc; base += 32 * U) {
    double v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int c = base + u * 32 + lane;
      v[u] = INFINITY;
      if (c < nc) v[u] = value(c);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      unsigned m = __ballot_sync(FULL, v[u] < kth);
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const double x = __shfl_sync(FULL, v[u], b);
        if (!(x < kth)) continue;
        const int xc = base + u * 32 + b;
        int pos = 0;
#pragma unroll
        for (int l = 0; l < L; ++l) pos += __popc(__ballot_sync(FULL, es[l] <= x));
#pragma unroll
        for (int l = L - 1; l >= 0; --l) {
          const int idx = l * 32 + lane;
          double us = __shfl_up_sync(FULL, es[l], 1);
          int uc = __shfl_up_sync(FULL, ec[l], 1);
          if (l > 0) {
            const double cs = __shfl_sync(FULL, es[l > 0 ? l - 1 : 0], 31);
            const int cc = __shfl_sync(FULL, ec[l > 0 ? l - 1 : 0], 31);
            if (lane == 0) {
              us = cs;
              uc = cc;
            }
          }
          if (idx > pos) {
            es[l] = us;
            ec[l] = uc;
          } else if (idx == pos) {
            es[l] = x;
            ec[l] = xc;
          }
        }
        double kv = es[0];
#pragma unroll
        for (int l = 1; l < L; ++l)
          if (l == kl) kv = es[l];
        kth = __shfl_sync(FULL, kv, klane);
      }
    }
  }
}

constexpr int PRUNE_CAP = 96;   // survivors of the threshold filter a warp can rank in shared memory

// One warp prunes frame t of sorted sentence `warp` (k_prune: warp-per-sentence kernel; k_single_f64: warp 0 of CTA 0).
template <int L, bool DYN>
__device__ __forceinline__ void prune_sentence(const BeamDev& d, int t, int warp, int W, int tstride, int use_lse) {
  const int lane = threadIdx.x & 31;
  const unsigned FULL = 0xffffffffu;
  const int fid = (int)d.fbase[warp] + t;
  const int lo = d.frame_lo[fid], hi = d.frame_hi[fid];
  const int64_t c0 = d.frame_cand_lo[fid];
  const int nc = d.frame_ncand[fid];
  const double* val = d.cand_val + c0;

  // DYN, _fix_neg_log (decoder_dynamic.py:150-175): every ancestor transition is re-scored with the
  // softmax over lattice_vocab[t].  dyn_chain[slot][t] already holds the sum of the path's LSEs under
  // that vocabulary (accumulated parent-to-child by k_dyn_prefix_lse), so a candidate's score is
  // (sum of LSEs) - (sum of logits) along its path, formed here from its parent slot.

  double es[L];
  int ec[L];
#pragma unroll
  for (int l = 0; l < L; ++l) {
    es[l] = INFINITY;
    ec[l] = -1;
  }
  auto value = [&](int c) -> double {
    if (DYN) return d.cand_score[c0 + c];      // formed by k_score_nodes (same expression, same operands)
    return val[c];
  };

  // ---- parallel selection ---------------------------------------------------------------------------
  // (1) every lane takes L minima over its strided share of the candidates; the W-th smallest of those
  //     32 L >= W values is an upper bound tau of the frame's W-th best score (they are candidates).
  // (2) the candidates <= tau - about 2W of them - are compacted, in ordinal order, into shared memory.
  // (3) each survivor counts the survivors that sort before it under (score, ordinal): that count is its
  //     rank in the reference's stable sort (decoder.py:227-229), and ranks < W are the kept paths.
  // No step has a serial dependence on the candidate count; the insertion list below remains for wider
  // beams and for the rare frame whose survivors overflow the buffer (many exact ties).
  bool selected = false;
  {
    constexpr int CAP = PRUNE_CAP * L;
    __shared__ double sv_all[4][CAP];
    __shared__ int sc_all[4][CAP];
    __shared__ double ov_all[4][32 * L];
    __shared__ int oc_all[4][32 * L];
    double* sv = sv_all[threadIdx.x >> 5];
    int* sc = sc_all[threadIdx.x >> 5];
    double* ov = ov_all[threadIdx.x >> 5];
    int* oc = oc_all[threadIdx.x >> 5];
    // loads are issued sixteen groups at a time: a frame with thousands of candidates (one long-tailed
    // sentence per batch sets the kernel's duration) must not pay one L2 round trip per 32 of them.
    // A lane keeps L minima (32-candidate group g goes to slot g % L) so that 32 L >= W values are ranked.
    constexpr int UN = 16;
    static_assert(UN % L == 0, "slot of a group must be a compile-time function of the unroll index");
    double lmin[L];
#pragma unroll
    for (int k = 0; k < L; ++k) lmin[k] = INFINITY;
    for (int base = 0; base < nc; base += 32 * UN) {
      double v[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int c = base + u * 32 + lane;
        v[u] = c < nc ? value(c) : INFINITY;
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) lmin[u % L] = fmin(lmin[u % L], v[u]);
    }
    int rk[L];
#pragma unroll
    for (int k = 0; k < L; ++k) rk[k] = 0;
#pragma unroll 4
    for (int j = 0; j < 32; ++j) {
#pragma unroll
      for (int k2 = 0; k2 < L; ++k2) {
        const double o = __shfl_sync(FULL, lmin[k2], j);
        const int oid = j * L + k2;
#pragma unroll
        for (int k = 0; k < L; ++k) rk[k] += (o < lmin[k] || (o == lmin[k] && oid < lane * L + k)) ? 1 : 0;
      }
    }
    double tau = INFINITY;
#pragma unroll
    for (int k = 0; k < L; ++k) {
      const unsigned holder = __ballot_sync(FULL, rk[k] == W - 1);
      if (holder) tau = __shfl_sync(FULL, lmin[k], __ffs(holder) - 1);
    }
    int ns = 0;
    for (int base = 0; base < nc; base += 32 * UN) {
      double v[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int c = base + u * 32 + lane;
        v[u] = c < nc ? value(c) : INFINITY;
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int c = base + u * 32 + lane;
        const bool keep = c < nc && v[u] <= tau;
        const unsigned m = __ballot_sync(FULL, keep);
        const int pos = ns + __popc(m & ((1u << lane) - 1u));
        if (keep && pos < CAP) {
          sv[pos] = v[u];
          sc[pos] = c;
        }
        ns += __popc(m);
      }
    }
    if (ns <= CAP) {
      __syncwarp();
      for (int i = lane; i < ns; i += 32) {
        const double x = sv[i];
        int r = 0;
        for (int j = 0; j < ns; ++j) {
          const double y = sv[j];
          r += (y < x || (y == x && j < i)) ? 1 : 0;
        }
        if (r < W) {          // ranks are a permutation of 0..ns-1: each kept slot is written exactly once
          ov[r] = x;
          oc[r] = sc[i];
        }
      }
      __syncwarp();
#pragma unroll
      for (int l = 0; l < L; ++l) {
        const int idx = l * 32 + lane;
        if (idx < W && idx < ns) {
          es[l] = ov[idx];
          ec[l] = oc[idx];
        }
      }
      selected = true;
    }
  }

  if (!selected) serial_select<L>(value, nc, W, lane, es, ec);

  const int cnt = d.bc[fid];
  const int64_t s0 = d.slot0[fid];
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const int idx = l * 32 + lane;
    if (idx < cnt) {
      const int64_t c = c0 + ec[l];
      // last node of the frame whose first candidate is <= c
      int a = lo, b = hi - 1;
      while (a < b) {
        const int mid = (a + b + 1) >> 1;
        if (d.cand_pos[mid] <= c) a = mid; else b = mid - 1;
      }
      const int par = (int)(d.slot0[d.node_pfid[a]] + (c - d.cand_pos[a]));
      d.slot_score[s0 + idx] = es[l];
      d.slot_parent[s0 + idx] = par;
      d.slot_node[s0 + idx] = a;
      d.slot_word[s0 + idx] = d.node_word[a];
      if (DYN) d.slot_cumy[s0 + idx] = d.slot_cumy[par] + d.cand_val[c];
    }
  }
}

template <int L, bool DYN>
__global__ void __launch_bounds__(128)
k_prune(BeamDev d, int t, int nact, int W, int tstride, int use_lse) {
  pdl_enter();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= nact) return;
  prune_sentence<L, DYN>(d, t, warp, W, tstride, use_lse);
}

// Block-per-sentence form of the same selection (W <= 128): the four warps split the candidate range, so the one
// long-tailed sentence that sets the kernel's duration (thousands of candidates) has four times the loads in
// flight, and the grid has four times the warps of the warp-per-sentence kernel.  Steps as in k_prune: thread
// minima -> tau (rank W-1 among the 128 minima) -> per-warp ordered compaction of the candidates <= tau (warp w owns
// a contiguous quarter of the ordinals, so warp order is ordinal order) -> rank by counting under (score, ordinal)
// -> kept list in shared memory -> slots.  A frame whose survivors overflow the buffers (mass ties) is finished by
// warp 0 with the insertion list.
constexpr int PB_CAP = 160;    // survivors per warp segment

template <int L, bool DYN>
__global__ void __launch_bounds__(128)
k_prune_block(BeamDev d, int t, int W, int tstride, int use_lse, double guard_eps, int guard_all, int topN) {
  pdl_enter();
  __shared__ double mins[128 * L];
  __shared__ double sv[4][PB_CAP];
  __shared__ int sc[4][PB_CAP];
  __shared__ int cnt_w[4];
  __shared__ double tau_s;
  __shared__ double ov[129];       // ranks 0..W (entry W: the best rejected candidate, for the near-tie guard)
  __shared__ int oc[129];
  __shared__ int minc[4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned FULL = 0xffffffffu;
  const int p = blockIdx.x;
  const int fid = (int)d.fbase[p] + t;
  const int lo = d.frame_lo[fid], hi = d.frame_hi[fid];
  const int64_t c0 = d.frame_cand_lo[fid];
  const int nc = d.frame_ncand[fid];
  const double* val = d.cand_val + c0;
  auto value = [&](int c) -> double {
    if (DYN) return d.cand_score[c0 + c];      // formed by k_score_nodes (same expression, same operands)
    return val[c];
  };
  // warp w owns ordinals [w*q, min(nc, (w+1)*q)), q a multiple of 32
  const int q = ((nc + 127) / 128) * 32;
  const int w_lo = warp * q, w_hi = min(nc, w_lo + q);
  constexpr int UN = 8;
  // Every thread keeps its L smallest candidates: 128 L distinct candidates, of which the one ranked W-1 bounds the
  // frame's W-th best score.  One value per thread is not enough for wide beams: when the best W candidates are the
  // W parents of ONE well-scored node they sit in at most 32 lanes of one warp, the bound then comes from some
  // mediocre candidate of another thread and the survivor buffers overflow (measured at beam 50: 45 % of sentences).
  double tm[L];
#pragma unroll
  for (int i = 0; i < L; ++i) tm[i] = INFINITY;
  for (int base = w_lo; base < w_hi; base += 32 * UN) {
    double v[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int c = base + u * 32 + lane;
      v[u] = c < w_hi ? value(c) : INFINITY;
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      double x = v[u];
#pragma unroll
      for (int i = 0; i < L; ++i) {      // sorted insertion, tm[0] <= tm[1] <= ...
        const double lo_ = fmin(tm[i], x);
        x = fmax(tm[i], x);
        tm[i] = lo_;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < L; ++i) mins[tid * L + i] = tm[i];
  __syncthreads();
  {
    int rk[L];
#pragma unroll
    for (int i = 0; i < L; ++i) rk[i] = 0;
    for (int j = 0; j < 128 * L; ++j) {
      const double o = mins[j];
#pragma unroll
      for (int i = 0; i < L; ++i) rk[i] += (o < tm[i] || (o == tm[i] && j < tid * L + i)) ? 1 : 0;
    }
#pragma unroll
    for (int i = 0; i < L; ++i)
      if (rk[i] == W - 1) tau_s = tm[i];      // ranks are a permutation of 0..128L-1 and W <= 128: exactly one writer
  }
  __syncthreads();
  const double tau = tau_s;
  int ns = 0;
  double above = INFINITY;      // guard: smallest (score, ordinal) beyond tau - the best rejected one when exactly W survive
  int above_c = -1;
  for (int base = w_lo; base < w_hi; base += 32 * UN) {
    double v[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int c = base + u * 32 + lane;
      v[u] = c < w_hi ? value(c) : INFINITY;
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int c = base + u * 32 + lane;
      const bool keep = c < w_hi && v[u] <= tau;
      if (!keep && v[u] < above) {      // ordinals ascend within a thread: strict < keeps the earliest of equals
        above = v[u];
        above_c = c;
      }
      const unsigned m = __ballot_sync(FULL, keep);
      const int pos = ns + __popc(m & ((1u << lane) - 1u));
      if (keep && pos < PB_CAP) {
        sv[warp][pos] = v[u];
        sc[warp][pos] = c;
      }
      ns += __popc(m);
    }
  }
  if (lane == 0) cnt_w[warp] = ns;
  if (guard_eps > 0.0) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      const double xv = __shfl_xor_sync(FULL, above, o);
      const int xc = __shfl_xor_sync(FULL, above_c, o);
      if (xv < above || (xv == above && xc >= 0 && (above_c < 0 || xc < above_c))) {
        above = xv;
        above_c = xc;
      }
    }
    if (lane == 0) {      // mins[] is free again: tau has been published
      mins[warp] = above;
      minc[warp] = above_c;
    }
    if (tid == 0) {
      ov[W] = INFINITY;
      oc[W] = -1;
    }
  }
  __syncthreads();
  const int n0 = cnt_w[0], n1 = cnt_w[1], n2 = cnt_w[2], n3 = cnt_w[3];
  const bool overflow = n0 > PB_CAP || n1 > PB_CAP || n2 > PB_CAP || n3 > PB_CAP;     // block-uniform
  if (overflow) {
    if (warp == 0) {        // mass ties: warp 0 runs the insertion list over the whole frame
      double es[L];
      int ec[L];
#pragma unroll
      for (int l = 0; l < L; ++l) {
        es[l] = INFINITY;
        ec[l] = -1;
      }
      serial_select<L>(value, nc, W, lane, es, ec);
#pragma unroll
      for (int l = 0; l < L; ++l) {
        ov[l * 32 + lane] = es[l];
        oc[l * 32 + lane] = ec[l];
      }
    }
  }
  const int total = overflow ? 0 : n0 + n1 + n2 + n3;
  // survivor g of the concatenation (warp order = ordinal order): rank it against all survivors
  for (int g = tid; g < total; g += 128) {
    int gw = 0, gi = g;
    if (gi >= n0) { gi -= n0; gw = 1; if (gi >= n1) { gi -= n1; gw = 2; if (gi >= n2) { gi -= n2; gw = 3; } } }
    const double x = sv[gw][gi];
    int r = 0, j = 0;
#pragma unroll
    for (int ww = 0; ww < 4; ++ww) {
      const int nw = cnt_w[ww];
      for (int i = 0; i < nw; ++i, ++j) {
        const double y = sv[ww][i];
        r += (y < x || (y == x && j < g)) ? 1 : 0;
      }
    }
    if (r < W) {
      ov[r] = x;
      oc[r] = sc[gw][gi];
    } else if (r == W && guard_eps > 0.0) {
      ov[W] = x;
      oc[W] = sc[gw][gi];
    }
  }
  __syncthreads();
  const int cnt = d.bc[fid];
  const int64_t s0 = d.slot0[fid];
  if (guard_eps > 0.0) {
    // every rank decision of the frame: kept[i] vs kept[i+1], and the last kept path vs the best rejected candidate
    // (rank W among the survivors, else the smallest score beyond tau).  Mass ties (overflow) are flagged outright.
    double gap = INFINITY;
    bool full = overflow;      // the sentence needs the float64 re-decode: no record can describe the decision
    // Which decisions are guarded.  guard_all: every adjacent pair of every frame (the per-frame rank order too).
    // Otherwise those that can change what decode() returns: the kept/rejected boundary of every frame (the SET of
    // kept paths) and, in the sentence's last frame, the pairs that order the first topN paths (the returned list).
    // Two kept paths swapping ranks inside an intermediate frame only renumber later candidates, which the stable
    // sort consults for exactly equal scores alone.
    const bool mine = guard_all || (t == d.sent_T[p] && tid < topN) || (tid == cnt - 1 && cnt == W);
    if (!overflow && tid < cnt && mine) {
      double next = INFINITY;
      int next_c = -1;
      if (tid + 1 < cnt) {
        next = ov[tid + 1];
        next_c = oc[tid + 1];
      } else if (cnt == W) {      // best rejected candidate: rank W among the survivors, else the smallest beyond tau
        next = ov[W];
        next_c = oc[W];
#pragma unroll
        for (int w4 = 0; w4 < 4; ++w4)
          if (minc[w4] >= 0 && (mins[w4] < next || (mins[w4] == next && (next_c < 0 || minc[w4] < next_c)))) {
            next = mins[w4];
            next_c = minc[w4];
          }
      }
      gap = next - ov[tid];
      if (gap < guard_eps) {
        const int slot = atomicAdd(d.guard_n, 1);
        if (slot < d.guard_cap) d.guard_rec[slot] = make_int4(p, t, oc[tid], next_c);
        else full = true;
      }
    }
    const int near = __syncthreads_or(gap < guard_eps);
    const int any_full = __syncthreads_or(full);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) gap = fmin(gap, __shfl_xor_sync(FULL, gap, o));
    __shared__ double gap_w[4];
    if (lane == 0) gap_w[warp] = gap;
    __syncthreads();
    if (tid == 0) {
      const double gmin = overflow ? 0.0 : fmin(fmin(gap_w[0], gap_w[1]), fmin(gap_w[2], gap_w[3]));
      if (gmin < d.guard_gap[p]) d.guard_gap[p] = gmin;      // one CTA per sentence, frames are stream-ordered
      if (near || any_full) d.guard_flag[p] |= (near ? 1 : 0) | (any_full ? 2 : 0);
    }
  }
  if (tid < cnt) {
    const int64_t c = c0 + oc[tid];
    int a = lo, b = hi - 1;
    while (a < b) {
      const int mid = (a + b + 1) >> 1;
      if (d.cand_pos[mid] <= c) a = mid; else b = mid - 1;
    }
    const int par = (int)(d.slot0[d.node_pfid[a]] + (c - d.cand_pos[a]));
    d.slot_score[s0 + tid] = ov[tid];
    d.slot_parent[s0 + tid] = par;
    d.slot_node[s0 + tid] = a;
    d.slot_word[s0 + tid] = d.node_word[a];
    if (DYN) d.slot_cumy[s0 + tid] = d.slot_cumy[par] + d.cand_val[c];
  }
}

// Near-tie guard: the node sequences of the two candidates of every queued record, so the host can re-score both
// paths in float64 (guard_resolve).  A candidate is (node, parent rank); its parent's slots hold the rest of the path.
__global__ void k_guard_paths(BeamDev d, int max_len) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = min(*d.guard_n, d.guard_cap);
  if (i >= 2 * n) return;
  const int4 rec = d.guard_rec[i >> 1];
  const int c = (i & 1) ? rec.w : rec.z;
  int32_t* out = d.guard_paths + (int64_t)i * (max_len + 1);
  if (c < 0) {
    out[0] = 0;
    return;
  }
  const int fid = (int)d.fbase[rec.x] + rec.y;
  const int64_t cg = d.frame_cand_lo[fid] + c;
  int a = d.frame_lo[fid], b = d.frame_hi[fid] - 1;
  while (a < b) {
    const int mid = (a + b + 1) >> 1;
    if (d.cand_pos[mid] <= cg) a = mid; else b = mid - 1;
  }
  const int par = (int)(d.slot0[d.node_pfid[a]] + (cg - d.cand_pos[a]));
  int len = 1;
  for (int s = par; s >= 0; s = d.slot_parent[s]) ++len;
  out[0] = len;
  int q = len - 1;
  out[1 + q] = a;
  --q;
  for (int s = par; s >= 0 && q >= 0; s = d.slot_parent[s], --q) out[1 + q] = d.slot_node[s];
}

// beam_width=None (decoder.py:227-229 not executed): every candidate of the frame becomes a kept path, in the
// reference's enumeration order (node order, then parent rank) - nothing is sorted, nothing is dropped.
template <bool DYN>
__global__ void __launch_bounds__(128)
k_keep_all(BeamDev d, int t, int tstride, int use_lse) {
  pdl_enter();
  const int p = blockIdx.x;
  const int fid = (int)d.fbase[p] + t;
  const int nc = d.frame_ncand[fid];
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const int lo = d.frame_lo[fid], hi = d.frame_hi[fid];
  const int64_t cg = d.frame_cand_lo[fid] + c;
  int a = lo, b = hi - 1;
  while (a < b) {
    const int mid = (a + b + 1) >> 1;
    if (d.cand_pos[mid] <= cg) a = mid; else b = mid - 1;
  }
  const int par = (int)(d.slot0[d.node_pfid[a]] + (cg - d.cand_pos[a]));
  double v = DYN ? d.cand_score[cg] : d.cand_val[cg];
  const int64_t s = d.slot0[fid] + c;
  d.slot_score[s] = v;
  d.slot_parent[s] = par;
  d.slot_node[s] = a;
  d.slot_word[s] = d.node_word[a];
  if (DYN) d.slot_cumy[s] = d.slot_cumy[par] + d.cand_val[cg];
}

// decoder.py:237: walk the back-pointers of the best paths of the last frame.
__global__ void k_backtrace(BeamDev d, int S, int topN, int max_len) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * topN) return;
  const int p = i / topN, k = i % topN;
  const int fid = (int)d.fbase[p] + d.sent_T[p];
  const int cnt = min(topN, d.bc[fid]);
  if (k == 0) d.out_npaths[p] = cnt;
  if (k >= cnt) {
    d.out_len[i] = 0;
    d.out_score[i] = INFINITY;
    return;
  }
  const int slot = (int)d.slot0[fid] + k;
  int depth = 0;
  for (int a = slot; a >= 0; a = d.slot_parent[a]) ++depth;
  d.out_len[i] = depth;
  d.out_score[i] = d.slot_score[slot];
  int q = depth - 1;
  for (int a = slot; a >= 0; a = d.slot_parent[a], --q)
    if (q < max_len) d.out_nodes[(int64_t)i * max_len + q] = d.slot_node[a];
}

// static vocab_select: one warp per LM row, LSE over the sentence's lattice_vocab logits.
// logit j of a job row: the first n_shared columns live in the dense shared block (float32), the rest in the job's own
struct RowLogits {
  const double* p;
  const float* y0;
  int n_shared;
  __device__ __forceinline__ double operator[](int j) const { return j < n_shared ? (double)y0[j] : p[j]; }
};

__global__ void __launch_bounds__(128)
k_job_rows_lse(const SubsetJob* __restrict__ jobs, const double* __restrict__ yv, double* __restrict__ lse_out,
               const float* __restrict__ y0, int ldy0, int n_shared) {
  pdl_enter();
  const SubsetJob job = jobs[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= job.rows) return;
  const RowLogits p{yv + job.out0 + (int64_t)r * job.ncols, y0 + (job.row0 + r) * ldy0, n_shared};
  double mx = -INFINITY;
  for (int j = lane; j < job.ncols; j += 32) mx = fmax(mx, p[j]);
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  double s = 0.0;
  for (int j = lane; j < job.ncols; j += 32) s += exp(p[j] - mx);
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) lse_out[job.row0 + r] = mx + log(s);
}

// DynamicDecoder: the row stepped at frame k is later scored under lattice_vocab[i] for every
// i > k (decoder_dynamic.py:112-148).  Columns are ordered by first appearance, so LSE_i is a
// running log-sum-exp cut at the per-frame boundaries; frame 0 also counts the duplicate entries the
// reference keeps in lattice_vocab[0].
constexpr int DYN_CAP = 1024;   // columns of a sentence's cumulative word list the scan path buffers per warp

// CAP: the scan buffer of a warp in columns (512 when every list of the step fits: 16 KB per CTA instead of 32 KB,
// twice the resident warps of a kernel that is latency-bound)
template <int CAP>
__global__ void __launch_bounds__(128)
k_dyn_prefix_lse(const SubsetJob* __restrict__ jobs, const DynJobInfo* __restrict__ info,
                 const int32_t* __restrict__ vfp, const double* __restrict__ yv, double* __restrict__ dyn_lse,
                 double* dyn_chain, const int32_t* __restrict__ slot_parent, int64_t slot_base, int k, int tstride,
                 const float* __restrict__ y0, int ldy0, int n_shared, int fast_exp) {
  pdl_enter();
  const SubsetJob job = jobs[blockIdx.y];
  const DynJobInfo inf = info[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= job.rows) return;
  const RowLogits p{yv + job.out0 + (int64_t)r * job.ncols, y0 + (job.row0 + r) * ldy0, n_shared};
  const int64_t slot = slot_base + job.row0 + r;
  const int par = slot_parent[slot];
  // Fast path: one maximum over every column the row will ever be scored under, then ONE pass of
  // exp(y - max) with a warp scan; the running sums at the per-frame boundaries are the softmax denominators
  // of lattice_vocab[i] (all terms positive, so a prefix of the sum is as accurate as a sum of its own),
  // and lane i takes the log for frame i.  The per-frame two-pass loop below costs two warp reductions and
  // four float64 transcendentals per future frame and remains for word lists longer than the buffer.
  __shared__ double csum_all[4][CAP];
  const int n_all = vfp[inf.vfp_off + inf.T + 1];
  if (n_all <= CAP && n_all > 0) {
    double* csum = csum_all[threadIdx.x >> 5];
    __shared__ double offs_all[4][32];
    double* offs = offs_all[threadIdx.x >> 5];
    const bool dup = (k == 0);
    double gmax = -INFINITY;
    // pass 1 (coalesced): the row goes into shared memory once; everything after reads it from there
    for (int j = lane; j < n_all; j += 32) {
      const double x = p[j];
      csum[j] = x;
      gmax = fmax(gmax, x);
    }
    if (dup)
      for (int j = inf.nv + lane; j < inf.nv + inf.nd; j += 32) gmax = fmax(gmax, p[j]);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    // tensor-core back end: the logits are float32-accurate, so exp(y - max) is taken with the float32 exp2 (relative
    // error 1e-7, against 1e-6 in the logits themselves) and only summed in float64
    auto ex = [&](double d) { return fast_exp ? (double)exp2f((float)(d * 1.4426950408889634)) : exp(d); };
    double sdup = 0.0;
    if (dup) {
      for (int j = inf.nv + lane; j < inf.nv + inf.nd; j += 32) sdup += ex(p[j] - gmax);
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) sdup += __shfl_xor_sync(0xffffffffu, sdup, o);
    }
    // pass 2: ONE scan per row.  Lane l owns the contiguous segment [l*seg, (l+1)*seg): it turns its logits into
    // running sums of exp(y - max) in place, the 32 segment totals go through a single warp scan, and a boundary's
    // denominator is the in-segment prefix plus the offset of its segment (a scan per 32 columns cost ten shuffles
    // of a double per chunk).  seg is odd so that the lanes' strided 8-byte accesses spread over the banks.
    __syncwarp();
    const int seg = ((n_all + 31) / 32) | 1;
    const int j0 = lane * seg, j1 = min(n_all, j0 + seg);
    double run = 0.0;
    for (int j = j0; j < j1; ++j) {
      run += ex(csum[j] - gmax);
      csum[j] = run;
    }
    double tot = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double up = __shfl_up_sync(0xffffffffu, tot, o);
      if (lane >= o) tot += up;
    }
    offs[lane] = tot - run;            // sum of the segments before this lane's
    __syncwarp();
    for (int i = k + 1 + lane; i <= inf.T; i += 32) {
      const int end = vfp[inf.vfp_off + i + 1];
      const double lse = gmax + log((end > 0 ? csum[end - 1] + offs[(end - 1) / seg] : 0.0) + sdup);
      dyn_lse[slot * tstride + i] = lse;
      dyn_chain[slot * tstride + i] = (par >= 0 ? dyn_chain[(int64_t)par * tstride + i] : 0.0) + lse;
    }
    return;
  }
  double M = -INFINITY, Ssum = 0.0;
  int pos = 0;
  for (int i = k + 1; i <= inf.T; ++i) {
    const int end = vfp[inf.vfp_off + i + 1];
    double mx = -INFINITY;
    for (int j = pos + lane; j < end; j += 32) mx = fmax(mx, p[j]);
    if (i == k + 1 && k == 0)
      for (int j = inf.nv + lane; j < inf.nv + inf.nd; j += 32) mx = fmax(mx, p[j]);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (mx > -INFINITY) {
      double s = 0.0;
      for (int j = pos + lane; j < end; j += 32) s += exp(p[j] - mx);
      if (i == k + 1 && k == 0)
        for (int j = inf.nv + lane; j < inf.nv + inf.nd; j += 32) s += exp(p[j] - mx);
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const double nm = fmax(M, mx);
      Ssum = Ssum * exp(M - nm) + s * exp(mx - nm);
      M = nm;
    }
    pos = end;
    if (lane == 0) {
      const double lse = M + log(Ssum);
      dyn_lse[slot * tstride + i] = lse;
      // sum of this path's LSEs under lattice_vocab[i]; the parent was stepped at an earlier frame
      dyn_chain[slot * tstride + i] = (par >= 0 ? dyn_chain[(int64_t)par * tstride + i] : 0.0) + lse;
    }
  }
}


// ------------------------------------------------------------------------------------------------
// Single sentence, float64 back end: the whole frame loop in ONE cooperative kernel
// ------------------------------------------------------------------------------------------------
// A sentence decoded alone (latency mode; the near-tie guard's re-decode of a flagged sentence) is <= 16 rows per
// LM step: eight launches per frame, each a few microseconds of work behind a launch, a ramp and a drain
// (profiles/r02/launch_summary_single_*.txt: 87 us per frame, 160 launches per 20 kana).  Here one CTA per SM stays
// resident for the whole sentence and the phases of a frame are separated by grid barriers:
//   prune (warp 0 of CTA 0: the lock-step kernel's own selection, prune_sentence)
//   LSTM cell   - warp per hidden unit: the four gate rows of [HM | IM] against the <= 16 gathered input rows,
//                 warp reduce-scatter of the 4 x 16 sums, sigmoid / tanh / cell update on the lanes that end up with
//                 a row's four gates (decoder/model.py:125-139)
//   stage 1     - warp per column of h . PM (model.py:145,162,184)
//   output      - thread per vocabulary column (the word's K <= 512 weight row streamed once, the stage-1 rows
//                 broadcast from shared memory), logits of the CTA's slice in shared memory, one (max, sum exp)
//                 per (row, CTA) (model.py:15-20 softmax as a log-sum-exp)
//   merge+score - the first CTAs merge the per-CTA partials and score the lattice nodes that start at this frame
//                 (score_item, the lock-step kernel's own dot products; Path.append_node, decoder.py:43-49)
// Every quantity that crosses a barrier is read with plain (coherent) loads.
namespace cg = cooperative_groups;

constexpr int SG_THREADS = 256;      // measured with 512 (128-register cap, 4-load batches): 1.94 vs 1.48 ms per 20-kana sentence
constexpr int SG_WARPS = SG_THREADS / 32;
constexpr int SG_MAXM = 16;          // rows per register tile
constexpr int SG_MAXW = 64;          // rows per LM step (beam width): taken SG_MAXM (MT) at a time
constexpr int SG_MAXSTEPS = 128;

struct SingleArgs {
  BeamDev d;
  SegTable seg;
  const float* Wg;        // [4H, Kg]
  const float* bg;        // [4H]
  const float* LM_in;     // [V, Ep]
  const double* P1;       // [Kt, Hp]
  const float* b2;        // [V]
  double* hx;             // [n_slots, Hp]
  double* cx;
  double* T;              // [SG_MAXM, Kt] stage-1 rows of the step
  double2* part;          // [gridDim.x, SG_MAXW] (max, sum exp) of the CTA's vocabulary slice
  int V, H, Hp, Ep, Kg, Kt;
  int W, use_lse, n_steps, sent_T;
  int per;                // vocabulary columns per CTA
  int r0;                 // doubles of shared-memory region 0 (see the kernel)
  int item0[SG_MAXSTEPS + 2];   // work items (lattice nodes starting at frame t): [item0[t], item0[t+1])
};

// N live values per lane -> N/2: lanes with bit OFF set keep the upper half, the others the lower half
template <int N, int OFF>
__device__ __forceinline__ void rs_halve(double* acc, int lane) {
  const bool up = (lane & OFF) != 0;
#pragma unroll
  for (int i = 0; i < N / 2; ++i) {
    const double send = up ? acc[i] : acc[i + N / 2];
    const double keep = up ? acc[i + N / 2] : acc[i];
    acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
  }
}

// Output-layer columns n .. n+NC-1 (one segment) of k_single_f64 for one thread: the words' weight rows stream through
// registers in 64-byte batches (the next batch in flight while this one is multiplied), the stage-1 rows are broadcast
// from shared memory - with NC = 2 every broadcast feeds eight FMAs instead of four (the phase is bound by FMA issue
// and those loads).
template <int MT, int NC>
__device__ __forceinline__ void sg_columns(const SingleArgs& a, const double* As, double* ys, int n, int c_lo, int M) {
  int sgi = 0;
#pragma unroll
  for (int i = 1; i < JLM_MAX_SEGMENTS; ++i)
    if (i < a.seg.n && n >= a.seg.start[i]) sgi = i;
  const int kpad = a.seg.kpad[sgi];
  const float* wrow = a.seg.W[sgi] + (int64_t)(n - a.seg.start[sgi]) * kpad;
  const double* arow = As + a.seg.koff[sgi];
  double acc[NC][MT];
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int m = 0; m < MT; ++m) acc[c][m] = 0.0;
  float4 wn[NC][4];
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int u = 0; u < 4; ++u) wn[c][u] = __ldg(reinterpret_cast<const float4*>(wrow + (size_t)c * kpad + 4 * u));
  for (int k0 = 0; k0 < kpad; k0 += 16) {
    float4 wv[NC][4];
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int u = 0; u < 4; ++u) wv[c][u] = wn[c][u];
    if (k0 + 16 < kpad) {
#pragma unroll
      for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int u = 0; u < 4; ++u)
          wn[c][u] = __ldg(reinterpret_cast<const float4*>(wrow + (size_t)c * kpad + k0 + 16 + 4 * u));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const double2* ap = reinterpret_cast<const double2*>(arow + (size_t)m * a.Kt + k0 + 4 * u);
        const double2 u0 = ap[0], u1 = ap[1];      // the same address in every lane: broadcast
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          double sacc = acc[c][m];
          sacc = fma(u0.x, (double)wv[c][u].x, sacc);
          sacc = fma(u0.y, (double)wv[c][u].y, sacc);
          sacc = fma(u1.x, (double)wv[c][u].z, sacc);
          sacc = fma(u1.y, (double)wv[c][u].w, sacc);
          acc[c][m] = sacc;
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const double bias = (double)__ldg(a.b2 + n + c);
#pragma unroll
    for (int m = 0; m < MT; ++m)
      if (m < M) ys[(size_t)m * a.per + (n + c - c_lo)] = acc[c][m] + bias;
  }
}

// MT: rows the register tiles are written for (>= beam width).  Rows M..MT-1 of the shared-memory operands are zero and
// are multiplied like the others: with `if (m < M)` around every row the compiler keeps each row's four dependent
// float64 FMAs in a basic block of their own and the two warps a scheduler has cannot cover their latency.
template <int MT>
__global__ void __launch_bounds__(SG_THREADS, 1) k_single_f64(const SingleArgs a) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ double sg_sm[];
  const BeamDev& d = a.d;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gwarp = blockIdx.x * SG_WARPS + warp, NW = gridDim.x * SG_WARPS;
  const int Hp = a.Hp, Kt = a.Kt;
  // shared memory: stage-1 rows | logits of this CTA's columns | per-row LSE | parents, words of the step's rows
  // region 0 is, in turn, the gathered gate input [M][Kg], the step's h rows [M][Hp], and the stage-1 rows [M][Kt]
  // followed by the logits of this CTA's columns [M][per]
  double* As = sg_sm;
  double* ys = As + (size_t)SG_MAXM * Kt;
  double* lse_sm = sg_sm + a.r0;                    // [SG_MAXW]
  int* s_par = reinterpret_cast<int*>(lse_sm + SG_MAXW);   // [SG_MAXW]
  int* s_word = s_par + SG_MAXW;                            // [SG_MAXW]

  // ---- prologue: k_init_frame0 (sentence 0) and k_build_items ----
  if (blockIdx.x == 0 && tid == 0) {
    const int fid = (int)d.fbase[0];
    const int64_t s0 = d.slot0[fid];
    const int node = d.frame_lo[fid];
    d.slot_score[s0] = 0.0;
    d.slot_lse[s0] = 0.0;
    d.slot_parent[s0] = -1;
    d.slot_node[s0] = node;
    d.slot_word[s0] = d.node_word[node];
    d.guard_gap[0] = INFINITY;
    d.guard_flag[0] = 0;
    *d.guard_n = 0;
  }
  for (int64_t i = (int64_t)blockIdx.x * SG_THREADS + tid; i < d.n_items; i += (int64_t)gridDim.x * SG_THREADS) {
    const int n = d.start_items[i];
    const int pf = d.node_pfid[n];
    ScoreItem it;
    it.word = d.node_word[n];
    it.rows = d.bc[pf];
    it.node = n;
    it.pad = 0;
    it.ps0 = d.slot0[pf];
    it.cpos = d.cand_pos[n];
    d.items[i] = it;
  }
  grid.sync();

  const int fb = (int)d.fbase[0];
  for (int t = 0; t < a.n_steps; ++t) {
    const int fid = fb + t;
    // Prune: EVERY CTA runs the (deterministic) selection with its warp 0 and writes the same slots - a grid barrier
    // and a phase in which 147 SMs wait for one warp cost more than the redundant scan of a few hundred candidates.
    if (t > 0) {
      if (warp == 0) {
        if (a.W <= 32) prune_sentence<1, false>(d, t, 0, a.W, 0, a.use_lse);
        else prune_sentence<2, false>(d, t, 0, a.W, 0, a.use_lse);
      }
      __syncthreads();
    }
    const int M = (t < a.sent_T) ? d.bc[fid] : 0;     // rows that take an LM step (plan data: the same in every CTA)
    if (M == 0) continue;
    const int64_t row0 = d.slot0[fid];
    if (tid < SG_MAXW) {
      s_par[tid] = tid < M ? d.slot_parent[row0 + tid] : -1;
      s_word[tid] = tid < M ? d.slot_word[row0 + tid] : 0;
    }
    __syncthreads();
    // The step's rows go through every phase MT at a time (beams wider than the register tile: the weights are re-read
    // from L2 per pass).
    for (int mp = 0; mp < M; mp += MT) {
    const int Mp = min(MT, M - mp);
    // gathered gate input [ h[parent] | LM_in[word] ] of the pass's rows -> shared memory (float64), once per CTA
    {
      const int Kg2 = a.Kg >> 1;
      const int total = MT * Kg2;                     // double2 pieces (rows M..MT-1: zeros)
      for (int i0 = tid; i0 < total; i0 += 4 * SG_THREADS) {
        double2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * SG_THREADS;
          v[u] = make_double2(0.0, 0.0);
          if (i < total) {
            const int m = i / Kg2, k = (i - m * Kg2) * 2;
            if (m >= Mp) {
            } else if (k < Hp) {
              const int p = s_par[mp + m];
              if (p >= 0) v[u] = *reinterpret_cast<const double2*>(a.hx + (int64_t)p * Hp + k);
            } else {
              const float2 e = __ldg(reinterpret_cast<const float2*>(a.LM_in + (int64_t)s_word[mp + m] * a.Ep + (k - Hp)));
              v[u] = make_double2((double)e.x, (double)e.y);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * SG_THREADS;
          if (i < total) reinterpret_cast<double2*>(As)[i] = v[u];
        }
      }
    }
    __syncthreads();

    // ---- LSTM cell: warp per hidden unit ----
    for (int j = gwarp; j < Hp; j += NW) {
      if (j >= a.H) {      // padding units of the state rows stay zero (stage 1 multiplies them by zero weights)
        if (lane < Mp) {
          a.hx[(row0 + mp + lane) * Hp + j] = 0.0;
          a.cx[(row0 + mp + lane) * Hp + j] = 0.0;
        }
        continue;
      }
      const float* w0 = a.Wg + (int64_t)j * a.Kg;
      const float* w1 = a.Wg + ((int64_t)a.H + j) * a.Kg;
      const float* w2 = a.Wg + ((int64_t)2 * a.H + j) * a.Kg;
      const float* w3 = a.Wg + ((int64_t)3 * a.H + j) * a.Kg;
      {
        double acc[4 * SG_MAXM];               // index = row * 4 + gate (i, f, o, g); rows >= MT stay zero
#pragma unroll
        for (int i = 0; i < 4 * SG_MAXM; ++i) acc[i] = 0.0;
        for (int k = lane * 4; k < a.Kg; k += 128) {
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(w0 + k));
          const float4 g1 = __ldg(reinterpret_cast<const float4*>(w1 + k));
          const float4 g2 = __ldg(reinterpret_cast<const float4*>(w2 + k));
          const float4 g3 = __ldg(reinterpret_cast<const float4*>(w3 + k));
          const double gw[4][4] = {{(double)g0.x, (double)g0.y, (double)g0.z, (double)g0.w},
                                   {(double)g1.x, (double)g1.y, (double)g1.z, (double)g1.w},
                                   {(double)g2.x, (double)g2.y, (double)g2.z, (double)g2.w},
                                   {(double)g3.x, (double)g3.y, (double)g3.z, (double)g3.w}};
#pragma unroll
          for (int r = 0; r < MT; ++r) {
            {
              const double2* xp = reinterpret_cast<const double2*>(As + (size_t)r * a.Kg + k);
              const double2 xu = xp[0], xv = xp[1];
#pragma unroll
              for (int gte = 0; gte < 4; ++gte) {
                double sacc = acc[r * 4 + gte];
                sacc = fma(xu.x, gw[gte][0], sacc);
                sacc = fma(xu.y, gw[gte][1], sacc);
                sacc = fma(xv.x, gw[gte][2], sacc);
                sacc = fma(xv.y, gw[gte][3], sacc);
                acc[r * 4 + gte] = sacc;
              }
            }
          }
        }
        rs_halve<64, 16>(acc, lane);
        rs_halve<32, 8>(acc, lane);
        rs_halve<16, 4>(acc, lane);
        rs_halve<8, 2>(acc, lane);
        rs_halve<4, 1>(acc, lane);
        // lane l holds the sums with index 2l and 2l+1: row l/2, gates (i, f) on even lanes, (o, g) on odd lanes
        const double o0 = __shfl_xor_sync(0xffffffffu, acc[0], 1), o1 = __shfl_xor_sync(0xffffffffu, acc[1], 1);
        const int m = lane >> 1;
        if ((lane & 1) == 0 && m < Mp) {
          const double pi = acc[0] + (double)a.bg[j], pf = acc[1] + (double)a.bg[a.H + j];
          const double po = o0 + (double)a.bg[2 * a.H + j], pg = o1 + (double)a.bg[3 * a.H + j];
          const double gi = 1.0 / (exp(-pi) + 1.0), gf = 1.0 / (exp(-pf) + 1.0), go = 1.0 / (exp(-po) + 1.0);
          const double gg = tanh(pg);
          const int p = s_par[mp + m];
          const double cp = p >= 0 ? a.cx[(int64_t)p * Hp + j] : 0.0;
          const double c = cp * gf + gg * gi;
          a.cx[(row0 + mp + m) * Hp + j] = c;
          a.hx[(row0 + mp + m) * Hp + j] = tanh(c) * go;
        }
      }
    }
    __syncthreads();      // the next pass overwrites the staged rows
    }
    grid.sync();

    // ---- stage 1: T[m][e] = h[m] . P1[e], warp per column; the pass's h rows go to shared memory first ----
    for (int mp = 0; mp < M; mp += MT) {
    const int Mp = min(MT, M - mp);
    {
      const int Hp2 = Hp >> 1;
      const int total = MT * Hp2, live = Mp * Hp2;
      for (int i0 = tid; i0 < total; i0 += 4 * SG_THREADS) {
        double2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * SG_THREADS;
          v[u] = make_double2(0.0, 0.0);
          if (i < live) v[u] = reinterpret_cast<const double2*>(a.hx + (row0 + mp) * Hp)[i];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * SG_THREADS;
          if (i < total) reinterpret_cast<double2*>(As)[i] = v[u];
        }
      }
    }
    __syncthreads();
    for (int e = gwarp; e < Kt; e += NW) {
      double acc[SG_MAXM];
#pragma unroll
      for (int i = 0; i < SG_MAXM; ++i) acc[i] = 0.0;
      const double* prow = a.P1 + (int64_t)e * Hp;
      for (int k0 = lane * 2; k0 < Hp; k0 += 256) {      // Hp is a multiple of 64: a group of four k-steps may end early
        double2 w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int k = k0 + 64 * u;
          w[u] = k < Hp ? *reinterpret_cast<const double2*>(prow + k) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int k = k0 + 64 * u;
          if (k < Hp) {
#pragma unroll
            for (int m = 0; m < MT; ++m) {
              {
                const double2 hv = *reinterpret_cast<const double2*>(As + (size_t)m * Hp + k);
                acc[m] = fma(hv.x, w[u].x, acc[m]);
                acc[m] = fma(hv.y, w[u].y, acc[m]);
              }
            }
          }
        }
      }
      rs_halve<16, 16>(acc, lane);
      rs_halve<8, 8>(acc, lane);
      rs_halve<4, 4>(acc, lane);
      rs_halve<2, 2>(acc, lane);
      const double v = acc[0] + __shfl_xor_sync(0xffffffffu, acc[0], 1);
      const int m = lane >> 1;
      if ((lane & 1) == 0 && m < Mp) a.T[(int64_t)(mp + m) * Kt + e] = v;
    }
    __syncthreads();
    }
    grid.sync();

    // ---- output layer: (max, sum exp) of this CTA's vocabulary slice for every row ----
    if (a.use_lse) {
      for (int mp = 0; mp < M; mp += MT) {
      const int Mp = min(MT, M - mp);
      {
        const int total = (MT * Kt) >> 1, live = (Mp * Kt) >> 1;
        for (int i0 = tid; i0 < total; i0 += 4 * SG_THREADS) {
          double2 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * SG_THREADS;
            v[u] = make_double2(0.0, 0.0);
            if (i < live) v[u] = reinterpret_cast<const double2*>(a.T + (size_t)mp * Kt)[i];
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * SG_THREADS;
            if (i < total) reinterpret_cast<double2*>(As)[i] = v[u];
          }
        }
      }
      __syncthreads();
      const int c_lo = blockIdx.x * a.per, c_hi = min(a.V, c_lo + a.per);
      // A thread takes two adjacent columns at a time: every broadcast of a stage-1 value then feeds eight FMAs instead
      // of four (measured per 20-kana sentence: one column per thread 1.36 ms, two 1.22 ms; three columns on the first
      // four warps only - one per scheduler - 1.43 ms).  A pair cut by the end of the slice or by a segment boundary
      // goes column by column.
      for (int n = c_lo + 2 * tid; n < c_hi; n += 2 * SG_THREADS) {
        bool pair = n + 1 < c_hi;
#pragma unroll
        for (int i = 1; i < JLM_MAX_SEGMENTS; ++i)
          if (i < a.seg.n && a.seg.start[i] == n + 1) pair = false;
        if (pair) {
          sg_columns<MT, 2>(a, As, ys, n, c_lo, Mp);
        } else {
          sg_columns<MT, 1>(a, As, ys, n, c_lo, Mp);
          if (n + 1 < c_hi) sg_columns<MT, 1>(a, As, ys, n + 1, c_lo, Mp);
        }
      }
      __syncthreads();
      const int nc = max(c_hi - c_lo, 0);
      {
        // (max, sum exp) of the slice per row.  Pass 1: row maxima.  Pass 2: the M x nc exponentials in 2M units of
        // half a row, up to four units interleaved per warp (independent chains of the ~60-instruction float64 exp), so
        // that all eight warps carry about the same number of them; a row's two halves are added in a fixed order.
        __shared__ double mx_s[SG_MAXM];
        __shared__ double us_s[2 * SG_MAXM];
        for (int m = warp; m < Mp; m += SG_WARPS) {
          const double* y = ys + (size_t)m * a.per;
          double mx = -INFINITY;
          for (int i = lane; i < nc; i += 32) mx = fmax(mx, y[i]);
#pragma unroll
          for (int o = 16; o >= 1; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          if (lane == 0) mx_s[m] = mx;
        }
        __syncthreads();
        const int h0 = (nc + 1) >> 1;                 // columns [0, h0) and [h0, nc)
        const double* yu[4];      // 2 x 16 units over eight warps
        double mu[4], su[4] = {0.0, 0.0, 0.0, 0.0};
        int nu[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int u = warp + q * SG_WARPS;
          const bool ok = u < 2 * Mp;
          const int m = ok ? (u >> 1) : 0;
          yu[q] = ys + (size_t)m * a.per + ((u & 1) ? h0 : 0);
          nu[q] = ok ? ((u & 1) ? nc - h0 : h0) : 0;
          mu[q] = mx_s[m];
        }
        for (int i = lane; i < h0; i += 32) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (i < nu[q]) su[q] += exp(yu[q][i] - mu[q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
          for (int o = 16; o >= 1; o >>= 1) su[q] += __shfl_xor_sync(0xffffffffu, su[q], o);
          const int u = warp + q * SG_WARPS;
          if (lane == 0 && u < 2 * Mp) us_s[u] = su[q];
        }
        __syncthreads();
        if (tid < Mp)
          a.part[(size_t)blockIdx.x * SG_MAXW + mp + tid] = make_double2(mx_s[tid], us_s[2 * tid] + us_s[2 * tid + 1]);
      }
      __syncthreads();
      }
      grid.sync();
    }

    // ---- merge the partials, score the nodes that start at this frame ----
    const int it0 = a.item0[t], n_it = a.item0[t + 1] - it0;
    if (blockIdx.x * SG_WARPS < n_it || blockIdx.x == 0) {
      if (a.use_lse) {
        for (int mg = 0; mg < M; mg += 2 * SG_WARPS) {
          // rows warp and warp + 8 of the group together (two independent chains of loads and float64 exponentials per lane)
          const int m0 = mg + warp, m1 = mg + warp + SG_WARPS;
          const bool r0 = m0 < M, r1 = m1 < M;
          const int ma = r0 ? m0 : 0, mb = r1 ? m1 : ma;
          double mxa = -INFINITY, mxb = -INFINITY;
          for (int c = lane; c < (int)gridDim.x; c += 32) {
            mxa = fmax(mxa, a.part[(size_t)c * SG_MAXW + ma].x);
            mxb = fmax(mxb, a.part[(size_t)c * SG_MAXW + mb].x);
          }
#pragma unroll
          for (int o = 16; o >= 1; o >>= 1) {
            mxa = fmax(mxa, __shfl_xor_sync(0xffffffffu, mxa, o));
            mxb = fmax(mxb, __shfl_xor_sync(0xffffffffu, mxb, o));
          }
          double sa = 0.0, sb = 0.0;
          for (int c = lane; c < (int)gridDim.x; c += 32) {
            const double2 pa = a.part[(size_t)c * SG_MAXW + ma], pb = a.part[(size_t)c * SG_MAXW + mb];
            if (pa.x > -INFINITY) sa += pa.y * exp(pa.x - mxa);
            if (pb.x > -INFINITY) sb += pb.y * exp(pb.x - mxb);
          }
#pragma unroll
          for (int o = 16; o >= 1; o >>= 1) {
            sa += __shfl_xor_sync(0xffffffffu, sa, o);
            sb += __shfl_xor_sync(0xffffffffu, sb, o);
          }
          if (lane == 0 && r0) {
            const double l = mxa + log(sa);
            lse_sm[m0] = l;
            if (blockIdx.x == 0) d.slot_lse[row0 + m0] = l;
          }
          if (lane == 0 && r1) {
            const double l = mxb + log(sb);
            lse_sm[m1] = l;
            if (blockIdx.x == 0) d.slot_lse[row0 + m1] = l;
          }
        }
        __syncthreads();
      }
      for (int item = gwarp; item < n_it; item += NW)
        score_item<double, false>(a.seg, a.T, Kt, d, a.b2, it0, item, row0, a.use_lse, 0, a.use_lse ? lse_sm : nullptr);
    }
    grid.sync();
  }
}

// ------------------------------------------------------------------------------------------------
// host: plan
// ------------------------------------------------------------------------------------------------
struct HostPlan {
  std::vector<int32_t> node_span;    // dynamic only
  std::vector<int32_t> node_word, node_pfid, frame_lo, frame_hi, frame_ncand, frame_minpf, bc, sent_T, start_items;
  std::vector<int64_t> cand_pos, frame_cand_lo, slot0, fbase;
  std::vector<int32_t> nstart;       // host only: nodes starting at each frame
  std::vector<int64_t> sent_cand;    // host only: prefix sum of candidates per sorted sentence
  std::vector<SubsetJob> vocab_jobs;
  std::vector<DynJobInfo> dyn_info;
  std::vector<int32_t> vocab_cols, vfp;
};

template <class T>
T* place(Arena& a, const std::vector<T>& v) { return a.take<T>(v.size() ? v.size() : 1); }

template <class F>
void run_threads(int nthreads, F&& work) {
  if (nthreads == 1) {
    work(0);
    return;
  }
  std::vector<std::thread> th;
  for (int k = 1; k < nthreads; ++k) th.emplace_back(work, k);
  work(0);
  for (auto& t : th) t.join();
}

int32_t build_plan(jlm_batch* b, const jlm_lattice_batch* lat, HostPlan& P) {
  jlm_handle* h = b->h;
  const int S = lat->n_sent;
  JLM_REQUIRE(S > 0, "decode: empty batch");
  JLM_REQUIRE(lat->sent_len && lat->frame_ptr_off && lat->frame_ptr && lat->node_start && lat->node_word,
              "decode: null lattice arrays");
  b->S = S;
  b->order.resize(S);
  std::iota(b->order.begin(), b->order.end(), 0);
  std::stable_sort(b->order.begin(), b->order.end(),
                   [&](int a, int c) { return lat->sent_len[a] > lat->sent_len[c]; });
  b->sent_T.resize(S);
  b->fbase.resize(S + 1);
  int64_t F = 0;
  int Tmax = 0;
  for (int p = 0; p < S; ++p) {
    const int T = lat->sent_len[b->order[p]];
    JLM_REQUIRE(T >= 0, "decode: negative sentence length");
    b->sent_T[p] = T;
    b->fbase[p] = F;
    F += T + 1;
    Tmax = std::max(Tmax, T);
  }
  b->fbase[S] = F;
  JLM_REQUIRE(F < (int64_t)1 << 30, "decode: too many frames");
  b->F = F;
  b->Tmax = Tmax;
  b->n_steps = Tmax + 1;
  b->max_len = Tmax + 1;

  int64_t N = 0;
  for (int s = 0; s < S; ++s) {
    const int64_t* fp = lat->frame_ptr + lat->frame_ptr_off[s];
    N = std::max<int64_t>(N, fp[lat->sent_len[s] + 1]);
  }
  JLM_REQUIRE(N < (int64_t)1 << 31, "decode: too many lattice nodes");
  b->N = N;
  P.node_word.assign(lat->node_word, lat->node_word + N);
  P.node_pfid.assign(N, -1);
  P.node_span.assign(b->dynamic ? N : 0, 0);
  P.cand_pos.assign(N, 0);
  P.frame_lo.assign(F, 0);
  P.frame_hi.assign(F, 0);
  P.frame_cand_lo.assign(F, 0);
  P.frame_ncand.assign(F, 0);
  P.frame_minpf.assign(F, 0);
  P.bc.assign(F, 0);
  P.slot0.assign(F, 0);
  P.sent_T = b->sent_T;
  P.fbase = b->fbase;

  // Sentences are independent: the per-node / per-frame tables are filled on host threads over contiguous
  // ranges of sorted positions, candidate offsets first relative to the sentence, then shifted by a
  // prefix sum over the sentences' candidate totals.
  P.nstart.assign(F, 0);
  std::vector<int32_t>& nstart = P.nstart;
  P.sent_cand.assign((size_t)S + 1, 0);
  int nthreads = (int)std::thread::hardware_concurrency();
  if (const char* e = getenv("JLM_HOST_THREADS")) nthreads = atoi(e);
  nthreads = std::max(1, std::min(std::min(nthreads, 8), S / 128));
  std::vector<std::string> errs(nthreads);
  auto plo = [&](int k) { return (int)((int64_t)S * k / nthreads); };
  const int64_t W = b->unlimited ? ((int64_t)1 << 31) : (int64_t)b->W;
  const int V = h->V;
  // DECODE_STATIC_VOCAB: every lattice word must be in the sentence's (sorted) list - the reference's
  // list.index raises ValueError otherwise (decoder.py:179-180)
  const bool check_vocab = b->mode == JLM_DECODE_STATIC_VOCAB && lat->vocab_ptr && lat->vocab_ids;
  run_threads(nthreads, [&](int k) {
    char msg[320];
    for (int p = plo(k); p < plo(k + 1); ++p) {
      const int s = b->order[p];
      const int T = b->sent_T[p];
      const int64_t* fp = lat->frame_ptr + lat->frame_ptr_off[s];
      const int64_t fb = b->fbase[p];
      if (!(fp[1] - fp[0] == 1 && lat->node_start[fp[0]] == -1)) {
        snprintf(msg, sizeof(msg), "decode: frame 0 of sentence %d must hold exactly the <eos> node", s);
        errs[k] = msg;
        return;
      }
      // the list is sorted (decoder.py:141,151); a linear scan covers callers that pass it unsorted
      const int32_t* v0 = check_vocab ? lat->vocab_ids + lat->vocab_ptr[s] : nullptr;
      const int32_t* v1 = check_vocab ? lat->vocab_ids + lat->vocab_ptr[s + 1] : nullptr;
      const bool v_sorted = check_vocab && std::is_sorted(v0, v1);
      int64_t cand = 0;
      for (int t = 0; t <= T; ++t) {
        if (fp[t + 1] < fp[t]) {
          snprintf(msg, sizeof(msg), "decode: frame_ptr not monotone (sentence %d)", s);
          errs[k] = msg;
          return;
        }
        P.frame_lo[fb + t] = (int32_t)fp[t];
        P.frame_hi[fb + t] = (int32_t)fp[t + 1];
        int64_t ncand = 0;
        int minpf = (int)(fb + t);
        P.frame_cand_lo[fb + t] = cand;
        for (int64_t n = fp[t]; n < fp[t + 1]; ++n) {
          const int w = lat->node_word[n];
          if (w < 0 || w >= V) {
            snprintf(msg, sizeof(msg), "decode: word id %d out of range", w);
            errs[k] = msg;
            return;
          }
          if (check_vocab) {
            const bool found = v_sorted ? std::binary_search(v0, v1, (int32_t)w) : std::find(v0, v1, (int32_t)w) != v1;
            if (!found) {
              snprintf(msg, sizeof(msg), "decode: %d is not in list (word of sentence %d missing from its vocabulary list)", w, s);
              errs[k] = msg;
              return;
            }
          }
          if (t == 0) continue;
          const int st = lat->node_start[n];
          if (st < 0 || st >= t) {
            snprintf(msg, sizeof(msg), "decode: node %lld of sentence %d has start %d at frame %d", (long long)n, s, st, t);
            errs[k] = msg;
            return;
          }
          P.node_pfid[n] = (int32_t)(fb + st);
          if (b->dynamic) P.node_span[n] = t - st;
          nstart[fb + st] += 1;
          P.cand_pos[n] = cand + ncand;
          ncand += P.bc[fb + st];
          minpf = std::min(minpf, (int)(fb + st));
        }
        P.frame_minpf[fb + t] = minpf;
        if (b->unlimited && ncand > JLM_MAX_UNPRUNED_PATHS) {
          snprintf(msg, sizeof(msg), "decode: beam_width=None keeps more than %lld paths in frame %d of sentence %d: the "
                   "unpruned search is exponential in the sentence length - pass a finite beam_width",
                   (long long)JLM_MAX_UNPRUNED_PATHS, t, s);
          errs[k] = msg;
          return;
        }
        if (ncand >= (int64_t)1 << 31) {
          errs[k] = "decode: too many candidates in one frame";
          return;
        }
        P.frame_ncand[fb + t] = (int32_t)ncand;
        P.bc[fb + t] = t == 0 ? 1 : (int32_t)std::min<int64_t>(W, ncand);      // unlimited: every candidate is kept
        cand += ncand;
      }
      P.sent_cand[p + 1] = cand;
    }
  });
  for (auto& e : errs) JLM_REQUIRE(e.empty(), "%s", e.c_str());
  if (b->unlimited) {
    // beam_width=None (decoder.py:227-229 skipped): nothing is sorted or pruned, paths multiply frame by frame
    int64_t kept = 0;
    int widest = 1;
    for (int32_t c : P.bc) {
      kept += c;
      widest = std::max(widest, (int)c);
    }
    JLM_REQUIRE(kept <= JLM_MAX_UNPRUNED_PATHS,
                "decode: beam_width=None keeps %lld paths for this batch (limit %lld): the unpruned search is exponential in "
                "the sentence length - pass a finite beam_width", (long long)kept, (long long)JLM_MAX_UNPRUNED_PATHS);
    b->W = widest;
  }
  for (int p = 0; p < S; ++p) P.sent_cand[p + 1] += P.sent_cand[p];
  const int64_t n_cand_total = P.sent_cand[S];
  b->bc = P.bc;
  b->n_cand = n_cand_total;

  // frame-major slot layout + step table
  b->steps.assign(b->n_steps, StepPlan{});
  int64_t slot = 0;
  for (int t = 0; t < b->n_steps; ++t) {
    StepPlan& sp = b->steps[t];
    sp.row0 = slot;
    for (int p = 0; p < S && b->sent_T[p] >= t; ++p) {
      const int64_t fid = b->fbase[p] + t;
      P.slot0[fid] = slot;
      slot += P.bc[fid];
      sp.nact++;
      sp.rows_all += P.bc[fid];
      if (b->sent_T[p] > t) {
        sp.nstep++;
        sp.rows_step += P.bc[fid];
      }
    }
    b->max_rows_step = std::max(b->max_rows_step, sp.rows_step);
  }
  JLM_REQUIRE(slot < (int64_t)1 << 31, "decode: too many beam slots");
  b->n_slots = slot;
  b->slot0 = P.slot0;

  // lattice nodes grouped by the lock-step frame they START at (sentence position, then node order):
  // the work list of the scoring kernel that follows that frame's LM step
  {
    std::vector<int64_t> item_base(F + 1, 0);
    int64_t item = 0;
    for (int t = 0; t < b->n_steps; ++t) {
      StepPlan& sp = b->steps[t];
      sp.item0 = item;
      for (int p = 0; p < sp.nact; ++p) {
        const int64_t fid = b->fbase[p] + t;
        item_base[fid] = item;
        item += nstart[fid];
      }
      JLM_REQUIRE(item - sp.item0 < (int64_t)1 << 31, "decode: too many nodes start at one frame");
      sp.n_items = (int)(item - sp.item0);
    }
    P.start_items.assign(std::max<int64_t>(item, 1), 0);
    // second pass over the sentences: shift the candidate offsets, fill the work list (a sentence's nodes
    // only touch its own frames' counters)
    run_threads(nthreads, [&](int k) {
      for (int p = plo(k); p < plo(k + 1); ++p) {
        const int s = b->order[p];
        const int T = b->sent_T[p];
        const int64_t* fp = lat->frame_ptr + lat->frame_ptr_off[s];
        const int64_t fb = b->fbase[p], base = P.sent_cand[p];
        for (int t = 0; t <= T; ++t) P.frame_cand_lo[fb + t] += base;
        for (int64_t n = fp[1]; n < fp[T + 1]; ++n) {
          P.cand_pos[n] += base;
          P.start_items[item_base[P.node_pfid[n]]++] = (int32_t)n;
        }
      }
    });
  }

  // per-step jobs
  const bool vocab_mode = b->mode != JLM_DECODE_FULL;
  if (vocab_mode) {
    JLM_REQUIRE(lat->vocab_ptr && lat->vocab_ids, "decode: vocabulary lists missing for mode %d", b->mode);
    JLM_REQUIRE(!h->untied, "decode: vocabulary selection with an untied projection raises in the reference "
                            "(decoder/model.py:189); not supported");
    if (b->dynamic) {
      JLM_REQUIRE(lat->vocab_frame_ptr && lat->dup_ptr, "decode: dynamic mode needs vocab_frame_ptr and dup_ptr");
      JLM_REQUIRE(h->n_seg == 1, "decode: DynamicDecoder with a segmented softmax reads permuted logits in the "
                                 "reference (SURVEY quirk 3); not supported");
    }
  }
  std::vector<int64_t> col_ptr(S + 1, 0);  // by caller's sentence index
  if (vocab_mode) {
    for (int s = 0; s < S; ++s) {
      const int64_t nv = lat->vocab_ptr[s + 1] - lat->vocab_ptr[s];
      const int64_t nd = b->dynamic ? lat->dup_ptr[s + 1] - lat->dup_ptr[s] : 0;
      JLM_REQUIRE(nv > 0 && nd >= 0, "decode: empty vocabulary list for sentence %d", s);
      col_ptr[s + 1] = col_ptr[s] + nv + nd;
    }
    P.vocab_cols.resize(col_ptr[S]);
    b->n_vocab_cols = col_ptr[S];
    for (int s = 0; s < S; ++s) {
      const int64_t nv = lat->vocab_ptr[s + 1] - lat->vocab_ptr[s];
      int32_t* dst = &P.vocab_cols[col_ptr[s]];
      for (int64_t j = 0; j < nv; ++j) dst[j] = lat->vocab_ids[lat->vocab_ptr[s] + j];
      if (b->dynamic)
        for (int64_t j = 0; j < lat->dup_ptr[s + 1] - lat->dup_ptr[s]; ++j) dst[nv + j] = lat->dup_ids[lat->dup_ptr[s] + j];
      for (int64_t j = 0; j < col_ptr[s + 1] - col_ptr[s]; ++j)
        JLM_REQUIRE(dst[j] >= 0 && dst[j] < h->V, "decode: vocabulary id %d out of range", dst[j]);
    }
    if (b->dynamic) {
      int64_t tot = 0;
      for (int s = 0; s < S; ++s) tot += lat->sent_len[s] + 2;
      P.vfp.assign(lat->vocab_frame_ptr, lat->vocab_frame_ptr + tot);
    }
    // longest prefix of word ids common to all sentences' lists
    int64_t shared = lat->vocab_ptr[1] - lat->vocab_ptr[0];
    const int32_t* first = lat->vocab_ids + lat->vocab_ptr[0];
    for (int s = 1; s < S && shared > 0; ++s) {
      const int32_t* ids = lat->vocab_ids + lat->vocab_ptr[s];
      const int64_t n = std::min<int64_t>(shared, lat->vocab_ptr[s + 1] - lat->vocab_ptr[s]);
      int64_t k = 0;
      while (k < n && ids[k] == first[k]) ++k;
      shared = k;
    }
    b->n_shared = (S >= 8 && shared >= 32) ? (int)std::min<int64_t>(shared, 1024) : 0;
  }
  int64_t job = 0;
  for (int t = 0; t < b->n_steps; ++t) {
    StepPlan& sp = b->steps[t];
    sp.job0 = job;
    int64_t yv = 0;
    for (int p = 0; p < sp.nstep; ++p) {
      const int64_t fid = b->fbase[p] + t;
      if (vocab_mode) {
        const int s = b->order[p];
        const int nc = (int)(col_ptr[s + 1] - col_ptr[s]);
        SubsetJob vj{P.slot0[fid] - sp.row0, P.bc[fid], col_ptr[s], nc, yv};
        yv += (int64_t)P.bc[fid] * nc;
        P.vocab_jobs.push_back(vj);
        sp.max_vocab_cols = std::max(sp.max_vocab_cols, nc);
        if (b->dynamic) {
          const int64_t nd = lat->dup_ptr[s + 1] - lat->dup_ptr[s];
          DynJobInfo di{lat->frame_ptr_off[s], (int32_t)(nc - nd), (int32_t)nd, b->sent_T[p], 0};
          P.dyn_info.push_back(di);
        }
      }
      ++job;
    }
    sp.yv_elems = yv;
    b->max_yv = std::max(b->max_yv, yv);
  }
  b->n_jobs = job;
  return 0;
}

// device memory layout; called twice (dry run to size, then for real)
void layout(jlm_batch* b, Arena& a, const HostPlan& P) {
  jlm_handle* h = b->h;
  BeamDev& d = b->d;
  d.node_word = place(a, P.node_word);
  d.node_pfid = place(a, P.node_pfid);
  d.node_span = b->dynamic ? place(a, P.node_span) : nullptr;
  d.cand_pos = place(a, P.cand_pos);
  d.frame_lo = place(a, P.frame_lo);
  d.frame_hi = place(a, P.frame_hi);
  d.frame_cand_lo = place(a, P.frame_cand_lo);
  d.frame_ncand = place(a, P.frame_ncand);
  d.frame_minpf = place(a, P.frame_minpf);
  d.bc = place(a, P.bc);
  d.slot0 = place(a, P.slot0);
  d.fbase = place(a, P.fbase);
  d.sent_T = place(a, P.sent_T);
  d.start_items = place(a, P.start_items);
  d.n_items = (int64_t)P.start_items.size();
  d.vocab_jobs = place(a, P.vocab_jobs);
  d.dyn_info = place(a, P.dyn_info);
  d.vocab_cols = place(a, P.vocab_cols);
  d.vfp = place(a, P.vfp);
  const size_t ns = (size_t)b->n_slots;
  d.slot_score = a.take<double>(ns);
  d.slot_lse = a.take<double>(ns);
  d.slot_parent = a.take<int32_t>(ns);
  d.slot_node = a.take<int32_t>(ns);
  d.slot_word = a.take<int32_t>(ns);
  d.slot_cumy = d.dyn_lse = d.dyn_chain = nullptr;
  d.cand_parent = nullptr;
  d.cand_score = nullptr;
  const size_t ncd = (size_t)std::max<int64_t>(b->n_cand, 1);
  if (b->dynamic) {
    d.cand_parent = a.take<int32_t>(ncd);
    d.cand_score = a.take<double>(ncd);
    d.slot_cumy = a.take<double>(ns);
    d.dyn_chain = a.take<double>(ns * (size_t)(b->Tmax + 1));
    d.dyn_lse = a.take<double>(ns * (size_t)(b->Tmax + 1));
  }
  d.cand_val = a.take<double>(ncd);
  d.items = a.take<ScoreItem>((size_t)std::max<int64_t>(d.n_items, 1));
  d.guard_gap = a.take<double>((size_t)b->S);          // guard + n-best blocks are consecutive: one D2H copy
  d.guard_flag = a.take<int32_t>((size_t)b->S);
  d.guard_n = a.take<int32_t>(1);
  d.guard_cap = b->guard_eps > 0.0 ? std::max(64, 4 * b->S) : 1;
  d.guard_rec = a.take<int4>((size_t)d.guard_cap);
  d.guard_paths = a.take<int32_t>((size_t)d.guard_cap * 2 * (b->max_len + 1));
  d.out_score = a.take<double>((size_t)b->S * b->topN);
  d.out_npaths = a.take<int32_t>((size_t)b->S);
  d.out_len = a.take<int32_t>((size_t)b->S * b->topN);
  d.out_nodes = a.take<int32_t>((size_t)b->S * b->topN * b->max_len);
  b->yv = b->max_yv ? a.take<double>((size_t)b->max_yv) : nullptr;
  if (b->backend == JLM_BACKEND_EXACT) {
    const size_t mr = (size_t)std::max(b->max_rows_step, 1);
    b->hx = a.take<double>(ns * h->Hp);
    b->cx = a.take<double>(ns * h->Hp);
    b->A = a.take<double>(mr * h->Kg);
    b->G = a.take<double>(mr * 4 * h->H);
    b->T = h->untied ? nullptr : a.take<double>(mr * h->Kt);
    b->part_tiles = 0;
    for (int i = 0; i < h->n_seg; ++i) b->part_tiles += exact_tiles_n(h->seg[i].end - h->seg[i].start);
    b->part = (b->mode == JLM_DECODE_FULL && b->use_lse) ? a.take<double2>(mr * b->part_tiles) : nullptr;
    b->spart = (b->S == 1 && b->mode == JLM_DECODE_FULL) ? a.take<double2>((size_t)SG_MAXW * std::max(h->sm_count, 1)) : nullptr;
  }
}

template <class T>
void stage(char* host, char* dev_base, T* dev_ptr, const std::vector<T>& v) {
  if (v.empty()) return;
  memcpy(host + (reinterpret_cast<char*>(dev_ptr) - dev_base), v.data(), v.size() * sizeof(T));
}

// Shared tail of an LM step: vocabulary-subset softmax statistics (static / dynamic selection) and
// the logits of the words that can follow each kept path (nodes starting at this frame).
template <typename TT>
int32_t launch_score(jlm_batch* b, int t, const TT* T, int ldt, cudaStream_t st, int defer) {
  const StepPlan& sp = b->steps[t];
  if (sp.n_items <= 0) return 0;
  jlm_handle* h = b->h;
  const int grid = ceil_div(sp.n_items, SC_WARPS);
  const int ul = b->use_lse ? 1 : 0;
  // shared-memory tiles for the stage-1 rows of two start frames per CTA (k_score_nodes): beam x widest segment each,
  // while two of them stay within the default 48 KB (JLM_SCORE_TILES=0: off)
  static const int tiles_on = [] {
    const char* e = getenv("JLM_SCORE_TILES");
    return e ? atoi(e) : 1;
  }();
  int kmax = 0;
  for (int i = 0; i < h->n_seg; ++i) kmax = std::max(kmax, h->seg[i].kpad);
  int tile_elems = b->W * kmax;
  // measured: beam 20 (cfg 4) 61 -> 40 us per frame with the tiles, beam 10 (cfg 2) 24 -> 32 us (staging + two barriers cost
  // more than ten re-read rows save): wide beams only
  // (tensor-core back end only - float32 stage-1 rows: the float64 instantiation exists but no measured or tested shape uses it)
  if (!tiles_on || sizeof(TT) != 4 || b->W <= 12 || h->untied || (size_t)2 * tile_elems * sizeof(TT) > 48 * 1024 ||
      (ldt * sizeof(TT)) % 16 != 0)
    tile_elems = 0;
  const size_t tile_bytes = (size_t)2 * tile_elems * sizeof(TT);
  if (b->dynamic)
    if (tile_elems > 0)
      JLM_CUDA(jlm_launch(k_score_nodes<TT, true, true>, dim3(grid), dim3(SC_WARPS * 32), tile_bytes, st, make_seg_table(h), T, ldt, b->d, h->b2, sp.item0, sp.n_items, sp.row0, ul, 0, t, b->Tmax + 1, tile_elems));
    else
      JLM_CUDA(jlm_launch(k_score_nodes<TT, true, false>, dim3(grid), dim3(SC_WARPS * 32), 0, st, make_seg_table(h), T, ldt, b->d, h->b2, sp.item0, sp.n_items, sp.row0, ul, 0, t, b->Tmax + 1, 0));
  else
    if (tile_elems > 0)
      JLM_CUDA(jlm_launch(k_score_nodes<TT, false, true>, dim3(grid), dim3(SC_WARPS * 32), tile_bytes, st, make_seg_table(h), T, ldt, b->d, h->b2, sp.item0, sp.n_items, sp.row0, ul, defer, t, b->Tmax + 1, tile_elems));
    else
      JLM_CUDA(jlm_launch(k_score_nodes<TT, false, false>, dim3(grid), dim3(SC_WARPS * 32), 0, st, make_seg_table(h), T, ldt, b->d, h->b2, sp.item0, sp.n_items, sp.row0, ul, defer, t, b->Tmax + 1, 0));
  JLM_CUDA(cudaGetLastError());
  b->launches += 1;
  return 0;
}

template <typename TT>
int32_t lm_step_tail(jlm_batch* b, int t, const TT* T, int ldt) {
  jlm_handle* h = b->h;
  cudaStream_t st = h->stream;
  const StepPlan& sp = b->steps[t];
  BeamDev& d = b->d;
  if (b->use_lse && b->mode != JLM_DECODE_FULL) {
    // tensor-core back end: gathered split-fp16 GEMM per sentence; else (or for shapes it does not take) float64
    int32_t on_tc = 2;
    if (b->backend == JLM_BACKEND_TC && sizeof(TT) == sizeof(float)) {
      on_tc = tc_vocab_logits(b, t, b->yv);
      if (on_tc == 1) return 1;
    }
    if (on_tc != 0)
      JLM_TRY(subset_logits<TT>(st, h, T, ldt, d.vocab_jobs + sp.job0, sp.nstep, sp.max_vocab_cols, d.vocab_cols, nullptr,
                                b->yv));
    dim3 grid(ceil_div(b->W, 4), sp.nstep);
    const int ns = (on_tc == 0) ? b->n_shared : 0;      // the shared block exists only when the tensor-core kernel ran
    if (b->dynamic)
      if (sp.max_vocab_cols <= DYN_CAP / 2)
        JLM_CUDA(jlm_launch(k_dyn_prefix_lse<DYN_CAP / 2>, dim3(grid), dim3(128), 0, st, d.vocab_jobs + sp.job0, d.dyn_info + sp.job0, d.vfp, b->yv, d.dyn_lse, d.dyn_chain, d.slot_parent, sp.row0, t, b->Tmax + 1, b->y0, b->ldy0, ns, on_tc == 0 ? 1 : 0));
      else
        JLM_CUDA(jlm_launch(k_dyn_prefix_lse<DYN_CAP>, dim3(grid), dim3(128), 0, st, d.vocab_jobs + sp.job0, d.dyn_info + sp.job0, d.vfp, b->yv, d.dyn_lse, d.dyn_chain, d.slot_parent, sp.row0, t, b->Tmax + 1, b->y0, b->ldy0, ns, on_tc == 0 ? 1 : 0));
    else
      JLM_CUDA(jlm_launch(k_job_rows_lse, dim3(grid), dim3(128), 0, st, d.vocab_jobs + sp.job0, b->yv, d.slot_lse + sp.row0, b->y0, b->ldy0, ns));
    JLM_CUDA(cudaGetLastError());
    b->launches += 2;
  }
  return launch_score<TT>(b, t, T, ldt, st, 0);
}

int32_t exact_lm_step(jlm_batch* b, int t) {
  jlm_handle* h = b->h;
  cudaStream_t st = h->stream;
  const StepPlan& sp = b->steps[t];
  const int M = sp.rows_step;
  if (M == 0) return 0;
  BeamDev& d = b->d;
  const int32_t* parent = d.slot_parent + sp.row0;
  const int32_t* word = d.slot_word + sp.row0;
  double* hrow = b->hx + sp.row0 * h->Hp;
  double* crow = b->cx + sp.row0 * h->Hp;
  if (b->timers) cudaEventRecord(b->events[3 * t + 1], st);
  JLM_TRY(exact_gather_gate_input(st, h, b->hx, parent, word, M, b->A));
  JLM_TRY(exact_gemm_f32w(st, b->A, h->Kg, h->Wg, h->Kg, h->bg, b->G, 4 * h->H, M, 4 * h->H, h->Kg, nullptr, 0, 0));
  JLM_TRY(exact_lstm_pointwise(st, h, b->G, b->cx, parent, M, hrow, crow));
  b->launches += 3;
  if (b->timers) cudaEventRecord(b->events[3 * t + 2], st);
  const double* T = hrow;
  int ldt = h->Hp;
  if (!h->untied) {
    JLM_TRY(exact_gemm_f64w(st, hrow, h->Hp, h->P1, h->Hp, b->T, h->Kt, M, h->Kt, h->Hp));
    b->launches += 1;
    T = b->T;
    ldt = h->Kt;
  }
  if (b->use_lse && b->mode == JLM_DECODE_FULL) {
    int tile0 = 0;
    for (int i = 0; i < h->n_seg; ++i) {
      const SegDev& s = h->seg[i];
      const int Vi = s.end - s.start;
      if (exact_use_q8(h, s, M))
        JLM_TRY(exact_gemm_q8w(st, T + s.koff, ldt, s.Wq, s.kpad, s.cb, h->b2 + s.start, nullptr, 0, M, Vi, s.kpad,
                               b->part, b->part_tiles, tile0));
      else
        JLM_TRY(exact_gemm_f32w(st, T + s.koff, ldt, s.W, s.kpad, h->b2 + s.start, nullptr, 0, M, Vi, s.kpad, b->part,
                                b->part_tiles, tile0));
      tile0 += exact_tiles_n(Vi);
      b->launches += 1;
    }
    JLM_TRY(exact_lse_merge(st, b->part, b->part_tiles, b->part_tiles, M, d.slot_lse + sp.row0, 0));
    b->launches += 1;
  }
  return lm_step_tail<double>(b, t, T, ldt);
}

// Runs the whole batch through k_single_f64 when it is one sentence the kernel's shapes cover; *done says whether it
// did (otherwise the caller goes through the per-frame launches).  JLM_SINGLE=0 switches the kernel off.
int32_t single_try_run(jlm_batch* b, bool* done) {
  *done = false;
  jlm_handle* h = b->h;
  static const int enabled = [] {
    const char* e = getenv("JLM_SINGLE");
    return e ? atoi(e) : 1;
  }();
  if (!enabled || b->backend != JLM_BACKEND_EXACT || b->S != 1 || b->mode != JLM_DECODE_FULL || b->dynamic || b->unlimited ||
      b->W > SG_MAXW || b->timers || h->untied || !h->P1 || !b->T || !b->spart || b->n_steps > SG_MAXSTEPS)
    return 0;
  for (int i = 0; i < h->n_seg; ++i)
    if (h->seg[i].kpad % 32 != 0 || exact_use_q8(h, h->seg[i], b->W)) return 0;
  if (h->Hp % 64 != 0 || h->Ep % 4 != 0 || h->Kt % 4 != 0 || h->Kg % 4 != 0) return 0;
  static int blocks_per_sm = -1;
  int dev_smem = 0;
  JLM_CUDA(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
  const int G = h->sm_count;
  const int per = ceil_div(h->V, G);
  const size_t r0 = (size_t)SG_MAXM * std::max<size_t>(std::max(h->Kg, h->Hp), (size_t)h->Kt + per);
  const size_t smem = (r0 + SG_MAXW) * sizeof(double) + 2 * SG_MAXW * sizeof(int);
  if (smem + 20 * 1024 > (size_t)dev_smem) return 0;     // + the prune phase's static buffers
  const void* kern = b->W <= 4 ? reinterpret_cast<const void*>(k_single_f64<4>)
                     : b->W <= 8 ? reinterpret_cast<const void*>(k_single_f64<8>)
                     : b->W <= 10 ? reinterpret_cast<const void*>(k_single_f64<10>)
                     : b->W <= 12 ? reinterpret_cast<const void*>(k_single_f64<12>)
                                  : reinterpret_cast<const void*>(k_single_f64<16>);
  static size_t smem_set = 0;
  if (smem > smem_set) {
    JLM_CUDA(cudaFuncSetAttribute(k_single_f64<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    JLM_CUDA(cudaFuncSetAttribute(k_single_f64<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    JLM_CUDA(cudaFuncSetAttribute(k_single_f64<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    JLM_CUDA(cudaFuncSetAttribute(k_single_f64<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    JLM_CUDA(cudaFuncSetAttribute(k_single_f64<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
    blocks_per_sm = -1;
  }
  if (blocks_per_sm < 0) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_single_f64<16>, SG_THREADS, smem) != cudaSuccess) {
      cudaGetLastError();
      nb = 0;
    }
    blocks_per_sm = nb;
  }
  if (blocks_per_sm < 1) return 0;      // a cooperative grid must be co-resident: one CTA per SM
  SingleArgs sa{};
  sa.d = b->d;
  sa.seg = make_seg_table(h);
  sa.Wg = h->Wg;
  sa.bg = h->bg;
  sa.LM_in = h->LM_in;
  sa.P1 = h->P1;
  sa.b2 = h->b2;
  sa.hx = b->hx;
  sa.cx = b->cx;
  sa.T = b->T;
  sa.part = b->spart;
  sa.V = h->V;
  sa.H = h->H;
  sa.Hp = h->Hp;
  sa.Ep = h->Ep;
  sa.Kg = h->Kg;
  sa.Kt = h->Kt;
  sa.W = b->W;
  sa.use_lse = b->use_lse ? 1 : 0;
  sa.n_steps = b->n_steps;
  sa.sent_T = b->sent_T[0];
  sa.per = per;
  sa.r0 = (int)r0;
  for (int t = 0; t < b->n_steps; ++t) sa.item0[t] = (int)b->steps[t].item0;
  sa.item0[b->n_steps] = (int)b->d.n_items;
  void* args[] = {&sa};
  static bool coop_failed = false;      // a device / context that refuses cooperative launches: per-frame launches from then on
  if (coop_failed) return 0;
  if (cudaLaunchCooperativeKernel(kern, dim3(G), dim3(SG_THREADS), args, smem, h->stream) != cudaSuccess) {
    cudaGetLastError();
    coop_failed = true;
    return 0;
  }
  b->launches += 1;
  *done = true;
  return 0;
}

bool score_overlap_enabled() {
  static const int v = [] {
    // Measured on B200 (cfg 2, 1024 sentences): with the scoring kernel co-resident the output GEMM slows from
    // 0.394 to 0.418 ms per launch - more than the 0.024 ms the overlap hides - so it is off unless asked for.
    const char* e = getenv("JLM_OVERLAP_SCORE");
    return e ? atoi(e) : 0;
  }();
  return v != 0;
}

template <bool DYN>
int32_t launch_prune(jlm_batch* b, int t) {
  const StepPlan& sp = b->steps[t];
  if (sp.nact == 0) return 0;
  const int grid = ceil_div(sp.nact, 4);
  const int L = ceil_div(b->W, 32);
  const int ul = b->use_lse ? 1 : 0;
  cudaStream_t st = b->h->stream;
  if (b->unlimited) {
    JLM_CUDA(jlm_launch(k_keep_all<DYN>, dim3(dim3(sp.nact, ceil_div(b->W, 128))), dim3(128), 0, st, b->d, t, b->Tmax + 1, ul));
    JLM_CUDA(cudaGetLastError());
    b->launches += 1;
    return 0;
  }
  static const int block_mode = [] {
    const char* e = getenv("JLM_PRUNE_BLOCK");
    return e ? atoi(e) : 1;
  }();
  if (block_mode || b->guard_eps > 0.0) {      // one CTA per sentence (the near-tie guard lives in this kernel)
    if (L <= 1)
      JLM_CUDA(jlm_launch(k_prune_block<1, DYN>, dim3(sp.nact), dim3(128), 0, st, b->d, t, b->W, b->Tmax + 1, ul, b->guard_eps, b->guard_all ? 1 : 0, b->topN));
    else if (L <= 2)
      JLM_CUDA(jlm_launch(k_prune_block<2, DYN>, dim3(sp.nact), dim3(128), 0, st, b->d, t, b->W, b->Tmax + 1, ul, b->guard_eps, b->guard_all ? 1 : 0, b->topN));
    else
      JLM_CUDA(jlm_launch(k_prune_block<4, DYN>, dim3(sp.nact), dim3(128), 0, st, b->d, t, b->W, b->Tmax + 1, ul, b->guard_eps, b->guard_all ? 1 : 0, b->topN));
  } else if (L <= 1)
    JLM_CUDA(jlm_launch(k_prune<1, DYN>, dim3(grid), dim3(128), 0, st, b->d, t, sp.nact, b->W, b->Tmax + 1, ul));
  else if (L <= 2)
    JLM_CUDA(jlm_launch(k_prune<2, DYN>, dim3(grid), dim3(128), 0, st, b->d, t, sp.nact, b->W, b->Tmax + 1, ul));
  else
    JLM_CUDA(jlm_launch(k_prune<4, DYN>, dim3(grid), dim3(128), 0, st, b->d, t, sp.nact, b->W, b->Tmax + 1, ul));
  JLM_CUDA(cudaGetLastError());
  b->launches += 1;
  return 0;
}

}  // namespace

// Host copy of a lattice batch: jlm_batch_fetch may have to re-decode a few sentences (near-tie guard) long
// after the caller's arrays are gone.  ~8 bytes per node, a fraction of a millisecond per 1024 sentences.
struct GuardLattice {
  std::vector<int32_t> sent_len, node_start, node_word, vocab_ids, vocab_frame_ptr, dup_ids;
  std::vector<int64_t> frame_ptr_off, frame_ptr, vocab_ptr, dup_ptr;
  // the sub-batch of the flagged sentences (arrays a jlm_lattice_batch view points into)
  std::vector<int32_t> f_sent_len, f_node_start, f_node_word, f_vocab_ids, f_vocab_frame_ptr, f_dup_ids;
  std::vector<int64_t> f_frame_ptr_off, f_frame_ptr, f_vocab_ptr, f_dup_ptr;
  std::vector<int64_t> f_node_delta;      // original node index - sub-batch node index, per flagged sentence
  // its n-best lists
  std::vector<double> r_score;
  std::vector<int32_t> r_npaths, r_len, r_nodes;
  int r_max_len = 0;

  void copy_from(const jlm_lattice_batch* lat, int mode) {
    const int S = lat->n_sent;
    sent_len.assign(lat->sent_len, lat->sent_len + S);
    frame_ptr_off.assign(lat->frame_ptr_off, lat->frame_ptr_off + S);
    int64_t fp_end = 0, n_nodes = 0;
    for (int s = 0; s < S; ++s) {
      fp_end = std::max(fp_end, lat->frame_ptr_off[s] + lat->sent_len[s] + 2);
      n_nodes = std::max(n_nodes, lat->frame_ptr[lat->frame_ptr_off[s] + lat->sent_len[s] + 1]);
    }
    frame_ptr.assign(lat->frame_ptr, lat->frame_ptr + fp_end);
    node_start.assign(lat->node_start, lat->node_start + n_nodes);
    node_word.assign(lat->node_word, lat->node_word + n_nodes);
    vocab_ptr.clear();
    dup_ptr.clear();
    if (mode != JLM_DECODE_FULL) {
      vocab_ptr.assign(lat->vocab_ptr, lat->vocab_ptr + S + 1);
      vocab_ids.assign(lat->vocab_ids, lat->vocab_ids + lat->vocab_ptr[S]);
      if (mode == JLM_DECODE_DYNAMIC) {
        vocab_frame_ptr.assign(lat->vocab_frame_ptr, lat->vocab_frame_ptr + fp_end);
        dup_ptr.assign(lat->dup_ptr, lat->dup_ptr + S + 1);
        dup_ids.assign(lat->dup_ids, lat->dup_ids + lat->dup_ptr[S]);
      }
    }
  }

  // view over the sentences `which` (caller's indices), node indices renumbered from 0
  jlm_lattice_batch subset(const std::vector<int>& which, int mode) {
    f_sent_len.clear(); f_node_start.clear(); f_node_word.clear(); f_vocab_ids.clear(); f_vocab_frame_ptr.clear();
    f_dup_ids.clear(); f_frame_ptr_off.clear(); f_frame_ptr.clear(); f_node_delta.clear();
    f_vocab_ptr.assign(1, 0);
    f_dup_ptr.assign(1, 0);
    for (int s : which) {
      const int T = sent_len[s];
      const int64_t* fp = &frame_ptr[frame_ptr_off[s]];
      const int64_t n0 = fp[0], n1 = fp[T + 1];
      const int64_t base = (int64_t)f_node_start.size();
      f_sent_len.push_back(T);
      f_frame_ptr_off.push_back((int64_t)f_frame_ptr.size());
      f_node_delta.push_back(n0 - base);
      for (int t = 0; t <= T + 1; ++t) f_frame_ptr.push_back(fp[t] - n0 + base);
      f_node_start.insert(f_node_start.end(), node_start.begin() + n0, node_start.begin() + n1);
      f_node_word.insert(f_node_word.end(), node_word.begin() + n0, node_word.begin() + n1);
      if (mode != JLM_DECODE_FULL) {
        f_vocab_ids.insert(f_vocab_ids.end(), vocab_ids.begin() + vocab_ptr[s], vocab_ids.begin() + vocab_ptr[s + 1]);
        f_vocab_ptr.push_back((int64_t)f_vocab_ids.size());
        if (mode == JLM_DECODE_DYNAMIC) {
          const int32_t* vf = &vocab_frame_ptr[frame_ptr_off[s]];
          f_vocab_frame_ptr.insert(f_vocab_frame_ptr.end(), vf, vf + T + 2);
          f_dup_ids.insert(f_dup_ids.end(), dup_ids.begin() + dup_ptr[s], dup_ids.begin() + dup_ptr[s + 1]);
          f_dup_ptr.push_back((int64_t)f_dup_ids.size());
        }
      }
    }
    jlm_lattice_batch v{};
    v.n_sent = (int32_t)which.size();
    v.sent_len = f_sent_len.data();
    v.frame_ptr_off = f_frame_ptr_off.data();
    v.frame_ptr = f_frame_ptr.data();
    v.node_start = f_node_start.data();
    v.node_word = f_node_word.data();
    if (mode != JLM_DECODE_FULL) {
      v.vocab_ptr = f_vocab_ptr.data();
      v.vocab_ids = f_vocab_ids.data();
      if (mode == JLM_DECODE_DYNAMIC) {
        v.vocab_frame_ptr = f_vocab_frame_ptr.data();
        v.dup_ptr = f_dup_ptr.data();
        v.dup_ids = f_dup_ids.data();
      }
    }
    return v;
  }
};

static void guard_drop_rerun(jlm_batch* b) {
  if (b->rerun) jlm_batch_destroy(b->rerun);
  b->rerun = nullptr;
  b->rerun_index.clear();
  b->n_flagged = b->n_pairs = b->n_rerun = 0;
}

// The guard's float64 work runs beside the NEXT batch's kernels (a streaming caller has already enqueued them on the
// main stream): its own stream, at the highest priority, so that its short dependent kernels are picked ahead of the
// next long tensor-core kernel whenever SMs free up instead of queueing behind the whole batch.
static int32_t guard_stream_create(jlm_handle* h) {
  if (h->guard_stream) return 0;
  int lo = 0, hi = 0;
  JLM_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));      // hi is the numerically lowest = highest priority
  JLM_CUDA(cudaStreamCreateWithPriority(&h->guard_stream, cudaStreamNonBlocking, hi));
  return 0;
}

// Tier 1 of the near-tie guard: re-score the two paths of every queued rank decision in float64.
// A path's score depends on its word sequence only (Path.append_node, decoder.py:43-49: sum of -log softmax along
// the path), so all paths of all sentences go into ONE trie of word prefixes; each trie node is a state of the
// float64 LM pool (jlm_pool: exact back end kernels), stepped level by level, with one log-sum-exp GEMM over all
// nodes at the end.  A decision is confirmed when the float64 scores order the pair the way the tensor-core scores
// did (equal scores: by enumeration ordinal, the stable sort's rule); a contradicted decision sends its sentence to
// tier 2, the float64 re-decode.  Runs on its own stream, beside whatever the main stream holds.
static int32_t guard_verify_pairs(jlm_batch* b, int rec_lo, int rec_hi, const int4* rec, const int32_t* paths, std::vector<char>& need_full) {
  const int n_rec = rec_hi - rec_lo;      // this call's slice of the record queue
  jlm_handle* h = b->h;
  const GuardLattice& G = *b->guard_lat;
  const int L = b->max_len + 1;
  struct TNode { int parent, word, depth; int64_t slot; };
  std::vector<TNode> trie;
  std::unordered_map<uint64_t, int> index;          // (parent + 1) << 32 | word -> trie node
  auto child = [&](int parent, int word) {
    const uint64_t key = ((uint64_t)(uint32_t)(parent + 1) << 32) | (uint32_t)word;
    auto it = index.find(key);
    if (it != index.end()) return it->second;
    const int id = (int)trie.size();
    trie.push_back(TNode{parent, word, parent < 0 ? 0 : trie[parent].depth + 1, -1});
    index.emplace(key, id);
    return id;
  };
  struct Query { int node, word; };                 // -log p(word | state of trie node)
  std::vector<Query> queries;
  std::unordered_map<uint64_t, int> qindex;
  // Vocabulary-selection modes: the softmax of a state runs over a word list that depends on the sentence (static:
  // its lattice_vocab, decoder.py:137-151) and, for DynamicDecoder, on the frame the candidate ends at (every
  // transition of a path is re-normalised over lattice_vocab[t], decoder_dynamic.py:150-175) - one log-sum-exp
  // request per (state, sentence, frame).
  const bool subset = b->mode != JLM_DECODE_FULL;
  struct LseReq { int node, sent, frame; };
  std::vector<LseReq> reqs;
  std::unordered_map<uint64_t, int> rindex;
  auto req_key = [&](int node, int p, int t) { return ((uint64_t)(uint32_t)node << 40) ^ ((uint64_t)(uint32_t)p << 12) ^ (uint64_t)(uint32_t)t; };
  struct PathRef { std::vector<int> q; std::vector<int> lse; };
  std::vector<PathRef> refs((size_t)2 * n_rec);      // indexed by record - rec_lo
  std::vector<char> need_lse;                       // per trie node
  for (int r = rec_lo; r < rec_hi; ++r) {
    if (need_full[rec[r].x]) continue;
    const int32_t* pa = paths + ((int64_t)2 * r) * L;
    const int32_t* pb = pa + L;
    if (pa[0] <= 0 || pb[0] <= 0 || pa[0] > b->max_len || pb[0] > b->max_len) {   // no second candidate recorded
      need_full[rec[r].x] = 1;
      continue;
    }
    // The two paths share their first `common` words: those transitions are the SAME float64 terms in both scores
    // and are not evaluated; at the first differing word both still read the same state, so its log-normaliser
    // cancels as well.  Only states that contain a differing word need their full-vocabulary log-sum-exp.
    int common = 0;
    while (common < pa[0] && common < pb[0] && G.node_word[pa[1 + common]] == G.node_word[pb[1 + common]]) ++common;
    for (int which = 0; which < 2; ++which) {
      const int32_t* pth = which ? pb : pa;
      const int len = pth[0];
      int node = -1;
      PathRef& pr = refs[(size_t)2 * (r - rec_lo) + which];
      for (int j = 0; j < len; ++j) {
        const int w = G.node_word[pth[1 + j]];
        if (j >= common && j > 0) {                 // transition into word j from the state of words 0..j-1
          const uint64_t qk = ((uint64_t)(uint32_t)node << 32) | (uint32_t)w;
          auto it = qindex.find(qk);
          if (it == qindex.end()) {
            it = qindex.emplace(qk, (int)queries.size()).first;
            queries.push_back(Query{node, w});
          }
          pr.q.push_back(it->second);
          int req = -1;
          if (j > common && b->use_lse) {
            if (subset) {
              const uint64_t rk = req_key(node, rec[r].x, rec[r].y);
              auto rt = rindex.find(rk);
              if (rt == rindex.end()) {
                rt = rindex.emplace(rk, (int)reqs.size()).first;
                reqs.push_back(LseReq{node, rec[r].x, rec[r].y});
              }
              req = rt->second;
            } else {
              if ((int)need_lse.size() < (int)trie.size()) need_lse.resize(trie.size(), 0);
              need_lse[node] = 1;
            }
          }
          pr.lse.push_back(req);
        }
        if (j + 1 < len) node = child(node, w);     // the last word's state is never consumed
      }
    }
  }
  need_lse.resize(trie.size(), 0);
  if (queries.empty()) {
    // every pair is a pair of identical word sequences (duplicate lattice nodes): equal float64 scores, and the
    // record's order (rank i before rank i+1 of a stable sort) already is the enumeration order
    for (int r = rec_lo; r < rec_hi; ++r)
      if (!need_full[rec[r].x]) b->n_pairs += 1;
    return 0;
  }
  constexpr int64_t POOL_CAP = 16384;
  if ((int64_t)trie.size() > POOL_CAP) {            // pathological batch: let tier 2 handle every flagged sentence
    for (int r = rec_lo; r < rec_hi; ++r) need_full[rec[r].x] = 1;
    return 0;
  }
  JLM_TRY(guard_stream_create(h));
  if (!h->guard_pool) JLM_TRY(jlm_pool_create(h, POOL_CAP, &h->guard_pool));
  // the pool works on the handle's stream: lend it the guard stream for the duration of the check
  cudaStream_t main_stream = h->stream;
  h->stream = h->guard_stream;
  int32_t rc = jlm_pool_reset(h->guard_pool);
  int max_depth = 0;
  for (auto& t : trie) max_depth = std::max(max_depth, t.depth);
  std::vector<std::vector<int>> by_depth(max_depth + 1);
  for (int i = 0; i < (int)trie.size(); ++i) by_depth[trie[i].depth].push_back(i);
  std::vector<int32_t> src, idx;
  for (int dpt = 0; dpt <= max_depth && !rc; ++dpt) {
    src.clear();
    idx.clear();
    for (int i : by_depth[dpt]) {
      src.push_back(trie[i].parent < 0 ? -1 : (int32_t)trie[trie[i].parent].slot);
      idx.push_back(trie[i].word);
    }
    int64_t first = 0;
    rc = pool_step_rows(h->guard_pool, (int32_t)src.size(), src.data(), idx.data(), &first, false);
    for (size_t k = 0; k < by_depth[dpt].size(); ++k) trie[by_depth[dpt][k]].slot = first + (int64_t)k;
  }
  std::vector<double> req_lse(reqs.size(), 0.0);
  if (!subset) {
    std::vector<int32_t> lse_slots;
    for (size_t i = 0; i < trie.size(); ++i)
      if (need_lse[i]) lse_slots.push_back((int32_t)trie[i].slot);
    if (!rc && !lse_slots.empty()) rc = pool_lse_slots(h->guard_pool, lse_slots.data(), (int32_t)lse_slots.size());
    b->n_lse_rows += (int)lse_slots.size();
  } else if (!reqs.empty()) {
    std::vector<int32_t> rslots(reqs.size()), rcols;
    std::vector<int64_t> rptr(reqs.size() + 1, 0);
    for (size_t i = 0; i < reqs.size(); ++i) {
      const int s = b->order[reqs[i].sent];
      const int32_t* ids = &G.vocab_ids[G.vocab_ptr[s]];
      int64_t n_list = G.vocab_ptr[s + 1] - G.vocab_ptr[s];
      if (b->dynamic) n_list = G.vocab_frame_ptr[G.frame_ptr_off[s] + reqs[i].frame + 1];   // lattice_vocab[frame]
      rcols.insert(rcols.end(), ids, ids + n_list);
      if (b->dynamic && trie[reqs[i].node].depth == 0)      // the <eos> row also counts lattice_vocab[0]'s duplicates (quirk 4)
        rcols.insert(rcols.end(), G.dup_ids.begin() + G.dup_ptr[s], G.dup_ids.begin() + G.dup_ptr[s + 1]);
      rslots[i] = (int32_t)trie[reqs[i].node].slot;
      rptr[i + 1] = (int64_t)rcols.size();
    }
    if (!rc) rc = pool_lse_subsets(h->guard_pool, (int32_t)reqs.size(), rslots.data(), rptr.data(), rcols.data(), req_lse.data());
    b->n_lse_rows += (int)reqs.size();
  }
  std::vector<int32_t> qs(queries.size()), qw(queries.size());
  std::vector<double> nll(queries.size());
  for (size_t i = 0; i < queries.size(); ++i) {
    qs[i] = (int32_t)trie[queries[i].node].slot;
    qw[i] = queries[i].word;
  }
  if (!rc) rc = jlm_pool_nll(h->guard_pool, (int32_t)queries.size(), qs.data(), qw.data(), nll.data());
  h->stream = main_stream;
  if (rc) return rc;
  for (int r = rec_lo; r < rec_hi; ++r) {
    if (need_full[rec[r].x]) continue;
    double sc[2];
    for (int which = 0; which < 2; ++which) {
      double v = 0.0;                               // Path.__init__: neg_log_prob = 0, then += per node (decoder.py:34,49)
      const PathRef& pr = refs[(size_t)2 * (r - rec_lo) + which];
      for (size_t k = 0; k < pr.q.size(); ++k) v += nll[pr.q[k]] + (pr.lse[k] >= 0 ? req_lse[pr.lse[k]] : 0.0);
      sc[which] = v;
    }
    b->n_pairs += 1;
    // the tensor-core ranking put candidate z before candidate w; the stable sort agrees iff score_z < score_w, or
    // the scores are equal and z comes first in enumeration order
    const bool ok = sc[0] < sc[1] || (sc[0] == sc[1] && rec[r].z < rec[r].w);
    if (!ok) need_full[rec[r].x] = 1;
    static const bool dbg = getenv("JLM_DEBUG_TIMING") != nullptr;
    if (dbg && !ok)
      fprintf(stderr, "[jlm] guard: contradicted decision: sentence %d (T=%d) frame %d, float64 suffix scores %.9g vs %.9g\n",
              b->order[rec[r].x], b->sent_T[rec[r].x], rec[r].y, sc[0], sc[1]);
  }
  return 0;
}

// Near-tie guard, host side (once per run; later fetches reuse the result).  Tier 1 re-scores the queued pairs
// (full-softmax decoding; the vocabulary-selection modes normalise over per-frame word lists the pool does not
// know); tier 2 re-decodes, on the float64 back end, the sentences tier 1 could not confirm.
static int32_t guard_resolve(jlm_batch* b, const char* host, const char* src) {
  if (b->guard_eps <= 0.0 || !b->guard_lat || !b->rerun_index.empty()) return 0;
  auto at = [&](const void* dev) { return host + (reinterpret_cast<const char*>(dev) - src); };
  const double* gap = reinterpret_cast<const double*>(at(b->d.guard_gap));
  const int32_t* flag = reinterpret_cast<const int32_t*>(at(b->d.guard_flag));
  const int n_rec = std::min(*reinterpret_cast<const int32_t*>(at(b->d.guard_n)), b->d.guard_cap);
  const int4* rec = reinterpret_cast<const int4*>(at(b->d.guard_rec));
  const int32_t* paths = reinterpret_cast<const int32_t*>(at(b->d.guard_paths));
  b->rerun_index.assign(b->S, -1);
  std::vector<char> need_full(b->S, 0);             // by sorted position
  double mg = INFINITY;
  b->n_flagged = b->n_pairs = b->n_rerun = 0;
  for (int p = 0; p < b->S; ++p) {
    mg = std::min(mg, gap[p]);
    if (flag[p]) b->n_flagged += 1;
    if (flag[p] & 2) need_full[p] = 1;
  }
  b->min_gap = mg;
  if (b->n_flagged == 0) return 0;
  // The guard's own kernels are launched plain: chained, their successors sit resident on the SMs (the float64
  // output-layer stream with 71 KB of shared memory each) while they wait, in the way of the main stream's next
  // tensor-core kernel - measured with three batches in flight: 1.44 -> 1.30 M chars/s at cfg 2; and at beam 50
  // (cfg 5, 50-row float64 GEMMs) the chained re-decode is slower even alone (20.9 vs 17.7 ms).
  static const int guard_pdl = [] {
    const char* e = getenv("JLM_PDL_GUARD");
    return e ? atoi(e) : 0;
  }();
  const PdlOff plain(!guard_pdl);
  static const bool dbg = getenv("JLM_DEBUG_TIMING") != nullptr;
  const auto t_0 = std::chrono::steady_clock::now();
  if (b->h->guard_verify) {
    // in slices of 256 records: a slice's trie of word prefixes (<= 2 x 256 paths) always fits the state pool
    b->n_lse_rows = 0;
    for (int r0 = 0; r0 < n_rec; r0 += 256) JLM_TRY(guard_verify_pairs(b, r0, std::min(n_rec, r0 + 256), rec, paths, need_full));
    if (dbg)
      fprintf(stderr, "[jlm] guard: %d sentences flagged, %d records, %d pairs re-scored, %d lse rows, %.3f ms\n", b->n_flagged,
              n_rec, b->n_pairs, b->n_lse_rows,
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_0).count());
  } else {
    for (int p = 0; p < b->S; ++p)
      if (flag[p]) need_full[p] = 1;
  }
  std::vector<int> which;
  for (int p = 0; p < b->S; ++p)
    if (need_full[p]) which.push_back(b->order[p]);
  b->n_rerun = (int)which.size();
  if (which.empty()) return 0;
  b->h->tier2_recent = 16;
  std::sort(which.begin(), which.end());
  GuardLattice& G = *b->guard_lat;
  const jlm_lattice_batch sub = G.subset(which, b->mode);
  // tier 2 on the guard stream as well: on the main stream it would wait behind every batch already enqueued
  jlm_handle* h = b->h;
  JLM_TRY(guard_stream_create(h));
  cudaStream_t main_stream = h->stream;
  h->stream = h->guard_stream;
  jlm_batch* r = nullptr;
  int32_t rc = jlm_batch_upload(h, &sub, b->unlimited ? JLM_BEAM_UNLIMITED : b->W, b->topN, b->mode, JLM_BACKEND_EXACT, &r);
  b->rerun = r;
  if (!rc) rc = jlm_batch_run(r);
  if (!rc) {
    const size_t n = which.size() * (size_t)r->topN;
    G.r_max_len = r->max_len;
    G.r_score.resize(n);
    G.r_npaths.resize(which.size());
    G.r_len.resize(n);
    G.r_nodes.resize(n * r->max_len);
    jlm_nbest nb{r->topN, r->max_len, G.r_score.data(), G.r_npaths.data(), G.r_len.data(), G.r_nodes.data()};
    rc = jlm_batch_fetch(r, &nb);
  }
  h->stream = main_stream;
  if (rc) return rc;
  if (dbg)
    fprintf(stderr, "[jlm] guard: tier 2 re-decoded %d sentences in float64, %.3f ms since the fetch\n", (int)which.size(),
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_0).count());
  for (size_t k = 0; k < which.size(); ++k) b->rerun_index[which[k]] = (int)k;
  return 0;
}

void beam_free_guard(jlm_handle* h) {
  if (h->guard_pool) jlm_pool_destroy(h->guard_pool);
  h->guard_pool = nullptr;
  if (h->guard_stream) cudaStreamDestroy(h->guard_stream);
  h->guard_stream = nullptr;
}

// Optional copy streams of the handle (JLM_COPY_STREAM=1): plan upload and n-best download beside the neighbouring
// batches' kernels instead of between them (two streams - on one, the upload of batch k+1 would queue behind the
// download of batch k, which waits for batch k's kernels).  OFF by default: the ~0.35 ms the in-stream copies leave
// the SMs idle per batch is where the near-tie guard's float64 kernels of the previous batch run; measured with three
// batches in flight (cfg 2): 1.30 M chars/s with the copies in-stream, 1.17 M with the copy streams.
static cudaStream_t copy_stream_of(jlm_handle* h, int which) {
  static const int on = [] {
    const char* e = getenv("JLM_COPY_STREAM");
    return e ? atoi(e) : 0;
  }();
  if (!on) return h->stream;
  cudaStream_t& cs = h->copy_stream[which];
  if (!cs && cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    cs = nullptr;
    return h->stream;
  }
  return cs;
}

void beam_free_plan_scratch(jlm_handle* h) {
  delete static_cast<HostPlan*>(h->plan_scratch);
  h->plan_scratch = nullptr;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" int32_t jlm_batch_upload(jlm_handle* h, const jlm_lattice_batch* lat, int32_t beam_width, int32_t top_n,
                                    int32_t mode, int32_t backend, jlm_batch** out) {
  JLM_REQUIRE(h && lat && out, "jlm_batch_upload: null argument");
  JLM_REQUIRE(beam_width == JLM_BEAM_UNLIMITED || (beam_width >= 1 && beam_width <= JLM_MAX_BEAM),
              "jlm_batch_upload: beam_width %d not in [1,%d] (or JLM_BEAM_UNLIMITED)", beam_width, JLM_MAX_BEAM);
  JLM_REQUIRE(top_n >= 1, "jlm_batch_upload: top_n must be >= 1");
  JLM_REQUIRE(mode >= 0 && mode <= 2, "jlm_batch_upload: bad mode %d", mode);
  JLM_REQUIRE(backend == JLM_BACKEND_AUTO || backend == JLM_BACKEND_EXACT || backend == JLM_BACKEND_TC,
              "jlm_batch_upload: bad backend %d", backend);
  JLM_CUDA(cudaSetDevice(h->device));
  *out = nullptr;
  jlm_batch* b = new jlm_batch();
  b->h = h;
  b->unlimited = beam_width == JLM_BEAM_UNLIMITED;
  b->W = beam_width;      // unlimited: set by build_plan to the widest frame
  b->topN = b->unlimited ? top_n : std::min(top_n, beam_width);
  b->mode = mode;
  b->dynamic = mode == JLM_DECODE_DYNAMIC;
  b->use_lse = h->cfg.self_norm == 0;
  // the plan's host vectors live with the handle so that steady-state uploads touch no fresh memory
  if (!h->plan_scratch) h->plan_scratch = new HostPlan();
  HostPlan& P = *static_cast<HostPlan*>(h->plan_scratch);
  P.vocab_jobs.clear();
  P.dyn_info.clear();
  P.vocab_cols.clear();
  P.vfp.clear();
  static const bool dbg = getenv("JLM_DEBUG_TIMING") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](auto a, auto c) { return std::chrono::duration<double, std::milli>(c - a).count(); };
  const auto t_0 = now();
  if (build_plan(b, lat, P)) {
    delete b;
    return 1;
  }
  const auto t_1 = now();
  if (backend == JLM_BACKEND_AUTO) backend = (b->max_rows_step >= 512) ? JLM_BACKEND_TC : JLM_BACKEND_EXACT;
  b->backend = backend;
  // near-tie guard: tensor-core back end only (the float64 back end IS the fallback); nothing to rank when unpruned
  b->guard_eps = (backend == JLM_BACKEND_TC && !b->unlimited) ? h->guard_eps : 0.0;
  b->guard_all = h->guard_all;
  if (b->guard_eps > 0.0) {
    b->guard_lat = new GuardLattice();
    b->guard_lat->copy_from(lat, mode);
  }
  int32_t rc = 0;
  Arena a;
  a.begin_plan();
  layout(b, a, P);
  if (backend == JLM_BACKEND_TC) rc = tc_batch_plan(b, a);
  if (!rc) {
    // reuse a cached arena: the smallest one that is large enough, else grow the largest
    const size_t need = a.off;
    int pick = -1;
    for (int i = 0; i < (int)h->batch_cache.size(); ++i) {
      const size_t cap = h->batch_cache[i].cap;
      if (pick < 0) pick = i;
      else {
        const size_t pc = h->batch_cache[pick].cap;
        const bool better = (cap >= need && (pc < need || cap < pc)) || (cap < need && pc < need && cap > pc);
        if (better) pick = i;
      }
    }
    if (pick >= 0) {
      std::swap(b->mem, h->batch_cache[pick]);
      h->batch_cache.erase(h->batch_cache.begin() + pick);
    }
    std::swap(a.buf, b->mem);
    rc = a.commit();
    std::swap(a.buf, b->mem);
  }
  if (!rc) {
    a.buf = b->mem;
    a.off = 0;
    a.dry = false;
    layout(b, a, P);
    const size_t plan_bytes = reinterpret_cast<char*>(b->d.slot_score) - reinterpret_cast<char*>(b->mem.p);
    if (backend == JLM_BACKEND_TC) rc = tc_batch_plan(b, a);
    a.buf = DevBuf();  // ownership stays with b->mem
    // pinned staging slot: wait only for the copy that last used this slot, never for the stream
    const int slot = h->stage_next;
    h->stage_next = (h->stage_next + 1) % jlm_handle::N_STAGE;
    if (!rc && h->stage_busy[slot]) {
      if (cudaEventSynchronize(h->stage_ev[slot]) != cudaSuccess) {
        jlm_set_error("jlm_batch_upload: staging event sync failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = 1;
      }
      h->stage_busy[slot] = false;
    }
    if (!rc && !h->stage_ev[slot] && cudaEventCreateWithFlags(&h->stage_ev[slot], cudaEventDisableTiming) != cudaSuccess) {
      jlm_set_error("jlm_batch_upload: cudaEventCreate failed");
      rc = 1;
    }
    if (!rc) rc = h->stage[slot].reserve(plan_bytes);
    if (!rc) {
      char* host = h->stage[slot].as<char>();
      char* base = static_cast<char*>(b->mem.p);
      stage(host, base, b->d.node_word, P.node_word);
      stage(host, base, b->d.node_pfid, P.node_pfid);
      if (b->dynamic) stage(host, base, b->d.node_span, P.node_span);
      stage(host, base, b->d.cand_pos, P.cand_pos);
      stage(host, base, b->d.frame_lo, P.frame_lo);
      stage(host, base, b->d.frame_hi, P.frame_hi);
      stage(host, base, b->d.frame_cand_lo, P.frame_cand_lo);
      stage(host, base, b->d.frame_ncand, P.frame_ncand);
      stage(host, base, b->d.frame_minpf, P.frame_minpf);
      stage(host, base, b->d.bc, P.bc);
      stage(host, base, b->d.slot0, P.slot0);
      stage(host, base, b->d.fbase, P.fbase);
      stage(host, base, b->d.sent_T, P.sent_T);
      stage(host, base, b->d.start_items, P.start_items);
      stage(host, base, b->d.vocab_jobs, P.vocab_jobs);
      stage(host, base, b->d.dyn_info, P.dyn_info);
      stage(host, base, b->d.vocab_cols, P.vocab_cols);
      stage(host, base, b->d.vfp, P.vfp);
      // The copy goes through the handle's copy stream and the compute stream waits for its event: enqueued on the
      // compute stream itself it would sit between the previous batch's last kernel and this batch's first one with
      // the SMs idle (~0.25 ms per 1024-sentence plan); the arena is idle (jlm_batch_destroy waited for its last user).
      cudaStream_t cs = copy_stream_of(h, 0);
      if (cudaMemcpyAsync(base, host, plan_bytes, cudaMemcpyHostToDevice, cs) != cudaSuccess ||
          cudaEventRecord(h->stage_ev[slot], cs) != cudaSuccess ||
          (cs != h->stream && cudaStreamWaitEvent(h->stream, h->stage_ev[slot], 0) != cudaSuccess)) {
        jlm_set_error("jlm_batch_upload: H2D copy failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = 1;
      } else {
        h->stage_busy[slot] = true;
      }
      b->h2d_bytes = (int64_t)plan_bytes;
    }
  }
  if (dbg) fprintf(stderr, "[jlm] upload: build_plan %.3f ms, layout+stage+H2D enqueue %.3f ms\n", ms(t_0, t_1), ms(t_1, now()));
  if (rc) {
    jlm_batch_destroy(b);
    return 1;
  }
  *out = b;
  return 0;
}

extern "C" int32_t jlm_batch_run(jlm_batch* b) {
  JLM_REQUIRE(b, "jlm_batch_run: null batch");
  jlm_handle* h = b->h;
  JLM_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  b->launches = 0;
  guard_drop_rerun(b);
  // Launch chaining (pdl_enter, jlm_common.cuh) removes the points where the device drains between kernels - and those
  // are where the near-tie guard's float64 kernels (own stream, highest priority) get their SMs.  Measured with three
  // batches in flight and a float64 re-decode in every batch (bench.py e2e arm, cfg 2 / cfg 5): 1.41 -> 1.30 and
  // 0.235 -> 0.213 M chars/s with chaining, while batches without re-decodes gain (cfg 3 e2e 1.98 -> 2.15 M, blocking
  // cfg 2 16.4 -> 15.9 ms).  So a batch is chained unless another batch is still in flight AND one of the last
  // sixteen batches needed a re-decode.
  const bool main_tc = b->backend == JLM_BACKEND_TC && st != h->guard_stream;
  static const int adaptive = [] {
    const char* e = getenv("JLM_PDL_ADAPTIVE");
    return e ? atoi(e) : 1;
  }();
  const PdlOff plain(adaptive && main_tc && h->batches_unfetched > 0 && h->tier2_recent > 0);
  if (main_tc) {
    if (h->tier2_recent > 0) --h->tier2_recent;
    if (!b->counted) {
      b->counted = true;
      ++h->batches_unfetched;
    }
  }
  if (b->backend == JLM_BACKEND_TC && score_overlap_enabled()) {
    if (!h->side_stream) JLM_CUDA(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
    if (!b->ev_fork) JLM_CUDA(cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming));
    if (!b->ev_join) JLM_CUDA(cudaEventCreateWithFlags(&b->ev_join, cudaEventDisableTiming));
  }
  if (b->timers && b->events.empty()) {
    b->events.resize(3 * (size_t)b->n_steps + 1);
    for (auto& e : b->events) JLM_CUDA(cudaEventCreate(&e));
    b->kev.resize(4 * (size_t)b->n_steps);
    for (auto& e : b->kev) JLM_CUDA(cudaEventCreate(&e));
  }
  bool single = false;
  JLM_TRY(single_try_run(b, &single));
  for (int t = 0; t < b->n_steps && !single; ++t) {
    if (b->timers) cudaEventRecord(b->events[3 * t], st);
    if (t == 0) {
      JLM_CUDA(jlm_launch(k_init_frame0, dim3(ceil_div(b->S, 128)), dim3(128), 0, st, b->d, b->S));
      JLM_CUDA(cudaGetLastError());
      JLM_CUDA(jlm_launch(k_build_items, dim3(ceil_div(b->d.n_items, 256)), dim3(256), 0, st, b->d));
      JLM_CUDA(cudaGetLastError());
      b->launches += 2;
    } else if (b->dynamic) {
      JLM_TRY(launch_prune<true>(b, t));
    } else {
      JLM_TRY(launch_prune<false>(b, t));
    }
    if (b->backend == JLM_BACKEND_EXACT) {
      JLM_TRY(exact_lm_step(b, t));
    } else {
      const float* T32 = nullptr;
      int ldt = 0;
      JLM_TRY(tc_batch_lm_state(b, t, &T32, &ldt));
      // Full softmax: the needed-word dot products depend on the stage-1 rows only, not on the LSE, so they
      // CAN run on a side stream under the output GEMM, a one-pass kernel adding parent score + LSE after the
      // merge (JLM_OVERLAP_SCORE=1; measured slower, see score_overlap_enabled).
      const bool overlap = T32 && b->use_lse && b->mode == JLM_DECODE_FULL && b->steps[t].n_items > 0 &&
                           score_overlap_enabled() && h->side_stream;
      if (overlap) {
        JLM_CUDA(cudaEventRecord(b->ev_fork, st));
        JLM_CUDA(cudaStreamWaitEvent(h->side_stream, b->ev_fork, 0));
        JLM_TRY(launch_score<float>(b, t, T32, ldt, h->side_stream, 1));
        JLM_CUDA(cudaEventRecord(b->ev_join, h->side_stream));
      }
      JLM_TRY(tc_batch_lm_lse(b, t));
      if (overlap) {
        const StepPlan& sp = b->steps[t];
        JLM_CUDA(cudaStreamWaitEvent(st, b->ev_join, 0));
        JLM_CUDA(jlm_launch(k_add_base, dim3(ceil_div((int64_t)sp.n_items * b->W, 256)), dim3(256), 0, st, b->d, sp.item0, sp.n_items, b->W));
        JLM_CUDA(cudaGetLastError());
        b->launches += 1;
      } else if (T32) {
        JLM_TRY(lm_step_tail<float>(b, t, T32, ldt));
      }
    }
  }
  if (b->timers) cudaEventRecord(b->events[3 * (size_t)b->n_steps], st);
  if (b->guard_eps > 0.0) {
    JLM_CUDA(jlm_launch(k_guard_paths, dim3(ceil_div(2 * (int64_t)b->d.guard_cap, 128)), dim3(128), 0, st, b->d, b->max_len));
    JLM_CUDA(cudaGetLastError());
    b->launches += 1;
  }
  JLM_CUDA(jlm_launch(k_backtrace, dim3(ceil_div((int64_t)b->S * b->topN, 128)), dim3(128), 0, st, b->d, b->S, b->topN, b->max_len));
  JLM_CUDA(cudaGetLastError());
  b->launches += 1;
  b->ran = true;
  b->d2h_queued = false;
  if (!b->done) JLM_CUDA(cudaEventCreateWithFlags(&b->done, cudaEventDisableTiming));
  JLM_CUDA(cudaEventRecord(b->done, st));
  return 0;
}

// Enqueues the device->host copy of the n-best block behind the batch's kernels and re-records the
// completion event.  Asynchronous: jlm_batch_fetch then only waits for this batch, so a caller can
// upload + run the NEXT batch before fetching this one and the device never idles between batches.
extern "C" int32_t jlm_batch_fetch_async(jlm_batch* b) {
  JLM_REQUIRE(b, "jlm_batch_fetch_async: null batch");
  JLM_REQUIRE(b->ran, "jlm_batch_fetch_async: jlm_batch_run has not been called");
  if (b->d2h_queued) return 0;
  jlm_handle* h = b->h;
  JLM_CUDA(cudaSetDevice(h->device));
  const size_t n = (size_t)b->S * b->topN;
  const char* src = reinterpret_cast<const char*>(b->d.guard_gap);
  // guard_gap, guard_flag, out_score .. out_nodes are consecutive 256-aligned arena blocks: one transfer
  const size_t total = (reinterpret_cast<const char*>(b->d.out_nodes) - src) + n * b->max_len * sizeof(int32_t);
  if (!b->out_host.p && !h->out_pool.empty()) {
    int pick = 0;   // smallest pooled buffer that fits, else the largest
    for (int i = 1; i < (int)h->out_pool.size(); ++i) {
      const size_t c = h->out_pool[i].cap, pc = h->out_pool[pick].cap;
      if ((c >= total && (pc < total || c < pc)) || (c < total && pc < total && c > pc)) pick = i;
    }
    b->out_host = h->out_pool[pick];
    h->out_pool.erase(h->out_pool.begin() + pick);
  }
  JLM_TRY(b->out_host.reserve(total));
  // behind this batch's kernels (b->done, recorded by jlm_batch_run) but on the copy stream: the next batch's
  // kernels, already enqueued on the compute stream, do not wait for the transfer
  cudaStream_t cs = copy_stream_of(h, 1);
  if (cs != h->stream) JLM_CUDA(cudaStreamWaitEvent(cs, b->done, 0));
  JLM_CUDA(cudaMemcpyAsync(b->out_host.p, src, total, cudaMemcpyDeviceToHost, cs));
  JLM_CUDA(cudaEventRecord(b->done, cs));
  b->d2h_bytes = (int64_t)total;
  b->d2h_queued = true;
  return 0;
}

extern "C" int32_t jlm_batch_fetch(jlm_batch* b, jlm_nbest* out) {
  JLM_REQUIRE(b && out, "jlm_batch_fetch: null argument");
  JLM_REQUIRE(b->ran, "jlm_batch_fetch: jlm_batch_run has not been called");
  JLM_REQUIRE(out->top_n >= b->topN && out->max_len >= b->max_len,
              "jlm_batch_fetch: output capacity too small (top_n %d < %d or max_len %d < %d)", out->top_n, b->topN,
              out->max_len, b->max_len);
  jlm_handle* h = b->h;
  JLM_CUDA(cudaSetDevice(h->device));
  JLM_TRY(jlm_batch_fetch_async(b));
  JLM_CUDA(cudaEventSynchronize(b->done));   // this batch only: later batches keep running
  if (b->counted) {
    b->counted = false;
    --h->batches_unfetched;
  }
  const char* src = reinterpret_cast<const char*>(b->d.guard_gap);
  const char* host = b->out_host.as<char>();
  const double* sc = reinterpret_cast<const double*>(host + (reinterpret_cast<const char*>(b->d.out_score) - src));
  const int32_t* np_ = reinterpret_cast<const int32_t*>(host + (reinterpret_cast<const char*>(b->d.out_npaths) - src));
  const int32_t* ln = reinterpret_cast<const int32_t*>(host + (reinterpret_cast<const char*>(b->d.out_len) - src));
  const int32_t* nd = reinterpret_cast<const int32_t*>(host + (reinterpret_cast<const char*>(b->d.out_nodes) - src));
  JLM_TRY(guard_resolve(b, host, src));
  for (int p = 0; p < b->S; ++p) {
    const int s = b->order[p];
    const int fk = b->rerun_index.empty() ? -1 : b->rerun_index[s];
    if (fk >= 0) {      // flagged by the near-tie guard: the float64 re-decode replaces the tensor-core lists
      const GuardLattice& G = *b->guard_lat;
      const int rt = b->rerun->topN, rl = G.r_max_len;
      out->n_paths[s] = G.r_npaths[fk];
      for (int k = 0; k < out->top_n; ++k) {
        const size_t o = (size_t)s * out->top_n + k;
        if (k < rt) {
          const size_t i = (size_t)fk * rt + k;
          out->scores[o] = G.r_score[i];
          out->path_len[o] = G.r_len[i];
          for (int q = 0; q < std::min(G.r_len[i], rl); ++q)
            out->path_nodes[o * out->max_len + q] = (int32_t)(G.r_nodes[i * rl + q] + G.f_node_delta[fk]);
        } else {
          out->scores[o] = INFINITY;
          out->path_len[o] = 0;
        }
      }
      continue;
    }
    out->n_paths[s] = np_[p];
    for (int k = 0; k < out->top_n; ++k) {
      const size_t o = (size_t)s * out->top_n + k;
      if (k < b->topN) {
        const size_t i = (size_t)p * b->topN + k;
        out->scores[o] = sc[i];
        out->path_len[o] = ln[i];
        memcpy(out->path_nodes + o * out->max_len, nd + i * b->max_len, sizeof(int32_t) * std::min(ln[i], b->max_len));
      } else {
        out->scores[o] = INFINITY;
        out->path_len[o] = 0;
      }
    }
  }
  if (b->timers && !b->events.empty()) {
    b->ms_lstm = b->ms_softmax = b->ms_beam = 0.f;
    for (int t = 0; t < b->n_steps; ++t) {
      if (b->steps[t].rows_step == 0) {
        float x = 0.f;
        cudaEventElapsedTime(&x, b->events[3 * t], b->events[3 * t + 3]);
        b->ms_beam += x;
        continue;
      }
      float x = 0.f, y = 0.f, z = 0.f;
      cudaEventElapsedTime(&x, b->events[3 * t], b->events[3 * t + 1]);
      cudaEventElapsedTime(&y, b->events[3 * t + 1], b->events[3 * t + 2]);
      cudaEventElapsedTime(&z, b->events[3 * t + 2], b->events[3 * t + 3]);
      b->ms_beam += x;
      b->ms_lstm += y;
      b->ms_softmax += z;
    }
    b->ms_gate_gemm = b->ms_proj_gemm = 0.f;
    b->n_gate_launches = b->n_proj_launches = 0;
    if (b->backend == JLM_BACKEND_TC) {
      const bool proj = b->mode == JLM_DECODE_FULL && b->use_lse;
      for (int t = 0; t < b->n_steps; ++t) {
        if (b->steps[t].rows_step == 0) continue;
        float x = 0.f;
        cudaEventElapsedTime(&x, b->kev[4 * t], b->kev[4 * t + 1]);
        b->ms_gate_gemm += x;
        b->n_gate_launches += 1;
        if (proj) {
          cudaEventElapsedTime(&x, b->kev[4 * t + 2], b->kev[4 * t + 3]);
          b->ms_proj_gemm += x;
          b->n_proj_launches += b->h->n_seg;
        }
      }
    }
  }
  return 0;
}

extern "C" int32_t jlm_batch_destroy(jlm_batch* b) {
  if (!b) return 0;
  cudaSetDevice(b->h->device);
  // wait for this batch's own work only (its arena and pinned buffer go back to the pools)
  if (b->done) {
    cudaEventSynchronize(b->done);
    cudaEventDestroy(b->done);
  } else {
    cudaStreamSynchronize(b->h->stream);
  }
  if (b->counted) --b->h->batches_unfetched;
  b->counted = false;
  guard_drop_rerun(b);
  delete b->guard_lat;
  b->guard_lat = nullptr;
  if (b->ev_fork) cudaEventDestroy(b->ev_fork);
  if (b->ev_join) cudaEventDestroy(b->ev_join);
  if (b->out_host.p) {
    if (b->h->out_pool.size() < 8) b->h->out_pool.push_back(b->out_host);
    else b->out_host.release();
    b->out_host = HostBuf();
  }
  tc_batch_free(b);
  for (auto& e : b->events) cudaEventDestroy(e);
  for (auto& e : b->kev) cudaEventDestroy(e);
  if (b->mem.p && b->h->batch_cache.size() < 8) {
    b->h->batch_cache.push_back(b->mem);      // keep the arena for the next upload
    b->mem = DevBuf();
  }
  b->mem.release();
  delete b;
  return 0;
}

extern "C" int32_t jlm_decode_batch(jlm_handle* h, const jlm_lattice_batch* lat, int32_t beam_width, int32_t top_n,
                                    int32_t mode, int32_t backend, jlm_nbest* out) {
  jlm_batch* b = nullptr;
  JLM_TRY(jlm_batch_upload(h, lat, beam_width, top_n, mode, backend, &b));
  int32_t rc = jlm_batch_run(b);
  if (!rc) rc = jlm_batch_fetch(b, out);
  jlm_batch_destroy(b);
  return rc;
}

extern "C" int32_t jlm_batch_get_info(jlm_batch* b, jlm_batch_info* info) {
  JLM_REQUIRE(b && info, "jlm_batch_get_info: null argument");
  int64_t stepped = 0;
  for (auto& s : b->steps) stepped += s.rows_step;
  info->n_slots = stepped;
  info->n_candidates = b->n_cand;
  info->n_nodes = b->N;
  info->n_steps = b->n_steps;
  info->backend = b->backend;
  info->kernel_launches = b->launches;
  info->beam_width = b->W;
  info->n_guard_flagged = b->n_flagged;
  info->n_guard_pairs = b->n_pairs;
  info->n_guard_rerun = b->n_rerun;
  info->guard_min_gap = b->guard_eps > 0.0 ? b->min_gap : 0.0;
  info->guard_eps = b->guard_eps;
  info->h2d_bytes = b->h2d_bytes;
  info->d2h_bytes = b->d2h_bytes;
  info->ms_lstm = b->ms_lstm;
  info->ms_softmax = b->ms_softmax;
  info->ms_beam = b->ms_beam;
  info->ms_gate_gemm = b->ms_gate_gemm;
  info->ms_proj_gemm = b->ms_proj_gemm;
  info->n_gate_launches = b->n_gate_launches;
  info->n_proj_launches = b->n_proj_launches;
  return 0;
}

extern "C" int32_t jlm_batch_enable_timers(jlm_batch* b, int32_t on) {
  JLM_REQUIRE(b, "jlm_batch_enable_timers: null batch");
  b->timers = on != 0;
  return 0;
}

extern "C" int32_t jlm_batch_get_beams(jlm_batch* b, int32_t sentence, int32_t* count, double* score,
                                       int32_t* parent_frame, int32_t* parent_rank, int32_t* node, double* lse,
                                       double* h_out, double* c_out) {
  JLM_REQUIRE(b && b->ran, "jlm_batch_get_beams: run the batch first");
  JLM_REQUIRE(sentence >= 0 && sentence < b->S, "jlm_batch_get_beams: sentence out of range");
  if (b->rerun && !b->rerun_index.empty() && b->rerun_index[sentence] >= 0) {
    // flagged by the near-tie guard: the beams that produced the returned n-best are the float64 re-decode's
    const int fk = b->rerun_index[sentence];
    JLM_TRY(jlm_batch_get_beams(b->rerun, fk, count, score, parent_frame, parent_rank, node, lse, h_out, c_out));
    if (node) {      // node indices of the sub-batch -> the caller's lattice
      const jlm_batch* r = b->rerun;
      int q = 0;
      while (r->order[q] != fk) ++q;
      for (int t = 0; t <= r->sent_T[q]; ++t)
        for (int k = 0; k < r->bc[r->fbase[q] + t]; ++k) node[(size_t)t * r->W + k] += (int32_t)b->guard_lat->f_node_delta[fk];
    }
    return 0;
  }
  jlm_handle* h = b->h;
  JLM_CUDA(cudaSetDevice(h->device));
  JLM_CUDA(cudaStreamSynchronize(h->stream));
  int p = 0;
  while (b->order[p] != sentence) ++p;
  const int T = b->sent_T[p];
  const int W = b->W;
  std::vector<int32_t> par(W);
  std::vector<double> tmp((size_t)W * h->Hp);
  for (int t = 0; t <= T; ++t) {
    const int64_t fid = b->fbase[p] + t;
    const int cnt = b->bc[fid];
    const int64_t s0 = b->slot0[fid];
    if (count) count[t] = cnt;
    if (cnt == 0) continue;
    const size_t o = (size_t)t * W;
    if (score) JLM_CUDA(cudaMemcpy(score + o, b->d.slot_score + s0, sizeof(double) * cnt, cudaMemcpyDeviceToHost));
    if (node) JLM_CUDA(cudaMemcpy(node + o, b->d.slot_node + s0, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost));
    if (lse) {
      if (b->dynamic && b->use_lse && t < T) {
        // report LSE over lattice_vocab[t+1], the first set the row is scored under
        for (int k = 0; k < cnt; ++k)
          JLM_CUDA(cudaMemcpy(lse + o + k, b->d.dyn_lse + (s0 + k) * (b->Tmax + 1) + t + 1, sizeof(double),
                              cudaMemcpyDeviceToHost));
      } else {
        JLM_CUDA(cudaMemcpy(lse + o, b->d.slot_lse + s0, sizeof(double) * cnt, cudaMemcpyDeviceToHost));
      }
    }
    if (parent_frame || parent_rank) {
      JLM_CUDA(cudaMemcpy(par.data(), b->d.slot_parent + s0, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost));
      for (int k = 0; k < cnt; ++k) {
        int pf = -1, pr = -1;
        if (par[k] >= 0) {
          for (int q = 0; q < t; ++q) {
            const int64_t f2 = b->fbase[p] + q;
            if (par[k] >= b->slot0[f2] && par[k] < b->slot0[f2] + b->bc[f2]) {
              pf = q;
              pr = (int)(par[k] - b->slot0[f2]);
            }
          }
        }
        if (parent_frame) parent_frame[o + k] = pf;
        if (parent_rank) parent_rank[o + k] = pr;
      }
    }
    if ((h_out || c_out) && t < T) {
      if (b->backend == JLM_BACKEND_EXACT) {
        if (h_out) {
          JLM_CUDA(cudaMemcpy(tmp.data(), b->hx + s0 * h->Hp, sizeof(double) * cnt * h->Hp, cudaMemcpyDeviceToHost));
          for (int k = 0; k < cnt; ++k) memcpy(h_out + (o + k) * h->H, &tmp[(size_t)k * h->Hp], sizeof(double) * h->H);
        }
        if (c_out) {
          JLM_CUDA(cudaMemcpy(tmp.data(), b->cx + s0 * h->Hp, sizeof(double) * cnt * h->Hp, cudaMemcpyDeviceToHost));
          for (int k = 0; k < cnt; ++k) memcpy(c_out + (o + k) * h->H, &tmp[(size_t)k * h->Hp], sizeof(double) * h->H);
        }
      } else {
        JLM_TRY(tc_batch_get_state(b, s0, cnt, h_out ? h_out + o * h->H : nullptr, c_out ? c_out + o * h->H : nullptr));
      }
    }
  }
  return 0;
}
