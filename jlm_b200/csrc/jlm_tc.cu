// Tensor-core back end (sm_100a): TMA-fed tcgen05 GEMMs with fused epilogues.
//
// Precision.  The reference evaluates hidden -> logits in float64 over float32 weights, and beam
// indices are only reproducible with ~fp32-accurate logits.  fp16 tensor-core
// operands are therefore 2-term splits x*s = hi + lo (hi = fp16(x*s), lo = fp16(x*s - hi), s a power
// of two that places the tensor's max near 2^15 so `lo` stays a normal fp16): 22 significant bits
// per operand.  One product is three MMAs (hi.lo + lo.hi + hi.hi; lo.lo ~ 2^-22 is dropped) into
// one fp32 TMEM accumulator, un-scaled in the epilogue.
//
// Kernel.  One persistent CTA per SM: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane)
// + TMEM owner, then 4 epilogue warps (one TMEM lane quarter each; the LSTM epilogue runs 8, two per
// quarter splitting the tile's columns, because its math is what paces that kernel).  128 x BN x 64
// tiles, smem ring of {A_hi, A_lo, B_hi, B_lo} stages in SWIZZLE_128B layout, two TMEM accumulator
// stages so the epilogue of tile i overlaps the MMAs of tile i+1.  Epilogues:
//   EPI_LSTM  : gate bias + sigmoid/tanh + cell update (decoder/model.py:132-139), writes h, c and
//               the fp16 split of h for the next GEMM.
//   EPI_STORE : un-scale (+bias), write fp32 and/or the fp16 split (stage-1 projection h.PM).
//   EPI_LSE   : + b2, online per-row (max, sum exp) over the tile's columns -> partials
//               (softmax, decoder/model.py:15-20, never materialising [B,V]).
#include <cmath>

#include "jlm_beam.cuh"
#include "jlm_tc_ptx.cuh"

struct TcOperand {
  __half* hi = nullptr;
  __half* lo = nullptr;
  int64_t rows = 0, K = 0;
  float scale = 1.f;
  CUtensorMap map_hi, map_lo;         // box = box_rows x 64 (one CTA loads the whole N tile)
  CUtensorMap pair_hi, pair_lo;       // box = box_rows/2 x 64 (each CTA of a pair loads half of the N tile)
  CUtensorMap q_hi, q_lo;             // box = box_rows/4 x 64 (narrow 64-column tiles)
};

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_TILE = BM * BK * 2;  // bytes of one fp16 A tile
constexpr float LOG2E = 1.4426950408889634f;

enum { EPI_LSE = 0, EPI_STORE = 1, EPI_LSTM = 2 };

template <int EPI>
struct EpiCfg {
  static constexpr int WARPS = (EPI == EPI_LSTM) ? 8 : 4;   // epilogue warps
  static constexpr int THREADS = 64 + 32 * WARPS;
  // LSTM epilogue: per-warp transpose buffer, 32 rows x (64 B payload + 16 B pad); see stage_rows / unstage_rows
  static constexpr int STG_ROW = 80;
  static constexpr int STG_WARP = 32 * STG_ROW;
  static constexpr int STG = (EPI == EPI_LSTM) ? WARPS * STG_WARP : 0;
};

// sigmoid / tanh straight on the MUFU pipe (ex2.approx + rcp.approx, ~2e-7 absolute): the gate
// epilogue is instruction-paced, and these stay an order of magnitude inside the split-fp16 GEMM's
// own error.  Saturation needs no clamps: ex2 -> +inf gives rcp -> 0.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x on the FMA / ALU pipes (no MUFU): round-to-nearest split x = n + f with the 1.5 * 2^23 trick, degree-6 polynomial
// for 2^f on [-0.5, 0.5] (Cephes exp2f coefficients, ~1e-7 relative), 2^n by adding n to the exponent field.  Twelve
// issue slots instead of one MUFU: used for a fraction of the logits where the MUFU queue is what paces an epilogue.
// x is clamped at -125 (result 2^-125 instead of 0: negligible against a sum that is >= 1).
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;
  const float f = x - (t - 12582912.f);
  float p = 1.535336188319500e-4f;
  p = fmaf(p, f, 1.339887440266574e-3f);
  p = fmaf(p, f, 9.618437357674640e-3f);
  p = fmaf(p, f, 5.550332471162809e-2f);
  p = fmaf(p, f, 2.402264791363012e-1f);
  p = fmaf(p, f, 6.931472028550421e-1f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ float fast_sigmoid(float x) { return rcp_approx(1.f + ex2_approx(-LOG2E * x)); }
__device__ __forceinline__ float fast_tanh(float x) {
  return fmaf(-2.f, rcp_approx(ex2_approx((2.f * LOG2E) * x) + 1.f), 1.f);   // 1 - 2/(e^{2x}+1)
}

// CG = 1: one CTA per tile (128 x BN).  CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) per
// 256 x BN tile - each CTA stages its own 128 rows of A and HALF of the B tile, so a stage is 64 KB
// instead of 96 KB and the per-MMA shared-memory traffic (operand reads + TMA writes), which is what
// caps the single-CTA kernel at ~88 % tensor-pipe, drops from 20 KB to 13.3 KB per 128 cycles.
// KS > 1: the K range is cut into KS chunks, each accumulated into its OWN TMEM accumulator, and the epilogue adds
// the partials (fp32, round to nearest).  tcgen05 accumulates in fp32 with truncation (measured, scripts/tc_bias_probe.py: the
// error of C = A.B^T grows linearly with K, -3e-9 K relative, biased toward zero), so a short chain per
// accumulator is what buys accuracy; used for the stage-1 projection h.PM whose error every logit inherits.
template <int BN, int CG = 1, int KS = 1>
struct TileCfg {
  static constexpr int B_TILE = (BN / CG) * BK * 2;
  static constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;
  static constexpr int STAGES = (CG == 2) ? 3 : ((BN == 256) ? 2 : 3);
  static constexpr int BIAS_BYTES = 2 * BN * 4;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM = STAGES * STAGE + BIAS_BYTES + BAR_BYTES + 1024;
  static constexpr int ACC_COLS = BN * KS;          // TMEM columns of one accumulator stage
  static constexpr int TMEM_COLS = 2 * ACC_COLS;
  static_assert(TMEM_COLS <= 512, "two accumulator stages must fit the 512 TMEM columns");
};

struct GemmArgs {
  int M, N, K;
  int K16;              // MMA k-steps are issued up to this column (0: all of K); the rest of K is zero padding
  int num_m_blocks, num_n_blocks;
  int a_row0;           // first row of the A operand inside its tensor map (per-slot arrays: the step's first slot)
  float inv_scale;      // 1 / (scale_A * scale_B)
  const float* bias;    // [N] (tile order) or nullptr
  // EPI_LSE
  float2* part;
  int part_ld, part_tile0;
  // EPI_STORE
  float* C32;
  int64_t ldc;
  __half* s_hi;
  __half* s_lo;
  int64_t lds;
  float split_scale;
  // EPI_LSTM (row pointers already offset to the step's first row)
  const float* c_src;
  const int32_t* parent;
  float* h_out;
  float* c_out;
  int64_t ld_state;
};

__device__ __forceinline__ int ceil_div_dev(int a, int b) { return (a + b - 1) / b; }

// 16 consecutive values -> two 32-byte runs (hi, lo): x*scale = hi + lo in fp16
__device__ __forceinline__ void split_store16(__half* hi, __half* lo, const float* v, float scale) {
  uint32_t ph[8], pl[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float a = v[2 * j] * scale, b = v[2 * j + 1] * scale;
    const __half2 h2 = __floats2half2_rn(a, b);          // one F2FP.PACK_AB
    const float2 hf = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn(a - hf.x, b - hf.y);
    ph[j] = *reinterpret_cast<const uint32_t*>(&h2);
    pl[j] = *reinterpret_cast<const uint32_t*>(&l2);
  }
  uint4* dh = reinterpret_cast<uint4*>(hi);
  uint4* dl = reinterpret_cast<uint4*>(lo);
  dh[0] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
  dh[1] = make_uint4(ph[4], ph[5], ph[6], ph[7]);
  dl[0] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
  dl[1] = make_uint4(pl[4], pl[5], pl[6], pl[7]);
}

template <int BN, int EPI, int CG, int KS = 1>
__global__ void __launch_bounds__(EpiCfg<EPI>::THREADS, 1)
k_tc_gemm(const __grid_constant__ CUtensorMap mAh, const __grid_constant__ CUtensorMap mAl,
          const __grid_constant__ CUtensorMap mBh, const __grid_constant__ CUtensorMap mBl, const GemmArgs g) {
  using C = TileCfg<BN, CG, KS>;
  static_assert(KS == 1 || EPI == EPI_STORE, "partial accumulators are summed by the store epilogue only");
  // CTA pair: rank 0 is the leader (issues the MMAs, owns the full / tmem-empty barriers that count)
  const uint32_t rank = (CG == 2) ? ptx::cluster_ctarank() : 0u;
  const int unit = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;       // tile-loop index of this CTA (pair)
  const int n_units = (CG == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  float* bias_s = reinterpret_cast<float*>(gen + C::STAGES * C::STAGE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen + C::STAGES * C::STAGE + C::BIAS_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);
  uint8_t* epi_stage = gen + C::STAGES * C::STAGE + C::BIAS_BYTES + C::BAR_BYTES;   // EpiCfg<EPI>::STG bytes
  const uint32_t bar0 = ptx::smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * C::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * C::STAGES + 2 + a); };

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tensormap(&mAh);
    ptx::prefetch_tensormap(&mAl);
    ptx::prefetch_tensormap(&mBh);
    ptx::prefetch_tensormap(&mBl);
  }
  if (warp == 1) {
    if (ptx::elect_one()) {
      for (int s = 0; s < C::STAGES; ++s) {
        ptx::mbar_init(full_bar(s), 1);
        ptx::mbar_init(empty_bar(s), 1);
      }
      for (int a = 0; a < 2; ++a) {
        ptx::mbar_init(tfull_bar(a), 1);
        ptx::mbar_init(tempty_bar(a), CG * EpiCfg<EPI>::WARPS);   // pair: both CTAs' epilogue warps arrive on the leader's
      }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    if (CG == 2) {
      ptx::tmem_alloc_pair(ptx::smem_u32(tmem_slot), C::TMEM_COLS);
      ptx::tmem_relinquish_pair();
    } else {
      ptx::tmem_alloc(ptx::smem_u32(tmem_slot), C::TMEM_COLS);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync();   // the peer's barriers must be initialised before anything arrives on them
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  // Barriers, TMEM and tensor maps are set up: from here on global memory is read and written.  The trigger comes late
  // (producer thread, when it starts this CTA's LAST tile): a persistent GEMM lives for hundreds of microseconds, and
  // dependents that became resident at its start would sit on the SMs' thread slots all that time - measured: the
  // near-tie guard's short float64 kernels (other stream, highest priority) then wait a whole GEMM each for a slot
  // and the streaming end-to-end rate drops 6 %.
  pdl_wait();

  // tiles: CG == 1 -> 128-row blocks; CG == 2 -> 256-row blocks, rows [rank*128, +128) of each belong to this CTA
  const int num_m_units = (CG == 2) ? (g.num_m_blocks + 1) / 2 : g.num_m_blocks;
  const int num_tiles = num_m_units * g.num_n_blocks;
  const int num_kb = g.K / BK;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = unit; tile < num_tiles; tile += n_units) {
        const int m_blk = (tile % num_m_units) * CG + (int)rank, n_blk = tile / num_m_units;
        if (tile + n_units >= num_tiles) pdl_trigger();
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = base + stage * C::STAGE;
          if (CG == 2) {
            // both CTAs' loads are counted on the LEADER's full barrier: it expects two stages' worth of bytes
            if (rank == 0) ptx::mbar_expect_tx(full_bar(stage), 2 * C::STAGE);
            const int b_row = n_blk * BN + (int)rank * (BN / 2);
            ptx::tma_load_2d_pair(sa, &mAh, full_bar(stage), kb * BK, g.a_row0 + m_blk * BM);
            ptx::tma_load_2d_pair(sa + A_TILE, &mAl, full_bar(stage), kb * BK, g.a_row0 + m_blk * BM);
            ptx::tma_load_2d_pair(sa + 2 * A_TILE, &mBh, full_bar(stage), kb * BK, b_row);
            ptx::tma_load_2d_pair(sa + 2 * A_TILE + C::B_TILE, &mBl, full_bar(stage), kb * BK, b_row);
          } else {
            ptx::mbar_expect_tx(full_bar(stage), C::STAGE);
            ptx::tma_load_2d(sa, &mAh, full_bar(stage), kb * BK, g.a_row0 + m_blk * BM);
            ptx::tma_load_2d(sa + A_TILE, &mAl, full_bar(stage), kb * BK, g.a_row0 + m_blk * BM);
            ptx::tma_load_2d(sa + 2 * A_TILE, &mBh, full_bar(stage), kb * BK, n_blk * BN);
            ptx::tma_load_2d(sa + 2 * A_TILE + C::B_TILE, &mBl, full_bar(stage), kb * BK, n_blk * BN);
          }
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (rank == 0 && ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::umma_idesc_f16(BM * CG, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = unit; tile < num_tiles; tile += n_units) {
        ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        ptx::tc_fence_after();
        const int kchunk = ceil_div_dev(num_kb, KS);      // k-blocks per partial accumulator
        const int k_lim = g.K16 > 0 ? g.K16 : g.K;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          const uint32_t sa = base + stage * C::STAGE;
          const int part = (KS == 1) ? 0 : kb / kchunk;
          const bool fresh = (KS == 1) ? kb == 0 : kb % kchunk == 0;
          const uint32_t d_tmem = tmem_base + acc * C::ACC_COLS + part * BN;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            if (kb * BK + k * 16 >= k_lim) continue;      // zero padding: nothing to multiply
            const uint64_t ah = ptx::umma_desc_sw128(sa + k * 32);
            const uint64_t al = ptx::umma_desc_sw128(sa + A_TILE + k * 32);
            const uint64_t bh = ptx::umma_desc_sw128(sa + 2 * A_TILE + k * 32);
            const uint64_t bl = ptx::umma_desc_sw128(sa + 2 * A_TILE + C::B_TILE + k * 32);
            const uint32_t accum = (fresh && k == 0) ? 0u : 1u;
            if (CG == 2) {
              ptx::mma_f16_ss_pair(d_tmem, ah, bl, idesc, accum);
              ptx::mma_f16_ss_pair(d_tmem, al, bh, idesc, 1u);
              ptx::mma_f16_ss_pair(d_tmem, ah, bh, idesc, 1u);
            } else {
              ptx::mma_f16_ss(d_tmem, ah, bl, idesc, accum);
              ptx::mma_f16_ss(d_tmem, al, bh, idesc, 1u);
              ptx::mma_f16_ss(d_tmem, ah, bh, idesc, 1u);
            }
          }
          if (CG == 2) ptx::tc_commit_pair(empty_bar(stage), 3);   // frees the stage in both CTAs
          else ptx::tc_commit(empty_bar(stage));
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (CG == 2) ptx::tc_commit_pair(tfull_bar(acc), 3);
        else ptx::tc_commit(tfull_bar(acc));
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ===================== epilogue warps (2..) =====================
    constexpr int ET = 32 * EpiCfg<EPI>::WARPS;
    const int q = warp & 3;               // TMEM lane quarter this warp may read
    const int te = threadIdx.x - 64;      // 0..ET-1
    int acc = 0;
    uint32_t acc_phase = 0;
    // the tile's bias slice is fetched one tile ahead into registers: with short K (one or two k-blocks per
    // tile) the epilogue is the critical path and must not start every tile with a global-load round trip
    constexpr int BPT = (BN + ET - 1) / ET;
    float bnext[BPT];
    auto fetch_bias = [&](int tl) {
      const int nb = tl / num_m_units;
#pragma unroll
      for (int i = 0; i < BPT; ++i) {
        const int c = te + i * ET;
        const int n = nb * BN + c;
        float v = (EPI == EPI_LSE) ? -INFINITY : 0.f;
        if (c < BN && n < g.N) v = g.bias ? g.bias[n] : 0.f;
        bnext[i] = v;
      }
    };
    if (unit < num_tiles) fetch_bias(unit);
    for (int tile = unit; tile < num_tiles; tile += n_units) {
      const int m_blk = (tile % num_m_units) * CG + (int)rank, n_blk = tile / num_m_units;
      float* bs = bias_s + acc * BN;
#pragma unroll
      for (int i = 0; i < BPT; ++i) {
        const int c = te + i * ET;
        if (c < BN) bs[c] = bnext[i];
      }
      asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");
      if (tile + n_units < num_tiles) fetch_bias(tile + n_units);
      const uint32_t taddr = tmem_base + acc * C::ACC_COLS + (static_cast<uint32_t>(q * 32) << 16);
      const int row = m_blk * BM + q * 32 + lane;
      const bool row_ok = row < g.M;
      // LSTM: the parent's cell row does not depend on the MMAs - fetch it while they run
      constexpr int UPT = BN / 4;                       // hidden units per tile
      constexpr int UPW = UPT / (EpiCfg<EPI>::WARPS / 4);   // units per epilogue warp
      const int ubase = ((warp - 2) >> 2) * UPW;        // this warp's first unit inside the tile
      // The epilogue's global traffic goes through a per-warp shared-memory transpose: in the TMEM layout a lane
      // owns a ROW, so direct 16-byte accesses touch 32 different 128-byte lines per instruction, and ncu shows
      // those L1 tag cycles coming straight out of the MMA's shared-memory operand bandwidth.  Staged, one
      // instruction covers 8 rows x 64 contiguous bytes (lane -> row (lane & 7), 16-byte piece (lane >> 3)).
      float cprev[EPI == EPI_LSTM ? UPW : 1];
      uint8_t* stg = epi_stage + (EPI == EPI_LSTM ? (warp - 2) * EpiCfg<EPI>::STG_WARP : 0);
      constexpr int SR = EpiCfg<EPI>::STG_ROW;
      const int t_row = lane & 7, t_piece = lane >> 3;
      if (EPI == EPI_LSTM) {
        const int par = row_ok ? g.parent[row] : -1;
#pragma unroll
        for (int uc = 0; uc < UPW / 16; ++uc) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = i * 8 + t_row;
            const int p = __shfl_sync(0xffffffffu, par, rr);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p >= 0)
              v = *reinterpret_cast<const float4*>(g.c_src + (int64_t)p * g.ld_state + n_blk * UPT + ubase + uc * 16 +
                                                   4 * t_piece);
            *reinterpret_cast<float4*>(stg + rr * SR + 16 * t_piece) = v;
          }
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 t = *reinterpret_cast<const float4*>(stg + lane * SR + 16 * j);
            cprev[uc * 16 + 4 * j] = t.x;
            cprev[uc * 16 + 4 * j + 1] = t.y;
            cprev[uc * 16 + 4 * j + 2] = t.z;
            cprev[uc * 16 + 4 * j + 3] = t.w;
          }
          __syncwarp();
        }
      }
      ptx::mbar_wait(tfull_bar(acc), acc_phase);
      ptx::tc_fence_after();

      if (EPI == EPI_LSE) {
        float m_run = -INFINITY, c_run = -INFINITY, s_run = 0.f;
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t r[32];
          ptx::tmem_ld_x32(taddr + c0, r);
          ptx::tmem_ld_wait();
          float v[32];
          float cm = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = fmaf(__uint_as_float(r[j]), g.inv_scale, bs[c0 + j]);
            cm = fmaxf(cm, v[j]);
          }
          if (cm > -INFINITY) {
            const float m_new = fmaxf(m_run, cm);
            const float c_new = m_new * LOG2E;
            s_run *= exp2f(c_run - c_new);
            float part = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) part += exp2f(fmaf(v[j], LOG2E, -c_new));
            s_run += part;
            m_run = m_new;
            c_run = c_new;
          }
        }
        if (row_ok) g.part[(int64_t)row * g.part_ld + g.part_tile0 + n_blk] = make_float2(c_run, s_run);
      } else if (EPI == EPI_STORE) {
        const int n_part = (KS == 1) ? 1 : ceil_div_dev(num_kb, ceil_div_dev(num_kb, KS));   // partials the MMA warp filled
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t r[32];
          ptx::tmem_ld_x32(taddr + c0, r);
          ptx::tmem_ld_wait();
          // partial accumulators: added in fp32 with round-to-nearest (3 additions: ~1e-7 relative, a quarter of what
          // one accumulation chain carries; float64 here would put 128 F2F conversions per row on the quarter-rate pipe)
          float sum[KS > 1 ? 32 : 1];
          if (KS > 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[j] = __uint_as_float(r[j]);
            for (int pp = 1; pp < n_part; ++pp) {
              ptx::tmem_ld_x32(taddr + pp * BN + c0, r);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) sum[j] += __uint_as_float(r[j]);
            }
          }
          const int n0 = n_blk * BN + c0;
          if (row_ok && n0 < g.N) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j)
              v[j] = fmaf((KS > 1) ? sum[j] : __uint_as_float(r[j]), g.inv_scale, bs[c0 + j]);
            if (n0 + 32 <= g.N && (g.ldc & 3) == 0) {
              if (g.C32) {
                float4* dst = reinterpret_cast<float4*>(g.C32 + (int64_t)row * g.ldc + n0);
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              }
              if (g.s_hi) {
                split_store16(g.s_hi + (int64_t)row * g.lds + n0, g.s_lo + (int64_t)row * g.lds + n0, v, g.split_scale);
                split_store16(g.s_hi + (int64_t)row * g.lds + n0 + 16, g.s_lo + (int64_t)row * g.lds + n0 + 16, v + 16,
                              g.split_scale);
              }
            } else if (g.C32) {
              for (int j = 0; j < 32; ++j)
                if (n0 + j < g.N) g.C32[(int64_t)row * g.ldc + n0 + j] = v[j];
            }
          }
        }
      } else {
        // EPI_LSTM: tile columns are [i | f | o | g] x (BN/4) units (weights permuted on the host);
        // decoder/model.py:132-139
#pragma unroll
        for (int uc = 0; uc < UPW / 16; ++uc) {
          const int ul = ubase + uc * 16;               // first unit of the chunk inside the tile
          uint32_t ri[16], rf[16], ro[16], rg[16];
          ptx::tmem_ld_x16(taddr + ul, ri);
          ptx::tmem_ld_x16(taddr + UPT + ul, rf);
          ptx::tmem_ld_x16(taddr + 2 * UPT + ul, ro);
          ptx::tmem_ld_x16(taddr + 3 * UPT + ul, rg);
          ptx::tmem_ld_wait();
          float hv[16], cv[16];      // rows past M hold the zero-filled tile: computed, never stored
          {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float pi = fmaf(__uint_as_float(ri[j]), g.inv_scale, bs[ul + j]);
              const float pf = fmaf(__uint_as_float(rf[j]), g.inv_scale, bs[UPT + ul + j]);
              const float po = fmaf(__uint_as_float(ro[j]), g.inv_scale, bs[2 * UPT + ul + j]);
              const float pg = fmaf(__uint_as_float(rg[j]), g.inv_scale, bs[3 * UPT + ul + j]);
              const float c = fmaf(cprev[uc * 16 + j], fast_sigmoid(pf), fast_tanh(pg) * fast_sigmoid(pi));
              cv[j] = c;
              hv[j] = fast_tanh(c) * fast_sigmoid(po);
            }
          }
          // ---- staged stores: c, h (fp32, 64 B per row) then the fp16 split of h (hi 32 B | lo 32 B per row) ----
          const int u0 = n_blk * UPT + ul;
          const int row_base = m_blk * BM + q * 32;
          {
            float* const outs[2] = {g.c_out, g.h_out};
#pragma unroll
            for (int which = 0; which < 2; ++which) {
              if (which == 1 && !g.h_out) continue;      // tied models keep h only as its fp16 split
              const float* src = which ? hv : cv;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<float4*>(stg + lane * SR + 16 * j) =
                    make_float4(src[4 * j], src[4 * j + 1], src[4 * j + 2], src[4 * j + 3]);
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int rr = i * 8 + t_row;
                const float4 v = *reinterpret_cast<const float4*>(stg + rr * SR + 16 * t_piece);
                if (row_base + rr < g.M)
                  *reinterpret_cast<float4*>(outs[which] + (int64_t)(row_base + rr) * g.ld_state + u0 + 4 * t_piece) = v;
              }
              __syncwarp();
            }
          }
          {
            uint32_t ph[8], pl[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float a = hv[2 * j] * g.split_scale, b2_ = hv[2 * j + 1] * g.split_scale;
              const __half2 h2 = __floats2half2_rn(a, b2_);
              const float2 hf = __half22float2(h2);
              const __half2 l2 = __floats2half2_rn(a - hf.x, b2_ - hf.y);
              ph[j] = *reinterpret_cast<const uint32_t*>(&h2);
              pl[j] = *reinterpret_cast<const uint32_t*>(&l2);
            }
            uint4* w4 = reinterpret_cast<uint4*>(stg + lane * SR);
            w4[0] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
            w4[1] = make_uint4(ph[4], ph[5], ph[6], ph[7]);
            w4[2] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
            w4[3] = make_uint4(pl[4], pl[5], pl[6], pl[7]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rr = i * 8 + t_row;
              const uint4 v = *reinterpret_cast<const uint4*>(stg + rr * SR + 16 * t_piece);
              __half* dst = (t_piece < 2 ? g.s_hi : g.s_lo) + (int64_t)(row_base + rr) * g.lds + u0 + 8 * (t_piece & 1);
              if (row_base + rr < g.M) *reinterpret_cast<uint4*>(dst) = v;
            }
            __syncwarp();
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) ptx::mbar_arrive_cluster(tempty_bar(acc), 0);   // the leader's barrier gates the next MMAs
        else ptx::mbar_arrive(tempty_bar(acc));
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }
  ptx::tc_fence_before();
  if (CG == 2) {
    ptx::cluster_sync();   // no CTA of the pair may exit while the other can still signal it or read its smem
    if (warp == 1) ptx::tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
  } else {
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// Row-stationary output GEMM + online log-sum-exp for SHORT K (D-softmax* tail segments, K = 64 / 128).
//
// With K = 64 a 256 x 256 tile is 12 MMAs (1536 cycles) but needs 128 KB of operands: fed tile by tile
// (k_tc_gemm) the pair asks L2 for 64 B/clk, the chip for ~8 TB/s - the segment is L2-bound, not tensor-bound.
// Here a CTA pair owns a contiguous range of tiles in (row block, column block) order, keeps the A operand of
// its current 256-row block RESIDENT in shared memory (32 KB per k-block per CTA) and streams only the weight
// tiles through the ring, so the operand traffic per tile halves.  The epilogue warps own fixed rows, so the
// running (max, sum exp) of a row also stays in registers across the column blocks of the run and one partial
// per (row, run) is written instead of one per (row, tile): 392 -> 2..3 partials per row at V = 100k.
// K is issued in 16-wide steps up to round_up(width, 16): the zero padding up to 64 is never multiplied.
// Measured and rejected (profiles/r02/README.md): with one exponential in four on the FMA pipe the cfg-5 tail kernel runs
// 0.631 ms per launch against 0.604 ms with all of them on the MUFU (cfg 3: 0.081 vs 0.076) - twelve issue slots per
// polynomial cost more than the MUFU queue gives back.  Build with -DJLM_RS_POLY=n to try other ratios.
#ifndef JLM_RS_POLY
#define JLM_RS_POLY 0
#endif
constexpr int RS_POLY = JLM_RS_POLY;      // 0: every exponential on the MUFU; n: one in n on the FMA pipe

template <int KB>
struct RsCfg {
  static constexpr int BN = 256;
  static constexpr int A_BYTES = KB * 2 * A_TILE;                 // hi + lo, 128 rows x 64 per k-block
  static constexpr int B_HALF = (BN / 2) * BK * 2;                // one k-block of this CTA's half of the N tile
  // the weight ring holds single k-blocks (hi + lo, 32 KB): four of them beside a one- or two-block A operand; with
  // K = 256 (A = 128 KB) three still fit, 3/4 of a tile ahead of the MMAs - the depth k_tc_gemm's pair ring has
  static constexpr int B_STAGE = 2 * B_HALF;
  static constexpr int STAGES = (KB <= 2) ? 4 : 3;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM = A_BYTES + STAGES * B_STAGE + BAR_BYTES + 1024;
  static_assert(SMEM <= 227 * 1024, "row-stationary kernel: A operand + weight ring exceed the shared memory of an SM");
  // Sixteen epilogue warps, four per TMEM lane quarter (each takes a quarter of the tile's columns): with one exp per
  // logit the epilogue paces the short-K tiles, and it is latency-bound, not pipe-bound - ncu on the 8-warp version
  // (profiles/r02/ncu_rs_cfg5_v2.csv): XU pipe 61 %, issue slots 45 %, tensor pipe 42 %, the stalls are fixed-latency
  // waits and MUFU / TMEM-load scoreboards of the two warps a scheduler had.
  static constexpr int EPI_WARPS = 16;      // also for K = 128 (measured with eight: 101 vs 94 us per launch at cfg 3)
  static constexpr int COL_PARTS = EPI_WARPS / 4;            // column slices of a tile, one partial each
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
};

struct RsArgs {
  int M, N, K16;            // K16: K rounded up to the MMA's 16
  int num_m_units, num_n_blocks;
  int a_row0;
  float inv_scale;
  const float* bias;
  float2* part;
  int part_ld, part_col0, n_slots;
};

template <int KB>
__global__ void __launch_bounds__(RsCfg<KB>::THREADS, 1)
k_tc_lse_rs(const __grid_constant__ CUtensorMap mAh, const __grid_constant__ CUtensorMap mAl,
            const __grid_constant__ CUtensorMap mBh, const __grid_constant__ CUtensorMap mBl, const RsArgs g) {
  using C = RsCfg<KB>;
  constexpr int BN = C::BN;
  const uint32_t rank = ptx::cluster_ctarank();
  const int unit = (int)(blockIdx.x >> 1), n_units = (int)(gridDim.x >> 1);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t a_base = base;
  const uint32_t b_base = base + C::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen + C::A_BYTES + C::STAGES * C::B_STAGE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 6);
  const uint32_t bar0 = ptx::smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * C::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * C::STAGES + 2 + a); };
  const uint32_t afull_bar = bar0 + 8u * (2 * C::STAGES + 4), aempty_bar = bar0 + 8u * (2 * C::STAGES + 5);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tensormap(&mAh);
    ptx::prefetch_tensormap(&mAl);
    ptx::prefetch_tensormap(&mBh);
    ptx::prefetch_tensormap(&mBl);
  }
  if (warp == 1) {
    if (ptx::elect_one()) {
      for (int s = 0; s < C::STAGES; ++s) {
        ptx::mbar_init(full_bar(s), 1);
        ptx::mbar_init(empty_bar(s), 1);
      }
      for (int a = 0; a < 2; ++a) {
        ptx::mbar_init(tfull_bar(a), 1);
        ptx::mbar_init(tempty_bar(a), 2 * C::EPI_WARPS);      // both CTAs' epilogue warps arrive on the leader's
      }
      ptx::mbar_init(afull_bar, 1);
      ptx::mbar_init(aempty_bar, 1);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc_pair(ptx::smem_u32(tmem_slot), 2 * BN);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  pdl_wait();      // trigger: producer thread at the pair's last tile (see k_tc_gemm)

  // this pair's contiguous tile range, tile id = m_unit * NT + n_blk
  const int NT = g.num_n_blocks;
  const int64_t T = (int64_t)g.num_m_units * NT;
  const int t_lo = (int)(T * unit / n_units), t_hi = (int)(T * (unit + 1) / n_units);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (ptx::elect_one()) {
      int stage = 0, cur_m = -1;
      uint32_t phase = 0, a_phase = 0;
      for (int tile = t_lo; tile < t_hi; ++tile) {
        const int m_unit = tile / NT, n_blk = tile - m_unit * NT;
        if (tile + 1 == t_hi) pdl_trigger();
        if (m_unit != cur_m) {      // new row block: wait until the MMAs of the previous one have read A
          ptx::mbar_wait(aempty_bar, a_phase ^ 1u);
          a_phase ^= 1u;
          if (rank == 0) ptx::mbar_expect_tx(afull_bar, 2 * C::A_BYTES);
          const int row = g.a_row0 + (m_unit * 2 + (int)rank) * BM;
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            ptx::tma_load_2d_pair(a_base + kb * 2 * A_TILE, &mAh, afull_bar, kb * BK, row);
            ptx::tma_load_2d_pair(a_base + kb * 2 * A_TILE + A_TILE, &mAl, afull_bar, kb * BK, row);
          }
          cur_m = m_unit;
        }
        const int b_row = n_blk * BN + (int)rank * (BN / 2);
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          if (kb * BK >= g.K16) break;                      // a k-block of nothing but zero padding is not fetched
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          if (rank == 0) ptx::mbar_expect_tx(full_bar(stage), 2 * C::B_STAGE);
          const uint32_t sb = b_base + stage * C::B_STAGE;
          ptx::tma_load_2d_pair(sb, &mBh, full_bar(stage), kb * BK, b_row);
          ptx::tma_load_2d_pair(sb + C::B_HALF, &mBl, full_bar(stage), kb * BK, b_row);
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    if (rank == 0 && ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::umma_idesc_f16(2 * BM, BN);
      int stage = 0, acc = 0, cur_m = -1;
      uint32_t phase = 0, acc_phase = 0, a_phase = 0;
      for (int tile = t_lo; tile < t_hi; ++tile) {
        const int m_unit = tile / NT;
        if (m_unit != cur_m) {
          ptx::mbar_wait(afull_bar, a_phase);
          a_phase ^= 1u;
          cur_m = m_unit;
        }
        ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        const uint32_t d_tmem = tmem_base + acc * BN;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          if (kb * BK >= g.K16) break;
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          const uint32_t sb = b_base + stage * C::B_STAGE;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            if (kb * BK + k * 16 < g.K16) {
              const uint64_t ah = ptx::umma_desc_sw128(a_base + kb * 2 * A_TILE + k * 32);
              const uint64_t al = ptx::umma_desc_sw128(a_base + kb * 2 * A_TILE + A_TILE + k * 32);
              const uint64_t bh = ptx::umma_desc_sw128(sb + k * 32);
              const uint64_t bl = ptx::umma_desc_sw128(sb + C::B_HALF + k * 32);
              ptx::mma_f16_ss_pair(d_tmem, ah, bl, idesc, (kb | k) != 0 ? 1u : 0u);
              ptx::mma_f16_ss_pair(d_tmem, al, bh, idesc, 1u);
              ptx::mma_f16_ss_pair(d_tmem, ah, bh, idesc, 1u);
            }
          }
          ptx::tc_commit_pair(empty_bar(stage), 3);      // this k-block's weight stage is free once its MMAs retire
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        // last tile of this row block in our range: once its MMAs are done the A operand may be overwritten
        if (tile + 1 == t_hi || (tile + 1) / NT != m_unit) ptx::tc_commit_pair(aempty_bar, 3);
        ptx::tc_commit_pair(tfull_bar(acc), 3);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ===================== epilogue warps: online (max, sum exp) per row, carried across the run =====================
    constexpr int HALF = BN / C::COL_PARTS;    // columns per warp
    const int q = warp & 3;                    // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;          // which slice of the tile's columns
    int acc = 0;
    uint32_t acc_phase = 0;
    // The bias slice of a chunk is read straight from global memory (every lane the same address: one broadcast
    // sector per load, L1-resident after the first warp) - no shared-memory staging and, above all, no per-tile
    // barrier keeping the sixteen warps in lock-step: their FMA/max phases and MUFU phases now overlap.
    const bool bias_vec = g.bias && ((reinterpret_cast<uintptr_t>(g.bias) & 15) == 0);
    float m_run = -INFINITY, c_run = -INFINITY, s_run = 0.f;
    int m_unit = t_lo / NT, n_blk = t_lo - m_unit * NT;      // kept incrementally: no division per tile
    for (int tile = t_lo; tile < t_hi; ++tile) {
      const uint32_t taddr = tmem_base + acc * BN + half * HALF + (static_cast<uint32_t>(q * 32) << 16);
      const int col0 = n_blk * BN + half * HALF;             // first column of this warp's slice
      ptx::mbar_wait(tfull_bar(acc), acc_phase);
      ptx::tc_fence_after();
      // TMEM loads are issued one 32-column chunk ahead of the math
      uint32_t r[2][32];
      ptx::tmem_ld_x32(taddr, r[0]);
#pragma unroll
      for (int ch = 0; ch < HALF / 32; ++ch) {
        float bias_r[32];
        const int cb = col0 + ch * 32;
        if (bias_vec && cb + 32 <= g.N) {
          const float4* bp = reinterpret_cast<const float4*>(g.bias + cb);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 t4 = __ldg(bp + i);
            bias_r[4 * i] = t4.x;
            bias_r[4 * i + 1] = t4.y;
            bias_r[4 * i + 2] = t4.z;
            bias_r[4 * i + 3] = t4.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) bias_r[j] = (cb + j < g.N) ? (g.bias ? __ldg(g.bias + cb + j) : 0.f) : -INFINITY;
        }
        ptx::tmem_ld_wait();
        if (ch + 1 < HALF / 32) ptx::tmem_ld_x32(taddr + (ch + 1) * 32, r[(ch + 1) & 1]);
        uint32_t(&rc)[32] = r[ch & 1];
        float v[32];
        float cm = -INFINITY;
        // The exponentials are taken against the row's RUNNING maximum (from the chunks before), so the MUFUs can be
        // issued as soon as a logit is formed instead of after the chunk's own maximum is known; the sum is rescaled
        // afterwards in the (rare, after the first chunks) case that this chunk raised the maximum.  ex2.approx.ftz:
        // 2^-inf = 0, arguments below -126 flush to 0.  A first chunk, or a jump of more than 60 (fp32 range), takes
        // the two-pass form.
        float part[4] = {0.f, 0.f, 0.f, 0.f};      // four independent chains
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = fmaf(__uint_as_float(rc[j]), g.inv_scale, bias_r[j]);
          cm = fmaxf(cm, v[j]);
          const float arg = fmaf(v[j], LOG2E, -c_run);
          // every RS_POLY-th exponential goes to the FMA pipe: the MUFU queue paces this loop (ncu: stall_mio on
          // every MUFU.EX2, issue slots 45 % busy)
          part[j & 3] += (RS_POLY > 0 && (j % (RS_POLY > 0 ? RS_POLY : 1)) == RS_POLY - 1) ? ex2_poly(arg) : ex2_approx(arg);
        }
        if (cm > m_run) {
          const float c_new = cm * LOG2E;
          if (m_run > -INFINITY && cm - m_run < 60.f) {
            s_run = (s_run + (part[0] + part[1]) + (part[2] + part[3])) * ex2_approx(c_run - c_new);
          } else {
            s_run *= ex2_approx(c_run - c_new);
            float p2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 32; ++j) p2[j & 3] += ex2_approx(fmaf(v[j], LOG2E, -c_new));
            s_run += (p2[0] + p2[1]) + (p2[2] + p2[3]);
          }
          m_run = cm;
          c_run = c_new;
        } else if (m_run > -INFINITY) {      // (no maximum yet and nothing but masked columns: part holds NaNs, skip)
          s_run += (part[0] + part[1]) + (part[2] + part[3]);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(tempty_bar(acc), 0);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
      const bool block_done = n_blk + 1 == NT;             // this tile closes the row block
      if (tile + 1 == t_hi || block_done) {
        // end of this pair's run over the row block: one partial per (row, column half).  Slot = how many pairs before
        // this one also work on the block; the pair that finishes the block neutralises the slots nobody writes.
        const int row = (m_unit * 2 + (int)rank) * BM + q * 32 + lane;
        const int64_t t0 = (int64_t)m_unit * NT;
        const int u_first = (int)(((t0 + 1) * n_units + T - 1) / T) - 1;
        const int slot = unit - u_first;
        if (row < g.M) {
          float2* p = g.part + (int64_t)row * g.part_ld + g.part_col0;
          if (slot < g.n_slots) p[C::COL_PARTS * slot + half] = make_float2(c_run, s_run);
          if (block_done)
            for (int k = slot + 1; k < g.n_slots; ++k) p[C::COL_PARTS * k + half] = make_float2(-INFINITY, 0.f);
        }
        m_run = -INFINITY;
        c_run = -INFINITY;
        s_run = 0.f;
      }
      if (block_done) {
        n_blk = 0;
        ++m_unit;
      } else {
        ++n_blk;
      }
    }
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (warp == 1) ptx::tmem_dealloc_pair(tmem_base, 2 * BN);
}

// A = [ h[parent] | LM_in[word] ] as fp16 hi/lo at the gate-input scale: pure 16-byte row copies - the LSTM
// epilogue leaves h as its split per slot and the embedding table is split once at load time (same formula, same
// bits as splitting on the fly).
__global__ void k_tc_gather_rows(const __half* __restrict__ H_hi, const __half* __restrict__ H_lo,
                                 const int32_t* __restrict__ parent, const int32_t* __restrict__ word,
                                 const __half* __restrict__ E_hi, const __half* __restrict__ E_lo, int Hp, int Ep, int M,
                                 __half* __restrict__ A_hi, __half* __restrict__ A_lo) {
  pdl_enter();
  const int Kg = Hp + Ep;
  const int64_t total = (int64_t)M * (Kg / 8);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / (Kg / 8));
    const int k = (int)(i % (Kg / 8)) * 8;
    uint4 vh = make_uint4(0, 0, 0, 0), vl = vh;
    if (k < Hp) {
      const int p = parent[m];
      if (p >= 0) {
        vh = *reinterpret_cast<const uint4*>(H_hi + (int64_t)p * Hp + k);
        vl = *reinterpret_cast<const uint4*>(H_lo + (int64_t)p * Hp + k);
      }
    } else {
      const int64_t o = (int64_t)word[m] * Ep + (k - Hp);
      vh = *reinterpret_cast<const uint4*>(E_hi + o);
      vl = *reinterpret_cast<const uint4*>(E_lo + o);
    }
    *reinterpret_cast<uint4*>(A_hi + (int64_t)m * Kg + k) = vh;
    *reinterpret_cast<uint4*>(A_lo + (int64_t)m * Kg + k) = vl;
  }
}

// Vocabulary-selection logits on the tensor cores (static vocab_select / DynamicDecoder, decoder.py:137-151,
// decoder_dynamic.py:93-148): per sentence, y[r][j] = T[r] . LM[cols[j]] + b2[cols[j]] for its own word list.
// The word rows are a GATHER, so there is no TMA tile: the CTA's threads copy 128 gathered rows of the
// fp16 hi/lo weight copies (A operand, M = 128 vocabulary entries) and the sentence's <= 32 stage-1 rows
// (B operand, N = 32) into shared memory in the SWIZZLE_128B layout the UMMA descriptors expect
// (row r at (r/8)*1024 + (r%8)*128, 16-byte chunk c at position c ^ (r%8)), make the writes visible to the
// async proxy, and one thread issues the 3-MMA split products (128 x 32 x 16, fp32 accumulators in TMEM).
// K is consumed in halves of 128 so a CTA needs 80 KB and two CTAs share an SM: one gathers while the other
// multiplies.  TMEM lane = vocabulary entry, column = sentence row, so the epilogue's stores are coalesced
// along the word list.  Replaces the float64 CUDA-core k_vocab_logits (414 us per frame at cfg 4) on this
// back end; the needed-word logits are still recomputed in float64 by k_score_nodes.
constexpr int VT_M = 128;                 // vocabulary entries per CTA
constexpr int VT_N = 32;                  // sentence rows (beam) per CTA, zero padded
constexpr int VT_KH = 128;                // K consumed per pass
constexpr int VT_A_TILE = VT_M * 128;     // bytes of one 64-wide k-block of A (hi or lo)
constexpr int VT_B_TILE = VT_N * 128;
constexpr int VT_SMEM = (VT_KH / BK) * (2 * VT_A_TILE + 2 * VT_B_TILE) + 1024 + 64;

__global__ void __launch_bounds__(128, 2)
k_tc_vocab_logits(const __half* __restrict__ W_hi, const __half* __restrict__ W_lo, int64_t ldw,
                  const __half* __restrict__ T_hi, const __half* __restrict__ T_lo, int64_t ldt, int K,
                  const SubsetJob* __restrict__ jobs, const int32_t* __restrict__ cols, const float* __restrict__ b2,
                  float inv_scale, double* __restrict__ out, int col_skip) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ int32_t wid[VT_M];
  __shared__ uint64_t bar_s;
  __shared__ uint32_t tmem_slot;
  const SubsetJob job = jobs[blockIdx.y];
  const int c0 = col_skip + blockIdx.x * VT_M;      // the first col_skip words are shared by all sentences: dense GEMM
  if (c0 >= job.ncols) {
    pdl_wait();
    return;
  }
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bar = ptx::smem_u32(&bar_s);
  wid[tid] = (c0 + tid < job.ncols) ? cols[job.col0 + c0 + tid] : -1;
  if (warp == 0) {
    if (ptx::elect_one()) {
      ptx::mbar_init(bar, 1);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), VT_N);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);
  pdl_enter();      // (jobs / cols above are plan data, uploaded before the batch's first kernel)
  constexpr int KBH = VT_KH / BK;                 // k-blocks per pass
  constexpr int PASS = 2 * VT_A_TILE + 2 * VT_B_TILE;
  uint32_t phase = 0;
  for (int kh = 0; kh < K; kh += VT_KH) {
    // ---- gather this K half: A rows by word id, B rows from the sentence's stage-1 block ----
#pragma unroll
    for (int kb = 0; kb < KBH; ++kb) {
      uint8_t* a_hi = gen + kb * PASS;
      uint8_t* a_lo = a_hi + VT_A_TILE;
      uint8_t* b_hi = a_lo + VT_A_TILE;
      uint8_t* b_lo = b_hi + VT_B_TILE;
      const int kcol = kh + kb * BK;
#pragma unroll
      for (int i = tid; i < VT_M * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        const int w = wid[r];
        uint4 vh = make_uint4(0, 0, 0, 0), vl = vh;
        if (w >= 0 && kcol + c * 8 < K) {
          vh = *reinterpret_cast<const uint4*>(W_hi + (int64_t)w * ldw + kcol + c * 8);
          vl = *reinterpret_cast<const uint4*>(W_lo + (int64_t)w * ldw + kcol + c * 8);
        }
        const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(a_hi + off) = vh;
        *reinterpret_cast<uint4*>(a_lo + off) = vl;
      }
#pragma unroll
      for (int i = tid; i < VT_N * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        uint4 vh = make_uint4(0, 0, 0, 0), vl = vh;
        if (r < job.rows && kcol + c * 8 < K) {
          vh = *reinterpret_cast<const uint4*>(T_hi + (job.row0 + r) * ldt + kcol + c * 8);
          vl = *reinterpret_cast<const uint4*>(T_lo + (job.row0 + r) * ldt + kcol + c * 8);
        }
        const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(b_hi + off) = vh;
        *reinterpret_cast<uint4*>(b_lo + off) = vl;
      }
    }
    ptx::fence_proxy_async();          // generic-proxy stores -> visible to the tensor core's async proxy
    __syncthreads();
    if (warp == 0 && ptx::elect_one()) {
      ptx::tc_fence_after();
      constexpr uint32_t idesc = ptx::umma_idesc_f16(VT_M, VT_N);
#pragma unroll
      for (int kb = 0; kb < KBH; ++kb) {
        const uint32_t sa = base + kb * PASS;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t ah = ptx::umma_desc_sw128(sa + k * 32);
          const uint64_t al = ptx::umma_desc_sw128(sa + VT_A_TILE + k * 32);
          const uint64_t bh = ptx::umma_desc_sw128(sa + 2 * VT_A_TILE + k * 32);
          const uint64_t bl = ptx::umma_desc_sw128(sa + 2 * VT_A_TILE + VT_B_TILE + k * 32);
          ptx::mma_f16_ss(tmem_base, ah, bl, idesc, (kh | kb | k) != 0 ? 1u : 0u);
          ptx::mma_f16_ss(tmem_base, al, bh, idesc, 1u);
          ptx::mma_f16_ss(tmem_base, ah, bh, idesc, 1u);
        }
      }
      ptx::tc_commit(bar);             // arrives when every MMA above has read its operands and written TMEM
    }
    __syncwarp();
    ptx::mbar_wait(bar, phase);        // shared memory may be overwritten / accumulators may be read
    phase ^= 1u;
    ptx::tc_fence_after();
  }
  // ---- epilogue: lane = vocabulary entry, columns = sentence rows ----
  {
    uint32_t r[32];
    ptx::tmem_ld_x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16), r);
    ptx::tmem_ld_wait();
    const int j = c0 + tid;
    if (j < job.ncols) {
      const double bias = (double)b2[wid[tid]];
      for (int q = 0; q < job.rows; ++q)
        out[job.out0 + (int64_t)q * job.ncols + j] = (double)(__uint_as_float(r[q]) * inv_scale) + bias;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, VT_N);
}

// The same product with the word rows of EVERY sentence's list gathered ONCE per batch into a dense operand
// (k_tc_gather_shared over jlm_batch's vocab_cols: row = position in the concatenated lists) instead of once per
// frame: a sentence's list does not change from frame to frame (DynamicDecoder cuts the running log-sum-exp of the
// full list at the per-frame boundaries), so the per-frame gather - 128 threads copying 16-byte pieces through the
// LSU, 320 MB per frame at cfg 4 - was the same rows twenty-five times.  Here one thread issues TMA boxes for the
// sentence's 128 word rows (A) and its <= 32 stage-1 rows (B); a tile that runs past the end of the list just reads
// the next sentence's rows (lanes beyond ncols are not stored).
__global__ void __launch_bounds__(128, 2)
k_tc_vocab_dense(const __grid_constant__ CUtensorMap mWh, const __grid_constant__ CUtensorMap mWl,
                 const __grid_constant__ CUtensorMap mTh, const __grid_constant__ CUtensorMap mTl, int K,
                 const SubsetJob* __restrict__ jobs, const float* __restrict__ wd_bias, float inv_scale,
                 double* __restrict__ out, int col_skip) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_s[2];      // [0] TMA bytes of a pass, [1] its MMAs retired
  __shared__ uint32_t tmem_slot;
  const SubsetJob job = jobs[blockIdx.y];
  const int c0 = col_skip + blockIdx.x * VT_M;
  if (c0 >= job.ncols) {
    pdl_wait();
    return;
  }
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bar_tma = ptx::smem_u32(&bar_s[0]), bar_mma = ptx::smem_u32(&bar_s[1]);
  if (warp == 0) {
    if (ptx::elect_one()) {
      ptx::prefetch_tensormap(&mWh);
      ptx::prefetch_tensormap(&mWl);
      ptx::prefetch_tensormap(&mTh);
      ptx::prefetch_tensormap(&mTl);
      ptx::mbar_init(bar_tma, 1);
      ptx::mbar_init(bar_mma, 1);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), VT_N);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);
  pdl_enter();
  constexpr int KBH = VT_KH / BK;                 // k-blocks per pass
  constexpr int PASS = 2 * VT_A_TILE + 2 * VT_B_TILE;
  uint32_t phase = 0;
  for (int kh = 0; kh < K; kh += VT_KH) {
    if (tid == 0) {
      ptx::mbar_expect_tx(bar_tma, KBH * PASS);
#pragma unroll
      for (int kb = 0; kb < KBH; ++kb) {
        const uint32_t sa = base + kb * PASS;
        const int kcol = kh + kb * BK;            // a k-block past K is out of bounds: zero filled, bytes still counted
        ptx::tma_load_2d(sa, &mWh, bar_tma, kcol, (int32_t)(job.col0 + c0));
        ptx::tma_load_2d(sa + VT_A_TILE, &mWl, bar_tma, kcol, (int32_t)(job.col0 + c0));
        ptx::tma_load_2d(sa + 2 * VT_A_TILE, &mTh, bar_tma, kcol, (int32_t)job.row0);
        ptx::tma_load_2d(sa + 2 * VT_A_TILE + VT_B_TILE, &mTl, bar_tma, kcol, (int32_t)job.row0);
      }
    }
    ptx::mbar_wait(bar_tma, phase);
    if (warp == 0 && ptx::elect_one()) {
      ptx::tc_fence_after();
      constexpr uint32_t idesc = ptx::umma_idesc_f16(VT_M, VT_N);
#pragma unroll
      for (int kb = 0; kb < KBH; ++kb) {
        const uint32_t sa = base + kb * PASS;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t ah = ptx::umma_desc_sw128(sa + k * 32);
          const uint64_t al = ptx::umma_desc_sw128(sa + VT_A_TILE + k * 32);
          const uint64_t bh = ptx::umma_desc_sw128(sa + 2 * VT_A_TILE + k * 32);
          const uint64_t bl = ptx::umma_desc_sw128(sa + 2 * VT_A_TILE + VT_B_TILE + k * 32);
          ptx::mma_f16_ss(tmem_base, ah, bl, idesc, (kh | kb | k) != 0 ? 1u : 0u);
          ptx::mma_f16_ss(tmem_base, al, bh, idesc, 1u);
          ptx::mma_f16_ss(tmem_base, ah, bh, idesc, 1u);
        }
      }
      ptx::tc_commit(bar_mma);
    }
    __syncwarp();
    ptx::mbar_wait(bar_mma, phase);    // shared memory may be overwritten / accumulators may be read
    phase ^= 1u;
    ptx::tc_fence_after();
  }
  {
    uint32_t r[32];
    ptx::tmem_ld_x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16), r);
    ptx::tmem_ld_wait();
    const int j = c0 + tid;
    if (j < job.ncols) {
      const double bias = (double)wd_bias[job.col0 + j];
      for (int q = 0; q < job.rows; ++q)
        out[job.out0 + (int64_t)q * job.ncols + j] = (double)(__uint_as_float(r[q]) * inv_scale) + bias;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, VT_N);
}

// rows ids[0..n) of the split weight block and their biases -> dense operand (zero rows up to n_pad)
__global__ void k_tc_gather_shared(const __half* __restrict__ W_hi, const __half* __restrict__ W_lo, int64_t ldw,
                                   const float* __restrict__ b2, const int32_t* __restrict__ ids, int n, int n_pad,
                                   __half* __restrict__ S_hi, __half* __restrict__ S_lo, float* __restrict__ bias) {
  pdl_enter();
  const int64_t total = (int64_t)n_pad * (ldw / 8);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / (ldw / 8)), k = (int)(i % (ldw / 8)) * 8;
    uint4 vh = make_uint4(0, 0, 0, 0), vl = vh;
    if (r < n) {
      const int64_t o = (int64_t)ids[r] * ldw + k;
      vh = *reinterpret_cast<const uint4*>(W_hi + o);
      vl = *reinterpret_cast<const uint4*>(W_lo + o);
    }
    *reinterpret_cast<uint4*>(S_hi + (int64_t)r * ldw + k) = vh;
    *reinterpret_cast<uint4*>(S_lo + (int64_t)r * ldw + k) = vl;
    if (k == 0) bias[r] = r < n ? b2[ids[r]] : 0.f;
  }
}

// partial (c = max*log2e, s = sum 2^(v*log2e - c)) per 256-column tile (k_tc_gemm) or per run (k_tc_lse_rs) ->
// natural-log LSE in float64.  Each output segment owns a fixed column range of the partial array and says how many
// of its columns this step's launch wrote.
struct LseCols {
  int nseg;
  int col0[JLM_MAX_SEGMENTS], cnt[JLM_MAX_SEGMENTS];
};

__global__ void k_tc_lse_merge(const float2* __restrict__ part, int part_ld, LseCols cols, int M,
                               double* __restrict__ lse) {
  pdl_enter();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= M) return;
  const float2* p = part + (int64_t)warp * part_ld;
  // one pass: the partials of a row (<= 8 per lane for V <= 65536) stay in registers between the max and the sum
  constexpr int PL = 8;
  float2 v[PL];
  float mx = -INFINITY;
  int held = 0;          // partials this lane keeps in registers (the first PL it meets)
  for (int sg = 0; sg < cols.nseg; ++sg) {
    const float2* ps = p + cols.col0[sg];
    for (int t = lane; t < cols.cnt[sg]; t += 32) {
      const float2 w = ps[t];
      mx = fmaxf(mx, w.x);
#pragma unroll
      for (int i = 0; i < PL; ++i)
        if (i == held) v[i] = w;
      ++held;
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  // 2^(c_tile - c_max) in fp32 (the tile sums themselves are fp32); products and the sum in float64
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < PL; ++i)
    if (i < held && v[i].x > -INFINITY) s += (double)v[i].y * (double)exp2f(v[i].x - mx);
  if (held > PL) {       // very wide vocabularies: the rest is read again
    int seen = 0;
    for (int sg = 0; sg < cols.nseg; ++sg) {
      const float2* ps = p + cols.col0[sg];
      for (int t = lane; t < cols.cnt[sg]; t += 32) {
        if (seen >= PL) {
          const float2 w = ps[t];
          if (w.x > -INFINITY) s += (double)w.y * (double)exp2f(w.x - mx);
        }
        ++seen;
      }
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) lse[warp] = 0.6931471805599453094 * ((double)mx + log2(s));
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp16 K-major operand [rows, K] with row pitch `pitch_elems`; box = 64 (K) x box_rows, SWIZZLE_128B
int32_t make_map(CUtensorMap* m, const __half* base, int64_t K, int64_t rows, int64_t pitch_elems, int box_rows) {
  EncodeTiledFn enc = get_encode();
  JLM_REQUIRE(enc, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)pitch_elems * sizeof(__half)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  JLM_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) K=%lld rows=%lld pitch=%lld box_rows=%d", (int)r,
              (long long)K, (long long)rows, (long long)pitch_elems, box_rows);
  return 0;
}

float pow2_scale_for(double max_abs) {
  if (!(max_abs > 0)) return 1.f;
  int e = (int)std::floor(std::log2(32768.0 / max_abs));
  if (e > 24) e = 24;
  if (e < -24) e = -24;
  return std::ldexp(1.f, e);
}

// split a host matrix (float64 values) [rows, K] into device fp16 hi/lo
int32_t upload_split(TcOperand* op, const std::vector<double>& w, int64_t rows, int64_t K, int box_rows,
                     float fixed_scale = 0.f) {
  double mx = 0;
  for (double v : w) mx = std::max(mx, std::fabs(v));
  op->scale = fixed_scale > 0.f ? fixed_scale : pow2_scale_for(mx);
  op->rows = rows;
  op->K = K;
  std::vector<__half> hi(w.size()), lo(w.size());
  for (size_t i = 0; i < w.size(); ++i) {
    const double x = w[i] * (double)op->scale;
    const __half a = __float2half_rn((float)x);
    hi[i] = a;
    lo[i] = __float2half_rn((float)(x - (double)__half2float(a)));
  }
  JLM_CUDA(cudaMalloc(&op->hi, w.size() * sizeof(__half)));
  JLM_CUDA(cudaMalloc(&op->lo, w.size() * sizeof(__half)));
  JLM_CUDA(cudaMemcpy(op->hi, hi.data(), w.size() * sizeof(__half), cudaMemcpyHostToDevice));
  JLM_CUDA(cudaMemcpy(op->lo, lo.data(), w.size() * sizeof(__half), cudaMemcpyHostToDevice));
  JLM_TRY(make_map(&op->map_hi, op->hi, K, rows, K, box_rows));
  JLM_TRY(make_map(&op->map_lo, op->lo, K, rows, K, box_rows));
  JLM_TRY(make_map(&op->pair_hi, op->hi, K, rows, K, box_rows / 2));
  JLM_TRY(make_map(&op->pair_lo, op->lo, K, rows, K, box_rows / 2));
  JLM_TRY(make_map(&op->q_hi, op->hi, K, rows, K, box_rows / 4));
  JLM_TRY(make_map(&op->q_lo, op->lo, K, rows, K, box_rows / 4));
  return 0;
}

void free_operand(TcOperand* op) {
  cudaFree(op->hi);
  cudaFree(op->lo);
  op->hi = op->lo = nullptr;
}

// CTA pairs (cta_group::2) are used whenever the problem has at least two 128-row blocks; JLM_TC_PAIR=0
// forces the single-CTA kernel (A/B comparison, debugging).
static bool tc_pair_enabled() {
  static const int v = [] {
    const char* e = getenv("JLM_TC_PAIR");
    return e ? atoi(e) : 1;
  }();
  return v != 0;
}

static bool tc_narrow_enabled() {
  static const int v = [] {
    const char* e = getenv("JLM_TC_NARROW");
    return e ? atoi(e) : 1;
  }();
  return v != 0;
}

template <int BN, int EPI, int KS = 1>
int32_t launch_gemm(jlm_handle* h, const CUtensorMap& Ah, const CUtensorMap& Al, const TcOperand& B, GemmArgs g) {
  static_assert(KS == 1 || BN < 256, "partial accumulators: the CTA-pair tiles fill TMEM with their two stages");
  JLM_REQUIRE(g.K % BK == 0 && g.K > 0, "tc gemm: K=%d must be a positive multiple of %d", g.K, BK);
  g.num_m_blocks = ceil_div(g.M, BM);
  g.num_n_blocks = ceil_div(g.N, BN);
  if (g.num_m_blocks * g.num_n_blocks <= 0) return 0;
  if constexpr (BN == 256) if (g.num_m_blocks >= 2 && h->sm_count >= 2 && tc_pair_enabled()) {
    using C = TileCfg<BN, 2>;
    static int max_pairs = -1;      // co-resident CTA pairs the device can hold for this kernel (0: cannot launch it)
    const int tiles = ceil_div(g.num_m_blocks, 2) * g.num_n_blocks;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * std::min(tiles, h->sm_count / 2));
    cfg.blockDim = dim3(EpiCfg<EPI>::THREADS);
    cfg.dynamicSmemBytes = C::SMEM + EpiCfg<EPI>::STG;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // see pdl_enter (jlm_common.cuh)
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = jlm_pdl_enabled() ? 2 : 1;
    if (max_pairs < 0) {
      JLM_CUDA(cudaFuncSetAttribute(k_tc_gemm<BN, EPI, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM + EpiCfg<EPI>::STG));
      int n = 0;
      // a partitioned or floor-swept device may not place two-CTA clusters: fall back to the single-CTA kernel
      if (cudaOccupancyMaxActiveClusters(&n, k_tc_gemm<BN, EPI, 2>, &cfg) != cudaSuccess) {
        cudaGetLastError();
        n = 0;
      }
      max_pairs = n;
    }
    if (max_pairs > 0) {
      cfg.gridDim = dim3(2 * std::min(tiles, std::min(h->sm_count / 2, max_pairs)));
      JLM_CUDA(cudaLaunchKernelEx(&cfg, k_tc_gemm<BN, EPI, 2>, Ah, Al, B.pair_hi, B.pair_lo, g));
      return 0;
    }
  }
  using C = TileCfg<BN, 1, KS>;
  static bool configured = false;
  if (!configured) {
    JLM_CUDA(cudaFuncSetAttribute(k_tc_gemm<BN, EPI, 1, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM + EpiCfg<EPI>::STG));
    configured = true;
  }
  const int tiles = g.num_m_blocks * g.num_n_blocks;
  const int grid = tiles < h->sm_count ? tiles : h->sm_count;
  // 64-column tiles read B through the quarter-box maps (operands are uploaded with 256-row boxes)
  JLM_CUDA(jlm_launch(k_tc_gemm<BN, EPI, 1, KS>, dim3(grid), dim3(EpiCfg<EPI>::THREADS), C::SMEM + EpiCfg<EPI>::STG, h->stream, Ah, Al, BN == 64 ? B.q_hi : B.map_hi, BN == 64 ? B.q_lo : B.map_lo, g));
  JLM_CUDA(cudaGetLastError());
  return 0;
}

static bool tc_rs_enabled() {
  static const int v = [] {
    const char* e = getenv("JLM_TC_RS");
    return e ? atoi(e) : 1;
  }();
  return v != 0;
}

static bool tc_vocab_dense_enabled() {
  static const int v = [] {
    const char* e = getenv("JLM_TC_VOCAB_DENSE");
    return e ? atoi(e) : 1;
  }();
  return v != 0;
}

static bool tc_rs256_enabled() {
  static const int v = [] {
    const char* e = getenv("JLM_TC_RS256");
    return e ? atoi(e) : 0;
  }();
  return v != 0;
}

// Row-stationary launch (k_tc_lse_rs).  Returns the number of partial slots per row it writes through *n_slots,
// or 0 there when the shape / device cannot take this kernel and the caller must use k_tc_gemm<EPI_LSE>.
template <int KB>
int32_t launch_lse_rs(jlm_handle* h, const CUtensorMap& Ah, const CUtensorMap& Al, const TcOperand& B, RsArgs g, int max_slots,
                      int* n_slots) {
  using C = RsCfg<KB>;
  *n_slots = 0;
  g.num_m_units = ceil_div(g.M, 2 * BM);
  g.num_n_blocks = ceil_div(g.N, C::BN);
  const int64_t T = (int64_t)g.num_m_units * g.num_n_blocks;
  if (T <= 0 || h->sm_count < 2) return 0;
  static int max_pairs = -1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * (unsigned)std::min<int64_t>(T, h->sm_count / 2));
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::SMEM;
  cfg.stream = h->stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = jlm_pdl_enabled() ? 2 : 1;
  if (max_pairs < 0) {
    JLM_CUDA(cudaFuncSetAttribute(k_tc_lse_rs<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, k_tc_lse_rs<KB>, &cfg) != cudaSuccess) {
      cudaGetLastError();
      n = 0;
    }
    max_pairs = n;
  }
  if (max_pairs <= 0) return 0;
  const int U = (int)std::min<int64_t>(T, std::min(h->sm_count / 2, max_pairs));
  const int64_t L = T / U;                                   // tiles per pair (some get one more)
  const int slots = (int)std::min<int64_t>(std::min<int64_t>(g.num_n_blocks, U), (g.num_n_blocks + L - 1) / L + 1);
  if (C::COL_PARTS * slots > max_slots) return 0;      // one partial per (row, run, column slice)
  g.n_slots = slots;
  cfg.gridDim = dim3(2 * U);
  JLM_CUDA(cudaLaunchKernelEx(&cfg, k_tc_lse_rs<KB>, Ah, Al, B.pair_hi, B.pair_lo, g));
  *n_slots = C::COL_PARTS * slots;
  return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// handle-level tensor-core weights
// ------------------------------------------------------------------------------------------------
struct TcWeights {
  TcOperand Wg;                       // [4*Hp, Kg], tile-permuted rows
  float* bg_perm = nullptr;           // [4*Hp]
  TcOperand P1;                       // [Kt, Hp]
  TcOperand seg[JLM_MAX_SEGMENTS];    // [V_i, kpad_i]
  TcOperand Emb;                      // [V, Ep] input embedding, split at the gate-input scale sA
  float sA = 1.f;                     // gate-input scale: h (kept per slot as its fp16 split) and embedding share it
  float sT = 1.f;                     // stage-1 output scale
  int lse_tiles = 0;                  // partial slots per row: sum of lse_cnt
  int lse_col0[JLM_MAX_SEGMENTS] = {}; // first partial slot of each output segment
  int lse_cnt[JLM_MAX_SEGMENTS] = {};  // slots reserved for it: one per 256-column tile (+1: row-stationary runs)
};

constexpr int GATE_BN = 256;

// First-order compensation of the tensor core's truncating fp32 accumulate.  Measured on B200
// (scripts/tc_bias_probe.py, profiles/r02/tc_bias_probe.txt): C_tc - C_f64 = -beta.C + noise with beta proportional to
// the length K of the accumulation chain - 3.0e-9 K when the partial sums wander around zero (zero-mean products:
// gate pre-activations, stage-1 rows) and 6.8e-9 K when they drift monotonically toward a large final value (the few
// large logits that carry a softmax denominator).  The un-scale factor of each epilogue is multiplied by
// (1 + beta K).  JLM_TC_DEBIAS is a bit mask (1 gate, 2 stage-1, 4 output GEMMs; default 6).  The gate GEMM is left
// alone: measured against the float64 back end (scripts/tc_error_probe.py, profiles/r02) its compensation moves h
// AWAY from the float64 state (sum|h| error -1.5e-5 -> +3.3e-5), while the output-GEMM factor alone takes the mean
// log-sum-exp error from -2.3e-5 to +1.7e-7 (std 2.1e-6).
constexpr double RZ_WANDER = 3.0e-9, RZ_DRIFT = 6.8e-9;
enum { RZ_GATE = 1, RZ_STAGE1 = 2, RZ_OUT = 4 };      // JLM_TC_DEBIAS bit mask: which GEMMs are compensated
static float rz_comp(int which, double per_k, int chain_k) {
  static const int mask = [] {
    const char* e = getenv("JLM_TC_DEBIAS");
    return e ? atoi(e) : (RZ_STAGE1 | RZ_OUT);
  }();
  return (mask & which) ? (float)(1.0 + per_k * chain_k) : 1.f;
}
constexpr int STAGE1_KS = 4;      // partial accumulators of the stage-1 projection GEMM (TileCfg)

static void free_tc_weights(TcWeights* w);
static int32_t tc_build_weights(jlm_handle* h, TcWeights* w);

// The weights are built into a local object and published to the handle only when every upload succeeded: a
// failed upload (e.g. out of memory at a large V) must not leave a half-initialised h->tc behind.
int32_t tc_prepare_weights(jlm_handle* h) {
  if (h->tc) return 0;
  JLM_CUDA(cudaSetDevice(h->device));
  TcWeights* w = new TcWeights();
  if (tc_build_weights(h, w)) {
    free_tc_weights(w);
    return 1;
  }
  h->tc = w;
  return 0;
}

static int32_t tc_build_weights(jlm_handle* h, TcWeights* w) {
  const int H = h->H, Hp = h->Hp, Kg = h->Kg, V = h->V;
  const int UPT = GATE_BN / 4;
  // pull the exact-layout weights back from the device (they are the single source of truth)
  std::vector<float> Wg((size_t)4 * H * Kg), bg((size_t)4 * H), LM((size_t)V * h->Ep);
  JLM_CUDA(cudaMemcpy(Wg.data(), h->Wg, Wg.size() * sizeof(float), cudaMemcpyDeviceToHost));
  JLM_CUDA(cudaMemcpy(bg.data(), h->bg, bg.size() * sizeof(float), cudaMemcpyDeviceToHost));
  JLM_CUDA(cudaMemcpy(LM.data(), h->LM_in, LM.size() * sizeof(float), cudaMemcpyDeviceToHost));
  {
    std::vector<double> P((size_t)4 * Hp * Kg, 0.0);
    std::vector<float> bp((size_t)4 * Hp, 0.f);
    for (int gate = 0; gate < 4; ++gate)
      for (int j = 0; j < H; ++j) {
        const size_t dst = (size_t)(j / UPT) * GATE_BN + (size_t)gate * UPT + (j % UPT);
        const float* src = &Wg[((size_t)gate * H + j) * Kg];
        for (int k = 0; k < Kg; ++k) P[dst * Kg + k] = src[k];
        bp[dst] = bg[(size_t)gate * H + j];
      }
    JLM_TRY(upload_split(&w->Wg, P, 4 * Hp, Kg, GATE_BN));
    JLM_CUDA(cudaMalloc(&w->bg_perm, bp.size() * sizeof(float)));
    JLM_CUDA(cudaMemcpy(w->bg_perm, bp.data(), bp.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  {
    double mx = 1.0;
    for (float v : LM) mx = std::max(mx, (double)std::fabs(v));
    w->sA = pow2_scale_for(mx);
    std::vector<double> Ed(LM.begin(), LM.end());
    JLM_TRY(upload_split(&w->Emb, Ed, V, h->Ep, 256, w->sA));
  }
  if (!h->untied) {
    std::vector<double> P1((size_t)h->Kt * Hp);
    JLM_CUDA(cudaMemcpy(P1.data(), h->P1, P1.size() * sizeof(double), cudaMemcpyDeviceToHost));
    JLM_TRY(upload_split(&w->P1, P1, h->Kt, Hp, 256));
    double bound = 0;
    for (int n = 0; n < h->Kt; ++n) {
      double s = 0;
      for (int k = 0; k < Hp; ++k) s += std::fabs(P1[(size_t)n * Hp + k]);
      bound = std::max(bound, s);
    }
    w->sT = pow2_scale_for(bound);
  } else {
    w->sT = w->sA;      // untied: the output GEMM reads the h split directly
  }
  w->lse_tiles = 0;
  for (int i = 0; i < h->n_seg; ++i) {
    const SegDev& s = h->seg[i];
    const int64_t Vi = s.end - s.start;
    std::vector<float> Wf((size_t)Vi * s.kpad);
    JLM_CUDA(cudaMemcpy(Wf.data(), s.W, Wf.size() * sizeof(float), cudaMemcpyDeviceToHost));
    std::vector<double> Wd(Wf.begin(), Wf.end());
    JLM_TRY(upload_split(&w->seg[i], Wd, Vi, s.kpad, 256));
    w->lse_col0[i] = w->lse_tiles;
    w->lse_cnt[i] = ceil_div(Vi, 256) + 1;
    w->lse_tiles += w->lse_cnt[i];
  }
  return 0;
}

static void free_tc_weights(TcWeights* w) {
  free_operand(&w->Wg);
  free_operand(&w->P1);
  free_operand(&w->Emb);
  for (auto& s : w->seg) free_operand(&s);
  cudaFree(w->bg_perm);
  delete w;
}

void tc_free_weights(jlm_handle* h) {
  if (!h->tc) return;
  free_tc_weights(h->tc);
  h->tc = nullptr;
}

// ------------------------------------------------------------------------------------------------
// batch-level state
// ------------------------------------------------------------------------------------------------
struct TcBatchState {
  float* h32 = nullptr;   // [n_slots, Hp], untied projections only (their needed-word dots read fp32 h rows)
  float* c32 = nullptr;   // [n_slots, Hp]
  __half *Ag_hi = nullptr, *Ag_lo = nullptr;   // [Mpad, Kg]
  __half *Hs_hi = nullptr, *Hs_lo = nullptr;   // [n_slots, Hp]: h * sA as hi + lo, written by the LSTM epilogue
  float* T32 = nullptr;                        // [Mpad, Kt]
  __half *Ts_hi = nullptr, *Ts_lo = nullptr;   // [Mpad, Kt]
  float2* part = nullptr;
  int64_t Mpad = 0;
  // vocabulary-selection modes: the word rows every sentence's list starts with, gathered once per run into a dense
  // operand (jlm_batch::n_shared), and their logits for the step's rows
  TcOperand Sh;                       // [n_shared_pad, kpad] views into the arena (not owned)
  float* sh_bias = nullptr;           // [n_shared_pad] b2 of those words
  float* Y0 = nullptr;                // [Mpad, ldy0]
  int n_shared_pad = 0;
  // ... and ALL word rows of all sentences' lists, gathered once per batch (k_tc_vocab_dense)
  __half *Wd_hi = nullptr, *Wd_lo = nullptr;   // [dense_rows, kpad]
  float* wd_bias = nullptr;                    // [dense_rows]
  int64_t dense_rows = 0;
  CUtensorMap mWd_hi, mWd_lo, mTv_hi, mTv_lo;  // 128-row boxes over Wd, 32-row boxes over the step's stage-1 rows
  CUtensorMap mAg_hi, mAg_lo, mHs_hi, mHs_lo;
  CUtensorMap mTs_hi[JLM_MAX_SEGMENTS], mTs_lo[JLM_MAX_SEGMENTS];
};

int32_t tc_batch_plan(jlm_batch* b, Arena& a) {
  jlm_handle* h = b->h;
  JLM_TRY(tc_prepare_weights(h));
  if (!b->tc) b->tc = new TcBatchState();
  TcBatchState* s = b->tc;
  const size_t ns = (size_t)b->n_slots;
  s->Mpad = round_up64(std::max(b->max_rows_step, 1), BM);
  const size_t mp = (size_t)s->Mpad;
  s->h32 = h->untied ? a.take<float>(ns * h->Hp) : nullptr;
  s->c32 = a.take<float>(ns * h->Hp);
  s->Ag_hi = a.take<__half>(mp * h->Kg);
  s->Ag_lo = a.take<__half>(mp * h->Kg);
  s->Hs_hi = a.take<__half>((ns + BM) * h->Hp);      // + one tile: the last step's last TMA box stays inside
  s->Hs_lo = a.take<__half>((ns + BM) * h->Hp);
  if (!h->untied) {
    s->T32 = a.take<float>(mp * h->Kt);
    s->Ts_hi = a.take<__half>(mp * h->Kt);
    s->Ts_lo = a.take<__half>(mp * h->Kt);
  } else {
    s->T32 = nullptr;
    s->Ts_hi = s->Hs_hi;
    s->Ts_lo = s->Hs_lo;
  }
  s->part = (b->mode == JLM_DECODE_FULL && b->use_lse) ? a.take<float2>(mp * h->tc->lse_tiles) : nullptr;
  // shared word rows: only where the tensor-core vocabulary kernel runs (tc_vocab_logits)
  const bool vocab_tc = !(b->mode == JLM_DECODE_FULL || h->untied || h->n_seg != 1 || b->W > VT_N || h->seg[0].kpad % BK != 0 || !b->use_lse);
  if (!vocab_tc) b->n_shared = 0;
  s->dense_rows = 0;
  s->Wd_hi = s->Wd_lo = nullptr;
  s->wd_bias = nullptr;
  if (vocab_tc && tc_vocab_dense_enabled() && b->n_vocab_cols > 0) {
    const int64_t rows = round_up64(b->n_vocab_cols, VT_M) + VT_M;      // a tile may start at any row: one spare tile
    if (rows * h->seg[0].kpad * 4 <= ((int64_t)8 << 30)) {             // hi + lo, 2 bytes each; beyond 8 GB: per-frame gather
      s->dense_rows = rows;
      s->Wd_hi = a.take<__half>((size_t)rows * h->seg[0].kpad);
      s->Wd_lo = a.take<__half>((size_t)rows * h->seg[0].kpad);
      s->wd_bias = a.take<float>((size_t)rows);
    }
  }
  s->n_shared_pad = (int)round_up64(b->n_shared, 256);
  if (b->n_shared > 0) {
    s->Sh.hi = a.take<__half>((size_t)s->n_shared_pad * h->seg[0].kpad);
    s->Sh.lo = a.take<__half>((size_t)s->n_shared_pad * h->seg[0].kpad);
    s->sh_bias = a.take<float>((size_t)s->n_shared_pad);
    s->Y0 = a.take<float>(mp * s->n_shared_pad);
    b->ldy0 = s->n_shared_pad;
    b->y0 = s->Y0;
  } else {
    b->y0 = nullptr;
    b->ldy0 = 0;
  }
  if (a.dry) return 0;
  if (b->n_shared > 0) {
    const int64_t K = h->seg[0].kpad;
    s->Sh.rows = s->n_shared_pad;
    s->Sh.K = K;
    s->Sh.scale = h->tc->seg[0].scale;
    JLM_TRY(make_map(&s->Sh.map_hi, s->Sh.hi, K, s->n_shared_pad, K, 256));
    JLM_TRY(make_map(&s->Sh.map_lo, s->Sh.lo, K, s->n_shared_pad, K, 256));
    JLM_TRY(make_map(&s->Sh.pair_hi, s->Sh.hi, K, s->n_shared_pad, K, 128));
    JLM_TRY(make_map(&s->Sh.pair_lo, s->Sh.lo, K, s->n_shared_pad, K, 128));
    JLM_TRY(make_map(&s->Sh.q_hi, s->Sh.hi, K, s->n_shared_pad, K, 64));
    JLM_TRY(make_map(&s->Sh.q_lo, s->Sh.lo, K, s->n_shared_pad, K, 64));
  }
  if (s->dense_rows > 0) {
    const int64_t K = h->seg[0].kpad;
    JLM_TRY(make_map(&s->mWd_hi, s->Wd_hi, K, s->dense_rows, K, VT_M));
    JLM_TRY(make_map(&s->mWd_lo, s->Wd_lo, K, s->dense_rows, K, VT_M));
    JLM_TRY(make_map(&s->mTv_hi, s->Ts_hi + h->seg[0].koff, K, s->Mpad, h->Kt, VT_N));
    JLM_TRY(make_map(&s->mTv_lo, s->Ts_lo + h->seg[0].koff, K, s->Mpad, h->Kt, VT_N));
  }
  JLM_TRY(make_map(&s->mAg_hi, s->Ag_hi, h->Kg, s->Mpad, h->Kg, BM));
  JLM_TRY(make_map(&s->mAg_lo, s->Ag_lo, h->Kg, s->Mpad, h->Kg, BM));
  const int64_t hs_rows = (int64_t)ns + BM;
  JLM_TRY(make_map(&s->mHs_hi, s->Hs_hi, h->Hp, hs_rows, h->Hp, BM));
  JLM_TRY(make_map(&s->mHs_lo, s->Hs_lo, h->Hp, hs_rows, h->Hp, BM));
  const int64_t ldT = h->untied ? h->Hp : h->Kt;
  const int64_t t_rows = h->untied ? hs_rows : s->Mpad;      // untied: T aliases the per-slot h split
  for (int i = 0; i < h->n_seg; ++i) {
    JLM_TRY(make_map(&s->mTs_hi[i], s->Ts_hi + h->seg[i].koff, h->seg[i].kpad, t_rows, ldT, BM));
    JLM_TRY(make_map(&s->mTs_lo[i], s->Ts_lo + h->seg[i].koff, h->seg[i].kpad, t_rows, ldT, BM));
  }
  return 0;
}

void tc_batch_free(jlm_batch* b) {
  delete b->tc;
  b->tc = nullptr;
}

// LM step, first half: gather -> gate GEMM (+LSTM epilogue) -> stage-1 projection.  Returns the fp32 stage-1 rows.
int32_t tc_batch_lm_state(jlm_batch* b, int t, const float** T_out, int* ldt_out) {
  jlm_handle* h = b->h;
  TcWeights* w = h->tc;
  TcBatchState* s = b->tc;
  cudaStream_t st = h->stream;
  const StepPlan& sp = b->steps[t];
  const int M = sp.rows_step;
  *T_out = nullptr;
  *ldt_out = 0;
  if (M == 0) return 0;
  const int32_t* parent = b->d.slot_parent + sp.row0;
  const int32_t* word = b->d.slot_word + sp.row0;
  float* hrow = s->h32 ? s->h32 + sp.row0 * h->Hp : nullptr;
  float* crow = s->c32 + sp.row0 * h->Hp;
  if (b->timers) cudaEventRecord(b->events[3 * t + 1], st);
  {
    const int64_t total = (int64_t)M * (h->Kg / 8);
    int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)h->sm_count * 16);
    JLM_CUDA(jlm_launch(k_tc_gather_rows, dim3(grid), dim3(256), 0, st, s->Hs_hi, s->Hs_lo, parent, word, w->Emb.hi, w->Emb.lo, h->Hp, h->Ep, M, s->Ag_hi, s->Ag_lo));
    JLM_CUDA(cudaGetLastError());
  }
  {
    GemmArgs g{};
    g.M = M;
    g.N = 4 * h->Hp;
    g.K = h->Kg;
    g.inv_scale = rz_comp(RZ_GATE, RZ_WANDER, h->Kg) / (w->sA * w->Wg.scale);
    g.bias = w->bg_perm;
    g.c_src = s->c32;
    g.parent = parent;
    g.h_out = hrow;
    g.c_out = crow;
    g.ld_state = h->Hp;
    g.s_hi = s->Hs_hi + sp.row0 * h->Hp;
    g.s_lo = s->Hs_lo + sp.row0 * h->Hp;
    g.lds = h->Hp;
    g.split_scale = w->sA;
    if (b->timers) cudaEventRecord(b->kev[4 * t], st);
    JLM_TRY((launch_gemm<GATE_BN, EPI_LSTM>(h, s->mAg_hi, s->mAg_lo, w->Wg, g)));
    if (b->timers) cudaEventRecord(b->kev[4 * t + 1], st);
  }
  b->launches += 2;
  if (b->timers) cudaEventRecord(b->events[3 * t + 2], st);
  const float* T32 = hrow;
  int ldt = h->Hp;
  if (!h->untied) {
    GemmArgs g{};
    g.M = M;
    g.N = h->Kt;
    g.K = h->Hp;
    g.inv_scale = 1.f / (w->sA * w->P1.scale);     // compensation factor chosen with the kernel below
    g.a_row0 = (int)sp.row0;
    g.C32 = s->T32;
    g.ldc = h->Kt;
    g.s_hi = s->Ts_hi;
    g.s_lo = s->Ts_lo;
    g.lds = h->Kt;
    g.split_scale = w->sT;
    // h.PM has few output columns (Kt = E): 256-column tiles give ceil(M/128) CTAs, each with an epilogue
    // nothing overlaps.  64-column tiles quadruple the tile count so every SM works and the store
    // epilogue of one tile runs under the MMAs of the next.
    // Every logit inherits the error of these rows, so the narrow kernel also cuts K into STAGE1_KS accumulation
    // chains summed in float64 by its epilogue (TileCfg): 4x shorter chains, 4x less truncation error.
    const int tiles256 = ceil_div(M, BM) * ceil_div(h->Kt, 256);
    if (tc_narrow_enabled()) {
      const int chain = ceil_div(h->Hp / BK, STAGE1_KS) * BK;
      g.inv_scale *= rz_comp(RZ_STAGE1, RZ_WANDER, chain);
      JLM_TRY((launch_gemm<64, EPI_STORE, STAGE1_KS>(h, s->mHs_hi, s->mHs_lo, w->P1, g)));
    } else {
      g.inv_scale *= rz_comp(RZ_STAGE1, RZ_WANDER, h->Hp);
      JLM_TRY((launch_gemm<256, EPI_STORE>(h, s->mHs_hi, s->mHs_lo, w->P1, g)));
    }
    (void)tiles256;
    b->launches += 1;
    T32 = s->T32;
    ldt = h->Kt;
  }
  *T_out = T32;
  *ldt_out = ldt;
  return 0;
}

// LM step, second half: full-vocabulary logits + online LSE (one GEMM per segment) -> per-row LSE.
int32_t tc_batch_lm_lse(jlm_batch* b, int t) {
  jlm_handle* h = b->h;
  TcWeights* w = h->tc;
  TcBatchState* s = b->tc;
  cudaStream_t st = h->stream;
  const StepPlan& sp = b->steps[t];
  const int M = sp.rows_step;
  if (M == 0) return 0;
  if (b->use_lse && b->mode == JLM_DECODE_FULL) {
    LseCols cols{};
    cols.nseg = h->n_seg;
    if (b->timers) cudaEventRecord(b->kev[4 * t + 2], st);
    for (int i = 0; i < h->n_seg; ++i) {
      const SegDev& sg = h->seg[i];
      const int N = sg.end - sg.start;
      const int a_row0 = h->untied ? (int)sp.row0 : 0;      // untied: A is the per-slot h split itself
      const int k16 = (int)round_up64(sg.width, 16);
      cols.col0[i] = w->lse_col0[i];
      int rs_slots = 0;
      // short K (D-softmax* tail segments): row-stationary kernel, A resident, only the weight tiles stream
      // JLM_TC_RS256=1 also sends K = 256 (the cfg-2 output block) here: with A resident the pair asks L2 for 256 KB per
      // tile instead of 512 KB and a row's 196 tile partials become <= 16 run partials - measured +3 % on the cfg-2 step
      // (13.8 -> 13.1 ms).  Off by default: where a pair's tile range cuts a row block depends on the block's position,
      // so a sentence's fp32 partial sums - hence its scores, at 1e-7 - would depend on its place in the batch, and
      // tests/test_gpu_parity.py::test_full_size_batch_properties demands bit-identical results under a permutation.
      const bool rs_shape = sg.kpad <= 2 * BK || (sg.kpad == 4 * BK && tc_rs256_enabled());
      if (rs_shape && M >= 2 * BM && tc_pair_enabled() && tc_rs_enabled()) {
        RsArgs r{};
        r.M = M;
        r.N = N;
        r.K16 = k16;
        r.a_row0 = a_row0;
        r.inv_scale = rz_comp(RZ_OUT, RZ_DRIFT, k16) / (w->sT * w->seg[i].scale);
        r.bias = h->b2 + sg.start;
        r.part = s->part;
        r.part_ld = w->lse_tiles;
        r.part_col0 = w->lse_col0[i];
        if (sg.kpad == BK) JLM_TRY((launch_lse_rs<1>(h, s->mTs_hi[i], s->mTs_lo[i], w->seg[i], r, w->lse_cnt[i], &rs_slots)));
        else if (sg.kpad == 2 * BK) JLM_TRY((launch_lse_rs<2>(h, s->mTs_hi[i], s->mTs_lo[i], w->seg[i], r, w->lse_cnt[i], &rs_slots)));
        else JLM_TRY((launch_lse_rs<4>(h, s->mTs_hi[i], s->mTs_lo[i], w->seg[i], r, w->lse_cnt[i], &rs_slots)));
      }
      if (rs_slots > 0) {
        cols.cnt[i] = rs_slots;
      } else {
        GemmArgs g{};
        g.M = M;
        g.N = N;
        g.K = sg.kpad;
        g.K16 = k16;
        g.inv_scale = rz_comp(RZ_OUT, RZ_DRIFT, k16) / (w->sT * w->seg[i].scale);
        g.bias = h->b2 + sg.start;
        g.part = s->part;
        g.part_ld = w->lse_tiles;
        g.part_tile0 = w->lse_col0[i];
        g.a_row0 = a_row0;
        JLM_TRY((launch_gemm<256, EPI_LSE>(h, s->mTs_hi[i], s->mTs_lo[i], w->seg[i], g)));
        cols.cnt[i] = ceil_div(N, 256);
      }
      b->launches += 1;
    }
    if (b->timers) cudaEventRecord(b->kev[4 * t + 3], st);
    JLM_CUDA(jlm_launch(k_tc_lse_merge, dim3(ceil_div((int64_t)M * 32, 256)), dim3(256), 0, st, s->part, w->lse_tiles, cols, M, b->d.slot_lse + sp.row0));
    JLM_CUDA(cudaGetLastError());
    b->launches += 1;
  }
  return 0;
}

int32_t tc_batch_get_state(jlm_batch* b, int64_t slot, int count, double* h_out, double* c_out) {
  jlm_handle* h = b->h;
  const size_t n = (size_t)count * h->Hp;
  if (c_out) {
    std::vector<float> tmp(n);
    JLM_CUDA(cudaMemcpy(tmp.data(), b->tc->c32 + slot * h->Hp, n * sizeof(float), cudaMemcpyDeviceToHost));
    for (int k = 0; k < count; ++k)
      for (int j = 0; j < h->H; ++j) c_out[(size_t)k * h->H + j] = (double)tmp[(size_t)k * h->Hp + j];
  }
  if (h_out) {
    // h lives as its fp16 split at the gate-input scale: h = (hi + lo) / sA
    std::vector<__half> hi(n), lo(n);
    JLM_CUDA(cudaMemcpy(hi.data(), b->tc->Hs_hi + slot * h->Hp, n * sizeof(__half), cudaMemcpyDeviceToHost));
    JLM_CUDA(cudaMemcpy(lo.data(), b->tc->Hs_lo + slot * h->Hp, n * sizeof(__half), cudaMemcpyDeviceToHost));
    const double inv = 1.0 / (double)h->tc->sA;
    for (int k = 0; k < count; ++k)
      for (int j = 0; j < h->H; ++j) {
        const size_t i = (size_t)k * h->Hp + j;
        h_out[(size_t)k * h->H + j] = ((double)__half2float(hi[i]) + (double)__half2float(lo[i])) * inv;
      }
  }
  return 0;
}

// Vocabulary-selection logits for step t on the tensor cores; returns 2 when this back end cannot take the
// job shape (segmented projection, beam > 32) and the float64 kernel must run instead, 1 on a CUDA error.
int32_t tc_vocab_logits(jlm_batch* b, int t, double* out) {
  jlm_handle* h = b->h;
  TcWeights* w = h->tc;
  TcBatchState* s = b->tc;
  const StepPlan& sp = b->steps[t];
  static const int enabled = [] {
    const char* e = getenv("JLM_TC_VOCAB");
    return e ? atoi(e) : 1;
  }();
  if (!enabled || !w || !s || h->untied || h->n_seg != 1 || b->W > VT_N || h->seg[0].kpad % BK != 0) return 2;
  if (sp.nstep <= 0 || sp.max_vocab_cols <= 0) return 0;
  static bool configured = false;
  if (!configured) {
    JLM_CUDA(cudaFuncSetAttribute(k_tc_vocab_logits, cudaFuncAttributeMaxDynamicSharedMemorySize, VT_SMEM));
    configured = true;
  }
  const float inv_scale = rz_comp(RZ_OUT, RZ_DRIFT, h->seg[0].kpad) / (w->sT * w->seg[0].scale);
  const int ns = b->n_shared;
  if (ns > 0) {
    if (t == 0) {      // the shared word rows of this batch (the first ns ids of any sentence's list)
      JLM_CUDA(jlm_launch(k_tc_gather_shared, dim3(std::min(ceil_div((int64_t)s->n_shared_pad * (h->seg[0].kpad / 8), 256), h->sm_count * 8)), dim3(256), 0, h->stream, w->seg[0].hi, w->seg[0].lo, h->seg[0].kpad, h->b2, b->d.vocab_cols, ns, s->n_shared_pad, s->Sh.hi, s->Sh.lo, s->sh_bias));
      JLM_CUDA(cudaGetLastError());
      b->launches += 1;
    }
    GemmArgs g{};
    g.M = sp.rows_step;
    g.N = s->n_shared_pad;
    g.K = h->seg[0].kpad;
    g.inv_scale = inv_scale;
    g.bias = s->sh_bias;
    g.C32 = s->Y0;
    g.ldc = s->n_shared_pad;
    JLM_TRY((launch_gemm<256, EPI_STORE>(h, s->mTs_hi[0], s->mTs_lo[0], s->Sh, g)));
    b->launches += 1;
  }
  const int gx = ceil_div(std::max(sp.max_vocab_cols - ns, 0), VT_M);
  if (s->dense_rows > 0 && gx > 0) {
    static bool dense_configured = false;
    if (!dense_configured) {
      JLM_CUDA(cudaFuncSetAttribute(k_tc_vocab_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, VT_SMEM));
      dense_configured = true;
    }
    if (t == 0) {      // every sentence's word rows, once per run
      JLM_CUDA(jlm_launch(k_tc_gather_shared, dim3(std::min(ceil_div(s->dense_rows * (h->seg[0].kpad / 8), 256), h->sm_count * 16)),
                          dim3(256), 0, h->stream, w->seg[0].hi, w->seg[0].lo, h->seg[0].kpad, h->b2, b->d.vocab_cols,
                          (int)b->n_vocab_cols, (int)s->dense_rows, s->Wd_hi, s->Wd_lo, s->wd_bias));
      b->launches += 1;
    }
    for (int j0 = 0; j0 < sp.nstep; j0 += 65535) {
      const int nj = std::min(sp.nstep - j0, 65535);
      JLM_CUDA(jlm_launch(k_tc_vocab_dense, dim3(gx, nj), dim3(128), VT_SMEM, h->stream, s->mWd_hi, s->mWd_lo, s->mTv_hi,
                          s->mTv_lo, h->seg[0].kpad, b->d.vocab_jobs + sp.job0 + j0, s->wd_bias, inv_scale, out, ns));
    }
    return 0;
  }
  for (int j0 = 0; j0 < sp.nstep && gx > 0; j0 += 65535) {
    const int nj = std::min(sp.nstep - j0, 65535);
    JLM_CUDA(jlm_launch(k_tc_vocab_logits, dim3(dim3(gx, nj)), dim3(128), VT_SMEM, h->stream,  w->seg[0].hi, w->seg[0].lo, h->seg[0].kpad, s->Ts_hi + h->seg[0].koff, s->Ts_lo + h->seg[0].koff, h->Kt, h->seg[0].kpad, b->d.vocab_jobs + sp.job0 + j0, b->d.vocab_cols, h->b2, inv_scale, out, ns));
  }
  JLM_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// self test: C = A . B^T through the split-fp16 kernel (EPI_STORE)
// ------------------------------------------------------------------------------------------------
extern "C" int32_t jlm_tc_gemm_selftest(jlm_handle* h, const float* A, const float* B, int32_t M, int32_t N, int32_t K,
                                        float* C, float* ms) {
  JLM_REQUIRE(h && A && B && C && M > 0 && N > 0 && K > 0 && K % BK == 0, "jlm_tc_gemm_selftest: bad argument");
  JLM_CUDA(cudaSetDevice(h->device));
  TcOperand a, bop;
  std::vector<double> Ad(A, A + (size_t)M * K), Bd(B, B + (size_t)N * K);
  int32_t rc = upload_split(&a, Ad, M, K, BM);
  if (!rc) rc = upload_split(&bop, Bd, N, K, 256);
  float* dC = nullptr;
  if (!rc && cudaMalloc(&dC, sizeof(float) * (size_t)M * N) != cudaSuccess) {
    jlm_set_error("selftest: cudaMalloc failed");
    rc = 1;
  }
  if (!rc) {
    GemmArgs g{};
    g.M = M;
    g.N = N;
    g.K = K;
    g.inv_scale = 1.f / (a.scale * bop.scale);
    g.C32 = dC;
    g.ldc = N;
    cudaEventRecord(h->ev[0], h->stream);
    rc = launch_gemm<256, EPI_STORE>(h, a.map_hi, a.map_lo, bop, g);
    cudaEventRecord(h->ev[1], h->stream);
    if (!rc && cudaStreamSynchronize(h->stream) != cudaSuccess) {
      jlm_set_error("selftest: kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
      rc = 1;
    }
    if (!rc) {
      float t = 0.f;
      cudaEventElapsedTime(&t, h->ev[0], h->ev[1]);
      if (ms) *ms = t;
      if (cudaMemcpy(C, dC, sizeof(float) * (size_t)M * N, cudaMemcpyDeviceToHost) != cudaSuccess) {
        jlm_set_error("selftest: D2H failed");
        rc = 1;
      }
    }
  }
  cudaFree(dC);
  free_operand(&a);
  free_operand(&bop);
  return rc;
}
