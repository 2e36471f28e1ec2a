// Exact back end: CUDA-core kernels with float64 accumulation over float32 weights.
//
// The reference keeps hidden/cell/logits in float64 (decoder/model.py:44-45; np.dot(float64,float32)
// promotes), so the few-rows-in-flight path (one sentence, <= beam rows) mirrors that arithmetic
// here.  These kernels are weight-streaming (HBM/L2 bound): every weight element is read once per
// LM step and used for all rows of the step.  Large lock-step batches use the tensor-core back end
// (jlm_tc.cu); this one doubles as its on-device float64 cross-check.
#include <algorithm>

#include "jlm_common.cuh"

namespace {

constexpr int BN = 64;
constexpr int BK = 32;

template <typename TB, int TM>
__global__ void __launch_bounds__(256)
k_gemm_f64(const double* __restrict__ A, int lda, const TB* __restrict__ B, int ldb,
           const float* __restrict__ bias, double* __restrict__ C, int64_t ldc, int M, int N, int K,
           double2* __restrict__ part, int part_ld, int part_tile0) {
  pdl_enter();
  constexpr int BM = 16 * TM;
  __shared__ double As[BM][BK + 2];
  __shared__ TB Bs[BK][BN + 1];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * BM;

  double acc[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

  for (int k0 = 0; k0 < K; k0 += BK) {
    // A tile: BM x BK doubles as double2
    for (int idx = tid; idx < BM * BK / 2; idx += 256) {
      int m = idx / (BK / 2), kq = idx % (BK / 2);
      double2 v = make_double2(0.0, 0.0);
      if (m0 + m < M) v = *reinterpret_cast<const double2*>(A + (int64_t)(m0 + m) * lda + k0 + kq * 2);
      As[m][kq * 2] = v.x;
      As[m][kq * 2 + 1] = v.y;
    }
    // B tile: BN x BK elements, transposed into Bs[k][n]
    if (sizeof(TB) == 4) {
      for (int idx = tid; idx < BN * BK / 4; idx += 256) {
        int n = idx / (BK / 4), kq = idx % (BK / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n0 + n < N)
          v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(B) + (int64_t)(n0 + n) * ldb + k0 + kq * 4);
        Bs[kq * 4 + 0][n] = (TB)v.x;
        Bs[kq * 4 + 1][n] = (TB)v.y;
        Bs[kq * 4 + 2][n] = (TB)v.z;
        Bs[kq * 4 + 3][n] = (TB)v.w;
      }
    } else {
      for (int idx = tid; idx < BN * BK / 2; idx += 256) {
        int n = idx / (BK / 2), kq = idx % (BK / 2);
        double2 v = make_double2(0.0, 0.0);
        if (n0 + n < N)
          v = *reinterpret_cast<const double2*>(reinterpret_cast<const double*>(B) + (int64_t)(n0 + n) * ldb + k0 + kq * 2);
        Bs[kq * 2 + 0][n] = (TB)v.x;
        Bs[kq * 2 + 1][n] = (TB)v.y;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < BK; ++k) {
      double a[TM], b[4];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[ty * TM + i][k];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = (double)Bs[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    double v[4];
    double mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx + 16 * j;
      v[j] = acc[i][j];
      if (n < N) {
        if (bias) v[j] += (double)bias[n];
        if (C && m < M) C[(int64_t)m * ldc + n] = v[j];
        mx = fmax(mx, v[j]);
      } else {
        v[j] = -INFINITY;
      }
    }
    if (part) {
#pragma unroll
      for (int o = 8; o >= 1; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) s += (v[j] == -INFINITY) ? 0.0 : exp(v[j] - mx);
#pragma unroll
      for (int o = 8; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (tx == 0 && m < M) part[(int64_t)m * part_ld + part_tile0 + blockIdx.x] = make_double2(mx, s);
    }
  }
}

// Few rows in flight (one sentence: M <= beam rows): the GEMM is a weight stream.  One CTA = one tile of
// COLS columns; thread = (CPT columns 32 apart, one K slice): it streams its slice of those weight rows
// (every weight byte is read once) and accumulates all M rows in float64 registers against broadcast reads
// of A, which sits in shared memory as doubles - no cross-lane reduction in the K loop.  The K slices are
// then summed through shared memory (fixed order) and, for 64-column tiles, the tile's (max, sum exp)
// partial is formed in the CTA.  HBM/L2-bound: algorithmic bytes = N*K*sizeof(weight) per launch.
// Weights are float32, float64 (stage-1 matrix) or 8-bit k-means codes (train/comp.py) decoded through a
// 256-entry codebook held in shared memory; the code path issues the same products in the same order as
// the float32 path, so its results are bit-identical.
// Shapes: <COLS=64, CPT=2, KG=8> when LSE partials are wanted or N is large; <32, 1, 16> for the small
// gate / stage-1 GEMMs, which need more CTAs and shorter K loops.
template <typename TB>
struct SkLoad {                                           // one 16-byte weight load
  static constexpr int VEC = 16 / sizeof(TB);
  static __device__ __forceinline__ void load(const TB* p, const double* cb, double (&w)[VEC]) {
    if constexpr (sizeof(TB) == 4) {
      const float4 v = *reinterpret_cast<const float4*>(p);
      w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    } else if constexpr (sizeof(TB) == 8) {
      const double2 v = *reinterpret_cast<const double2*>(p);
      w[0] = v.x; w[1] = v.y;
    } else {
      const uint4 v = *reinterpret_cast<const uint4*>(p);
      const uint32_t q[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        w[4 * j] = cb[q[j] & 0xff];
        w[4 * j + 1] = cb[(q[j] >> 8) & 0xff];
        w[4 * j + 2] = cb[(q[j] >> 16) & 0xff];
        w[4 * j + 3] = cb[q[j] >> 24];
      }
    }
  }
};

template <typename TB, int MT, int COLS, int CPT, int KG>
__global__ void __launch_bounds__(COLS / CPT * KG)
k_skinny_f64(const double* __restrict__ A, int lda, const TB* __restrict__ B, int ldb,
             const float* __restrict__ codebook, const float* __restrict__ bias, double* __restrict__ C,
             int64_t ldc, int M, int N, int K, double2* __restrict__ part, int part_ld, int part_tile0) {
  pdl_enter();
  extern __shared__ __align__(16) double smem_sk[];
  constexpr int SK_THREADS = COLS / CPT * KG;
  constexpr int VEC = SkLoad<TB>::VEC;
  constexpr int WPG = COLS / CPT / 32;                    // warps per K slice
  double* As = smem_sk;                                   // [M][K]
  double* Ps = smem_sk + (size_t)M * K;                   // [KG][MT][COLS] partial sums, then values
  double* cb = Ps + (size_t)KG * MT * COLS;               // [256] codebook (8-bit weights only)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (sizeof(TB) == 1)
    for (int i = tid; i < 256; i += SK_THREADS) cb[i] = (double)codebook[i];
  for (int i = tid * 2; i < M * K; i += SK_THREADS * 2) {
    const int m = i / K, k = i % K;
    *reinterpret_cast<double2*>(&As[i]) = *reinterpret_cast<const double2*>(A + (int64_t)m * lda + k);
  }
  __syncthreads();
  const int g = warp / WPG;                               // K slice
  const int c = (warp % WPG) * 32 * CPT + lane;           // first column inside the tile
  const int n = blockIdx.x * COLS + c;
  const int kq = K / KG;
  double acc[CPT][MT];
#pragma unroll
  for (int j = 0; j < CPT; ++j)
#pragma unroll
    for (int m = 0; m < MT; ++m) acc[j][m] = 0.0;
  if (n < N) {
    const TB* brow[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) brow[j] = B + (int64_t)min(n + 32 * j, N - 1) * ldb + g * kq;
    const double* arow = As + g * kq;
#pragma unroll 2
    for (int k = 0; k < kq; k += VEC) {
      double w[CPT][VEC];
#pragma unroll
      for (int j = 0; j < CPT; ++j) SkLoad<TB>::load(brow[j] + k, cb, w[j]);
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        if (m < M) {
#pragma unroll
          for (int i = 0; i < VEC; i += 2) {
            const double2 a2 = *reinterpret_cast<const double2*>(arow + m * K + k + i);
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
              acc[j][m] = fma(a2.x, w[j][i], acc[j][m]);
              acc[j][m] = fma(a2.y, w[j][i + 1], acc[j][m]);
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < CPT; ++j)
#pragma unroll
    for (int m = 0; m < MT; ++m) Ps[(g * MT + m) * COLS + c + 32 * j] = acc[j][m];
  __syncthreads();
  // sum the K slices (fixed order): value (m, c) lands in Ps[m][c] of slice 0
  for (int i = tid; i < M * COLS; i += SK_THREADS) {
    const int m = i / COLS, cc = i % COLS;
    const int nn = blockIdx.x * COLS + cc;
    double v = -INFINITY;
    if (nn < N) {
      v = 0.0;
#pragma unroll
      for (int q = 0; q < KG; ++q) v += Ps[(q * MT + m) * COLS + cc];
      if (bias) v += (double)bias[nn];
      if (C) C[(int64_t)m * ldc + nn] = v;
    }
    if (COLS == BN) Ps[m * COLS + cc] = v;
  }
  if (COLS == BN && part) {
    __syncthreads();
    for (int m = warp; m < M; m += SK_THREADS / 32) {      // one warp per row: 64 columns, 2 per lane
      const double v0 = Ps[m * BN + lane], v1 = Ps[m * BN + 32 + lane];
      double mx = fmax(v0, v1);
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      double sm = (v0 == -INFINITY ? 0.0 : exp(v0 - mx)) + (v1 == -INFINITY ? 0.0 : exp(v1 - mx));
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
      if (lane == 0) part[(int64_t)m * part_ld + part_tile0 + blockIdx.x] = make_double2(mx, sm);
    }
  }
}

// Weight-streaming GEMM for few rows, float32 weights with K <= 256 (the output blocks: K = E or a D-softmax segment
// width) - the kernel behind single-sentence latency, the near-tie guard's float64 re-scoring and its re-decodes.
// k_skinny_f64 reads a weight row per LANE (32 different 128-byte lines per load instruction), which holds it to
// ~1 TB/s out of an L2-resident block.  Here a persistent CTA copies whole 64-column weight tiles into shared
// memory with cp.async - every warp reads 512 contiguous bytes of ONE weight row, 16-byte pieces stored XOR-swizzled
// by row so that the per-column reads below are conflict-free - double-buffered: tile i+1 lands while tile i is
// multiplied.  Compute as before: thread = 2 columns x one K slice, all rows accumulate in float64 registers against
// broadcast reads of A; the K slices are summed through shared memory in a fixed order; 64-column (max, sum exp)
// partials for the log-sum-exp.  Same products, same summation order as k_skinny_f64<.., 64, 2, KG=4>.
constexpr int ST_COLS = 64;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int MT, int ST_KG>
__global__ void __launch_bounds__(32 * ST_KG)
k_stream_f64(const double* __restrict__ A, int lda, const float* __restrict__ B, int ldb, const float* __restrict__ bias,
             double* __restrict__ C, int64_t ldc, int M, int N, int K, double2* __restrict__ part, int part_ld,
             int part_tile0) {
  pdl_enter();
  constexpr int ST_THREADS = 32 * ST_KG;
  extern __shared__ __align__(128) unsigned char smem_st[];
  const int tile_bytes = ST_COLS * K * 4;
  float* Wt[2] = {reinterpret_cast<float*>(smem_st), reinterpret_cast<float*>(smem_st + tile_bytes)};
  double* As = reinterpret_cast<double*>(smem_st + 2 * tile_bytes);   // [M][K]
  double* Ps = As + (size_t)M * K;                                    // [KG][MT][COLS]
  const int tid = threadIdx.x, lane = tid & 31, g = tid >> 5;
  const int n_tiles = (N + ST_COLS - 1) / ST_COLS;
  const int cpr = K / 4;                                              // 16-byte pieces per weight row
  auto issue = [&](int tile, float* dst) {
    // piece j of column c goes to c*K*4 + (j & ~7 | (j ^ c) & 7) * 16: lanes reading piece j of 8 consecutive columns
    // then touch 8 different 16-byte bank groups
    const int n0 = tile * ST_COLS;
    for (int c = g; c < ST_COLS; c += ST_KG) {                        // a warp copies whole weight rows
      const float* src = B + (int64_t)min(n0 + c, N - 1) * ldb;       // columns past N repeat the last row; never stored
      char* drow = reinterpret_cast<char*>(dst) + (size_t)c * K * 4;
      for (int j = lane; j < cpr; j += 32) cp_async16(drow + (((j & ~7) | ((j ^ c) & 7)) << 4), src + j * 4);
    }
    cp_async_commit();
  };
  int tile = blockIdx.x;
  if (tile < n_tiles) issue(tile, Wt[0]);
  for (int i = tid * 2; i < M * K; i += ST_THREADS * 2) {
    const int m = i / K, k = i % K;
    *reinterpret_cast<double2*>(&As[i]) = *reinterpret_cast<const double2*>(A + (int64_t)m * lda + k);
  }
  const int kq = K / ST_KG;                                           // this thread's K slice [g*kq, (g+1)*kq)
  int buf = 0;
  for (; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
    const int next = tile + gridDim.x;
    if (next < n_tiles) {
      issue(next, Wt[buf ^ 1]);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();                                                  // tile (and, first time, A) visible to all
    double acc[2][MT];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int m = 0; m < MT; ++m) acc[j][m] = 0.0;
    const char* w0 = reinterpret_cast<const char*>(Wt[buf]) + (size_t)lane * K * 4;
    const char* w1 = w0 + (size_t)32 * K * 4;                         // column lane + 32: same (c & 7)
    const double* arow = As + g * kq;
    for (int k = 0; k < kq; k += 4) {
      const int j = (g * kq + k) >> 2;
      const int off = ((j & ~7) | ((j ^ lane) & 7)) << 4;
      const float4 a0 = *reinterpret_cast<const float4*>(w0 + off);
      const float4 a1 = *reinterpret_cast<const float4*>(w1 + off);
      const double w[2][4] = {{(double)a0.x, (double)a0.y, (double)a0.z, (double)a0.w},
                              {(double)a1.x, (double)a1.y, (double)a1.z, (double)a1.w}};
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        if (m < M) {
#pragma unroll
          for (int i = 0; i < 4; i += 2) {
            const double2 a2 = *reinterpret_cast<const double2*>(arow + m * K + k + i);
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              acc[jj][m] = fma(a2.x, w[jj][i], acc[jj][m]);
              acc[jj][m] = fma(a2.y, w[jj][i + 1], acc[jj][m]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int m = 0; m < MT; ++m) Ps[(g * MT + m) * ST_COLS + lane + 32 * j] = acc[j][m];
    __syncthreads();
    for (int i = tid; i < M * ST_COLS; i += ST_THREADS) {
      const int m = i / ST_COLS, cc = i % ST_COLS;
      const int nn = tile * ST_COLS + cc;
      double v = -INFINITY;
      if (nn < N) {
        v = 0.0;
#pragma unroll
        for (int q = 0; q < ST_KG; ++q) v += Ps[(q * MT + m) * ST_COLS + cc];
        if (bias) v += (double)bias[nn];
        if (C) C[(int64_t)m * ldc + nn] = v;
      }
      Ps[m * ST_COLS + cc] = v;                                       // slice 0's slot of (m, cc): read by no other thread before the sync
    }
    if (part) {
      __syncthreads();
      for (int m = g; m < M; m += ST_KG) {                            // one warp per row: 64 columns, 2 per lane
        const double v0 = Ps[m * ST_COLS + lane], v1 = Ps[m * ST_COLS + 32 + lane];
        double mx = fmax(v0, v1);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        double sm = (v0 == -INFINITY ? 0.0 : exp(v0 - mx)) + (v1 == -INFINITY ? 0.0 : exp(v1 - mx));
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
        if (lane == 0) part[(int64_t)m * part_ld + part_tile0 + tile] = make_double2(mx, sm);
      }
    }
    __syncthreads();                                                  // Ps and Wt[buf] are free for the next round
  }
}

// Small-N float64 GEMM for few rows (the gate pre-activations, N = 4H, and the stage-1 projection, N = E): ONE WARP PER
// OUTPUT COLUMN.  k_skinny_f64's column-per-lane layout gives N / 32 CTAs - 64 for the gate GEMM, 8 for stage-1 - on a
// 148-SM part, and both were pure latency (15 and 13 us per launch for 6 MB and 1 MB of weights).  Here the lanes of a
// warp split K with 16-byte loads (a warp reads 512 contiguous bytes of its weight row), every lane keeps one
// float64 accumulator per row against its own K slice of A (read through L1: the same few KB for every warp of the
// SM), and the rows are reduced across the warp at the end.
template <typename TB, int MT>
__global__ void __launch_bounds__(256)
k_warpcol_f64(const double* __restrict__ A, int lda, const TB* __restrict__ B, int ldb, const float* __restrict__ bias,
              double* __restrict__ C, int64_t ldc, int M, int N, int K) {
  pdl_enter();
  constexpr int VEC = 16 / (int)sizeof(TB);
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  const TB* brow = B + (int64_t)n * ldb;
  double acc[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) acc[m] = 0.0;
#pragma unroll 4      // the iterations are independent: their weight and A loads overlap instead of paying one L2 round trip each
  for (int k = lane * VEC; k < K; k += 32 * VEC) {
    double w[VEC];
    SkLoad<TB>::load(brow + k, nullptr, w);
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      if (m < M) {
        const double* ap = A + (int64_t)m * lda + k;
#pragma unroll
        for (int i = 0; i < VEC; i += 2) {
          const double2 a2 = __ldg(reinterpret_cast<const double2*>(ap + i));
          acc[m] = fma(a2.x, w[i], acc[m]);
          acc[m] = fma(a2.y, w[i + 1], acc[m]);
        }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < MT; ++m) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], o);
  }
  if (lane == 0) {
    const double bv = bias ? (double)bias[n] : 0.0;
#pragma unroll
    for (int m = 0; m < MT; ++m)
      if (m < M) C[(int64_t)m * ldc + n] = acc[m] + bv;
  }
}

__global__ void k_lse_merge(const double2* __restrict__ part, int part_ld, int n_tiles, int M,
                            double* __restrict__ lse, int self_norm) {
  pdl_enter();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= M) return;
  if (self_norm) {
    if (lane == 0) lse[warp] = 0.0;
    return;
  }
  const double2* p = part + (int64_t)warp * part_ld;
  double mx = -INFINITY;
  for (int t = lane; t < n_tiles; t += 32) mx = fmax(mx, p[t].x);
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  double s = 0.0;
  for (int t = lane; t < n_tiles; t += 32) s += p[t].y * exp(p[t].x - mx);
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) lse[warp] = mx + log(s);
}

__global__ void k_rows_lse(const double* __restrict__ y, int64_t ld, int M, int N, double* __restrict__ lse) {
  pdl_enter();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= M) return;
  const double* p = y + (int64_t)warp * ld;
  double mx = -INFINITY;
  for (int t = lane; t < N; t += 32) mx = fmax(mx, p[t]);
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  double s = 0.0;
  for (int t = lane; t < N; t += 32) s += exp(p[t] - mx);
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) lse[warp] = mx + log(s);
}

__global__ void k_softmax_rows(const double* __restrict__ y, int64_t ld, int M, int N,
                               const double* __restrict__ lse, double* __restrict__ pred) {
  pdl_enter();
  const int64_t total = (int64_t)M * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / N);
    const int n = (int)(i % N);
    const double v = y[(int64_t)m * ld + n];
    pred[i] = lse ? exp(v - lse[m]) : exp(v);
  }
}

__global__ void k_gather_gate_input(const double* __restrict__ h_src, const int32_t* __restrict__ parent,
                                    const int32_t* __restrict__ word, const float* __restrict__ LM_in, int Hp,
                                    int Ep, int M, double* __restrict__ A) {
  pdl_enter();
  const int Kg = Hp + Ep;
  const int64_t total = (int64_t)M * Kg;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / Kg);
    const int k = (int)(i % Kg);
    double v;
    if (k < Hp) {
      const int p = parent ? parent[m] : m;
      v = (p >= 0) ? h_src[(int64_t)p * Hp + k] : 0.0;
    } else {
      v = (double)LM_in[(int64_t)word[m] * Ep + (k - Hp)];
    }
    A[i] = v;
  }
}

// decoder/model.py:132-139: i,f,o = sigmoid ; g = tanh ; c = c*f + g*i ; h = tanh(c)*o
__global__ void k_lstm_pointwise(const double* __restrict__ gates, const double* __restrict__ c_src,
                                 const int32_t* __restrict__ parent, int H, int Hp, int M,
                                 double* __restrict__ h_out, double* __restrict__ c_out) {
  pdl_enter();
  const int64_t total = (int64_t)M * Hp;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(idx / Hp);
    const int j = (int)(idx % Hp);
    if (j >= H) {
      h_out[idx] = 0.0;
      c_out[idx] = 0.0;
      continue;
    }
    const double* g = gates + (int64_t)m * 4 * H;
    const double gi = 1.0 / (exp(-g[j]) + 1.0);
    const double gf = 1.0 / (exp(-g[H + j]) + 1.0);
    const double go = 1.0 / (exp(-g[2 * H + j]) + 1.0);
    const double gg = tanh(g[3 * H + j]);
    const int p = parent ? parent[m] : m;
    const double cp = (p >= 0) ? c_src[(int64_t)p * Hp + j] : 0.0;
    const double c = cp * gf + gg * gi;
    c_out[idx] = c;
    h_out[idx] = tanh(c) * go;
  }
}

// Logits of a word subset for the rows of a job (LSTM_Model.project with `vocab`, decoder/model.py:
// 144-193): out[job.out0 + r*ncols + j] = T[row0+r] . W[cols[j]] + b2[bias_idx[j]], float64
// accumulation.  One CTA = 128 columns of one job: thread = column (its word row of the output block
// staged in shared memory, K in chunks of 32), all rows of the job accumulate in registers against
// broadcast reads of the shared T chunk - no cross-lane reduction.  Columns of different segments
// (different K slice of T) are handled in separate passes.
constexpr int VL_COLS = 128;    // columns (= threads) per CTA
constexpr int VL_KC = 32;       // K chunk (float4 per row: VL_KC/4)
constexpr int VL_RC = 32;       // rows accumulated per pass

template <typename TT>
__global__ void __launch_bounds__(VL_COLS)
k_vocab_logits(SegTable seg, const TT* __restrict__ T, int64_t ldt, const SubsetJob* __restrict__ jobs,
               const int32_t* __restrict__ cols, const int32_t* __restrict__ bias_idx,
               const float* __restrict__ b2, double* __restrict__ out) {
  pdl_enter();
  __shared__ __align__(16) double Ts[VL_RC][VL_KC];
  __shared__ __align__(16) float Ws[VL_COLS][VL_KC + 2];
  __shared__ int32_t Wid[VL_COLS];
  const SubsetJob job = jobs[blockIdx.y];
  const int c0 = blockIdx.x * VL_COLS;
  if (c0 >= job.ncols) return;
  const int tid = threadIdx.x;
  const int j = c0 + tid;
  const bool valid = j < job.ncols;
  const int w = valid ? cols[job.col0 + j] : -1;
  int myseg = -1;
  if (valid) {
    myseg = 0;
#pragma unroll
    for (int i = 1; i < JLM_MAX_SEGMENTS; ++i)
      if (i < seg.n && w >= seg.start[i]) myseg = i;
  }
  Wid[tid] = w;
  const double bias = valid ? (double)b2[bias_idx ? bias_idx[job.col0 + j] : w] : 0.0;
  for (int r0 = 0; r0 < job.rows; r0 += VL_RC) {
    const int nr = min(VL_RC, job.rows - r0);
    double acc[VL_RC];
#pragma unroll
    for (int r = 0; r < VL_RC; ++r) acc[r] = 0.0;
    for (int s = 0; s < seg.n; ++s) {
      if (!__syncthreads_or(myseg == s)) continue;
      const int kpad = seg.kpad[s];
      const int sstart = seg.start[s], send = seg.end[s];
      const float* Wseg = seg.W[s];
      for (int k0 = 0; k0 < kpad; k0 += VL_KC) {
        for (int i = tid; i < nr * VL_KC; i += VL_COLS) {
          const int r = i / VL_KC, k = i % VL_KC;
          Ts[r][k] = (double)T[(job.row0 + r0 + r) * ldt + seg.koff[s] + k0 + k];
        }
        constexpr int QPR = VL_KC / 4;              // float4 per staged row
        constexpr int RPI = VL_COLS / QPR;          // rows staged per iteration
#pragma unroll 4
        for (int it = 0; it < VL_COLS / RPI; ++it) {
          const int c = it * RPI + tid / QPR, q = tid % QPR;
          const int wc = Wid[c];
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (wc >= sstart && wc < send)
            v = *reinterpret_cast<const float4*>(Wseg + (int64_t)(wc - sstart) * kpad + k0 + q * 4);
          float2* dst = reinterpret_cast<float2*>(&Ws[c][q * 4]);
          dst[0] = make_float2(v.x, v.y);
          dst[1] = make_float2(v.z, v.w);
        }
        __syncthreads();
        if (myseg == s) {
#pragma unroll 2
          for (int k = 0; k < VL_KC; k += 2) {
            const float2 wv = *reinterpret_cast<const float2*>(&Ws[tid][k]);
            const double w0 = (double)wv.x, w1 = (double)wv.y;
#pragma unroll
            for (int r = 0; r < VL_RC; ++r) {
              if (r < nr) {
                const double2 t = *reinterpret_cast<const double2*>(&Ts[r][k]);
                acc[r] = fma(t.x, w0, acc[r]);
                acc[r] = fma(t.y, w1, acc[r]);
              }
            }
          }
        }
        __syncthreads();
      }
    }
    if (valid) {
#pragma unroll
      for (int r = 0; r < VL_RC; ++r)
        if (r < nr) out[job.out0 + (int64_t)(r0 + r) * job.ncols + j] = acc[r] + bias;
    }
  }
}

int grid_1d(int64_t total, int block, int cap) {
  int64_t g = (total + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

static bool warpcol_enabled() {
  static const int v = [] {
    const char* e = getenv("JLM_WARPCOL");
    return e ? atoi(e) : 1;
  }();
  return v != 0;
}

template <typename TB, int MT>
int32_t launch_skinny(cudaStream_t st, const double* A, int lda, const TB* B, int ldb, const float* codebook,
                      const float* bias, double* C, int64_t ldc, int M, int N, int K, double2* part, int part_ld,
                      int part_tile0) {
  static bool configured = false;
  if (!configured) {
    JLM_CUDA(cudaFuncSetAttribute(k_skinny_f64<TB, MT, 64, 2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    JLM_CUDA(cudaFuncSetAttribute(k_skinny_f64<TB, MT, 32, 1, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    // three 72 KB CTAs per SM need the whole shared-memory carve-out (ncu showed two resident: the kernel is
    // latency-bound at 3.5 warps per scheduler)
    JLM_CUDA(cudaFuncSetAttribute(k_skinny_f64<TB, MT, 64, 2, 8>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  cudaSharedmemCarveoutMaxShared));
    configured = true;
  }
  constexpr int VEC = 16 / (int)sizeof(TB);
  if (part || N > 8192 || sizeof(TB) == 1 || K % (16 * VEC) != 0) {
    const size_t smem = ((size_t)M * K + (size_t)8 * MT * 64 + 256) * sizeof(double);
    JLM_CUDA(jlm_launch(k_skinny_f64<TB, MT, 64, 2, 8>, dim3(ceil_div(N, 64)), dim3(256), smem, st, A, lda, B, ldb, codebook, bias, C, ldc, M, N, K, part, part_ld, part_tile0));
  } else if (warpcol_enabled() && C && K % (32 * VEC) == 0) {
    JLM_CUDA(jlm_launch(k_warpcol_f64<TB, MT>, dim3(ceil_div(N, 8)), dim3(256), 0, st, A, lda, B, ldb, bias, C, ldc, M, N, K));
  } else {
    const size_t smem = ((size_t)M * K + (size_t)16 * MT * 32 + 256) * sizeof(double);
    JLM_CUDA(jlm_launch(k_skinny_f64<TB, MT, 32, 1, 16>, dim3(ceil_div(N, 32)), dim3(512), smem, st, A, lda, B, ldb, codebook, bias, C, ldc, M, N, K, nullptr, 0, 0));
  }
  JLM_CUDA(cudaGetLastError());
  return 0;
}

template <typename TB>
int32_t skinny_dispatch(cudaStream_t st, const double* A, int lda, const TB* B, int ldb, const float* codebook,
                        const float* bias, double* C, int64_t ldc, int M, int N, int K, double2* part, int part_ld,
                        int part_tile0) {
  if (M <= 4) return launch_skinny<TB, 4>(st, A, lda, B, ldb, codebook, bias, C, ldc, M, N, K, part, part_ld, part_tile0);
  if (M <= 8) return launch_skinny<TB, 8>(st, A, lda, B, ldb, codebook, bias, C, ldc, M, N, K, part, part_ld, part_tile0);
  if (M <= 12) return launch_skinny<TB, 12>(st, A, lda, B, ldb, codebook, bias, C, ldc, M, N, K, part, part_ld, part_tile0);
  return launch_skinny<TB, 16>(st, A, lda, B, ldb, codebook, bias, C, ldc, M, N, K, part, part_ld, part_tile0);
}

// the skinny path needs M <= 16, K a multiple of 8 slices x 16 weights, and A + partial sums in shared memory
bool skinny_ok(int M, int K) {
  return M >= 1 && M <= 16 && K % 128 == 0 && ((size_t)M * K + 16 * 16 * 32 + 256) * sizeof(double) <= 200 * 1024;
}

// k_stream_f64 for up to 16 rows; more rows go through it 16 at a time (the weight block is L2-resident: re-streaming
// it per chunk beats the register-tiled k_gemm_f64 up to a few hundred rows)
template <int MT>
int32_t launch_stream_mt(cudaStream_t st, const double* A, int lda, const float* B, int ldb, const float* bias, double* C,
                         int64_t ldc, int M, int N, int K, double2* part, int part_ld, int part_tile0, int sm_count) {
  // eight K slices (eight warps) while their partial sums fit beside the two weight tiles and A; four for 16 rows
  constexpr int KG = (MT <= 12) ? 8 : 4;
  static bool configured = false;
  if (!configured) {
    JLM_CUDA(cudaFuncSetAttribute(k_stream_f64<MT, KG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    configured = true;
  }
  const size_t smem = (size_t)2 * ST_COLS * K * 4 + ((size_t)M * K + (size_t)KG * MT * ST_COLS) * sizeof(double);
  const int grid = std::min(ceil_div(N, ST_COLS), sm_count);
  JLM_CUDA(jlm_launch(k_stream_f64<MT, KG>, dim3(grid), dim3(32 * KG), smem, st, A, lda, B, ldb, bias, C, ldc, M, N, K, part, part_ld, part_tile0));
  JLM_CUDA(cudaGetLastError());
  return 0;
}

bool stream_ok(int M, int N, int K) {
  static const int on = [] {
    const char* e = getenv("JLM_STREAM_GEMM");
    return e ? atoi(e) : 1;
  }();
  // Up to 16 rows k_skinny_f64 is still the faster of the two (measured, single sentence at cfg 2: 43.5 vs 52.7 us per
  // launch, profiles/r02/launch_summary_single_*.txt); beyond that this kernel runs 16 rows at a time.
  // JLM_STREAM_GEMM=2 also routes <= 16 rows here (A/B runs)
  return on && (M > 16 || on == 2) && M <= 512 && K <= 256 && K % 64 == 0 && N >= 1024;      // K / 8 slices of whole 16-byte pieces
}

int32_t launch_stream(cudaStream_t st, const double* A, int lda, const float* B, int ldb, const float* bias, double* C,
                      int64_t ldc, int M, int N, int K, double2* part, int part_ld, int part_tile0) {
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count <= 0) sm_count = 148;
  }
  for (int m0 = 0; m0 < M; m0 += 16) {
    const int mc = std::min(16, M - m0);
    const double* Ac = A + (int64_t)m0 * lda;
    double* Cc = C ? C + (int64_t)m0 * ldc : nullptr;
    double2* pc = part ? part + (int64_t)m0 * part_ld : nullptr;
    if (mc <= 4) JLM_TRY(launch_stream_mt<4>(st, Ac, lda, B, ldb, bias, Cc, ldc, mc, N, K, pc, part_ld, part_tile0, sm_count));
    else if (mc <= 8) JLM_TRY(launch_stream_mt<8>(st, Ac, lda, B, ldb, bias, Cc, ldc, mc, N, K, pc, part_ld, part_tile0, sm_count));
    else if (mc <= 12) JLM_TRY(launch_stream_mt<12>(st, Ac, lda, B, ldb, bias, Cc, ldc, mc, N, K, pc, part_ld, part_tile0, sm_count));
    else JLM_TRY(launch_stream_mt<16>(st, Ac, lda, B, ldb, bias, Cc, ldc, mc, N, K, pc, part_ld, part_tile0, sm_count));
  }
  return 0;
}

template <typename TB>
int32_t launch_gemm(cudaStream_t st, const double* A, int lda, const TB* B, int ldb, const float* bias, double* C,
                    int64_t ldc, int M, int N, int K, double2* part, int part_ld, int part_tile0) {
  JLM_REQUIRE(K % BK == 0 && lda % 2 == 0 && ldb % 4 == 0, "exact gemm: unaligned K=%d lda=%d ldb=%d", K, lda, ldb);
  if (M <= 0 || N <= 0) return 0;
  if constexpr (sizeof(TB) == 4) {
    if (stream_ok(M, N, K))
      return launch_stream(st, A, lda, reinterpret_cast<const float*>(B), ldb, bias, C, ldc, M, N, K, part, part_ld, part_tile0);
  }
  if (skinny_ok(M, K)) {
    // weight-streaming path for one sentence's rows
    JLM_TRY(skinny_dispatch<TB>(st, A, lda, B, ldb, nullptr, bias, C, ldc, M, N, K, part, part_ld, part_tile0));
  } else if (M <= 16) {
    dim3 grid(ceil_div(N, BN), ceil_div(M, 16));
    JLM_CUDA(jlm_launch(k_gemm_f64<TB, 1>, dim3(grid), dim3(256), 0, st, A, lda, B, ldb, bias, C, ldc, M, N, K, part, part_ld, part_tile0));
  } else if (M <= 32) {
    dim3 grid(ceil_div(N, BN), ceil_div(M, 32));
    JLM_CUDA(jlm_launch(k_gemm_f64<TB, 2>, dim3(grid), dim3(256), 0, st, A, lda, B, ldb, bias, C, ldc, M, N, K, part, part_ld, part_tile0));
  } else {
    dim3 grid(ceil_div(N, BN), ceil_div(M, 64));
    JLM_CUDA(jlm_launch(k_gemm_f64<TB, 4>, dim3(grid), dim3(256), 0, st, A, lda, B, ldb, bias, C, ldc, M, N, K, part, part_ld, part_tile0));
  }
  JLM_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

int exact_tiles_n(int N) { return ceil_div(N, BN); }

int32_t exact_gemm_f32w(cudaStream_t st, const double* A, int lda, const float* B, int ldb, const float* bias,
                        double* C, int64_t ldc, int M, int N, int K, double2* part, int part_ld,
                        int part_tile0) {
  return launch_gemm<float>(st, A, lda, B, ldb, bias, C, ldc, M, N, K, part, part_ld, part_tile0);
}

bool exact_use_q8(const jlm_handle* h, const SegDev& s, int M) {
  if (!s.Wq || !skinny_ok(M, s.kpad) || h->q8_policy == 0) return false;
  if (h->q8_policy == 1) return true;
  return (size_t)(s.end - s.start) * s.kpad * sizeof(float) > ((size_t)96 << 20);
}

int32_t exact_gemm_q8w(cudaStream_t st, const double* A, int lda, const uint8_t* Bq, int ldb, const float* codebook,
                       const float* bias, double* C, int64_t ldc, int M, int N, int K, double2* part, int part_ld,
                       int part_tile0) {
  JLM_REQUIRE(skinny_ok(M, K) && ldb % 16 == 0, "q8 gemm: unsupported shape M=%d K=%d ldb=%d", M, K, ldb);
  if (N <= 0) return 0;
  return skinny_dispatch<uint8_t>(st, A, lda, Bq, ldb, codebook, bias, C, ldc, M, N, K, part, part_ld, part_tile0);
}

int32_t exact_gemm_f64w(cudaStream_t st, const double* A, int lda, const double* B, int ldb, double* C,
                        int64_t ldc, int M, int N, int K) {
  return launch_gemm<double>(st, A, lda, B, ldb, nullptr, C, ldc, M, N, K, nullptr, 0, 0);
}

int32_t exact_lse_merge(cudaStream_t st, const double2* part, int part_ld, int n_tiles, int M, double* lse,
                        int self_norm) {
  if (M <= 0) return 0;
  JLM_CUDA(jlm_launch(k_lse_merge, dim3(ceil_div((int64_t)M * 32, 256)), dim3(256), 0, st, part, part_ld, n_tiles, M, lse, self_norm));
  JLM_CUDA(cudaGetLastError());
  return 0;
}

int32_t exact_rows_lse(cudaStream_t st, const double* y, int64_t ld, int M, int N, double* lse) {
  if (M <= 0) return 0;
  JLM_CUDA(jlm_launch(k_rows_lse, dim3(ceil_div((int64_t)M * 32, 256)), dim3(256), 0, st, y, ld, M, N, lse));
  JLM_CUDA(cudaGetLastError());
  return 0;
}

int32_t exact_softmax_rows(cudaStream_t st, const double* y, int64_t ld, int M, int N, const double* lse,
                           double* pred) {
  if (M <= 0 || N <= 0) return 0;
  JLM_CUDA(jlm_launch(k_softmax_rows, dim3(grid_1d((int64_t)M * N, 256, 148 * 16)), dim3(256), 0, st, y, ld, M, N, lse, pred));
  JLM_CUDA(cudaGetLastError());
  return 0;
}

int32_t exact_gather_gate_input(cudaStream_t st, const jlm_handle* h, const double* h_src, const int32_t* parent,
                                const int32_t* word, int M, double* A) {
  if (M <= 0) return 0;
  JLM_CUDA(jlm_launch(k_gather_gate_input, dim3(grid_1d((int64_t)M * h->Kg, 256, 148 * 16)), dim3(256), 0, st, h_src, parent, word, h->LM_in, h->Hp, h->Ep, M, A));
  JLM_CUDA(cudaGetLastError());
  return 0;
}

int32_t exact_lstm_pointwise(cudaStream_t st, const jlm_handle* h, const double* gates, const double* c_src,
                             const int32_t* parent, int M, double* h_out, double* c_out) {
  if (M <= 0) return 0;
  JLM_CUDA(jlm_launch(k_lstm_pointwise, dim3(grid_1d((int64_t)M * h->Hp, 256, 148 * 16)), dim3(256), 0, st, gates, c_src, parent, h->H, h->Hp, M, h_out, c_out));
  JLM_CUDA(cudaGetLastError());
  return 0;
}

template <typename TT>
int32_t subset_logits(cudaStream_t st, const jlm_handle* h, const TT* T, int64_t ldt, const SubsetJob* jobs,
                      int n_jobs, int max_cols, const int32_t* cols, const int32_t* bias_idx, double* out) {
  if (n_jobs <= 0 || max_cols <= 0) return 0;
  const int gx = ceil_div(max_cols, VL_COLS);
  for (int j0 = 0; j0 < n_jobs; j0 += 65535) {
    const int nj = n_jobs - j0 < 65535 ? n_jobs - j0 : 65535;
    dim3 grid(gx, nj);
    JLM_CUDA(jlm_launch(k_vocab_logits<TT>, dim3(grid), dim3(VL_COLS), 0, st, make_seg_table(h), T, ldt, jobs + j0, cols, bias_idx, h->b2, out));
  }
  JLM_CUDA(cudaGetLastError());
  return 0;
}

template int32_t subset_logits<double>(cudaStream_t, const jlm_handle*, const double*, int64_t, const SubsetJob*, int,
                                       int, const int32_t*, const int32_t*, double*);
template int32_t subset_logits<float>(cudaStream_t, const jlm_handle*, const float*, int64_t, const SubsetJob*, int,
                                      int, const int32_t*, const int32_t*, double*);
