// Internal declarations shared by the translation units of libjlm_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/jlm_b200.h"

void jlm_set_error(const char* fmt, ...);

#define JLM_CUDA(x)                                                                          \
  do {                                                                                       \
    cudaError_t e__ = (x);                                                                   \
    if (e__ != cudaSuccess) {                                                                \
      jlm_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #x, cudaGetErrorString(e__));      \
      return 1;                                                                              \
    }                                                                                        \
  } while (0)

#define JLM_REQUIRE(cond, ...)      \
  do {                              \
    if (!(cond)) {                  \
      jlm_set_error(__VA_ARGS__);   \
      return 1;                     \
    }                               \
  } while (0)

#define JLM_TRY(x)            \
  do {                        \
    int32_t r__ = (x);        \
    if (r__) return r__;      \
  } while (0)

static inline int64_t round_up64(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------- programmatic dependent launch
// Every kernel of the library starts with pdl_enter(): it lets the NEXT kernel of the stream be scheduled as soon as all
// CTAs of this one are running (griddepcontrol.launch_dependents) and then waits until everything the PREVIOUS kernels
// wrote is visible (griddepcontrol.wait) before it touches global memory.  With jlm_launch() (below) the launch
// latency, the CTA ramp and - for the GEMM kernels, which call it after their barrier / TMEM set-up - the prologue of
// kernel k+1 hide under the tail of kernel k; a lock-step frame is 7-9 dependent launches.  Both instructions are
// no-ops for a kernel launched without the attribute.  A kernel that allocates TMEM calls pdl_enter() AFTER the
// allocation: a dependent CTA that became co-resident and took the columns first would wait forever on a
// predecessor that cannot get them.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// Every CTA of a kernel must pass a wait before it exits (also one that has nothing to do): a grid that finished
// without waiting would let ITS dependent start reading what the grid before it is still writing.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
  pdl_trigger();
  pdl_wait();
}
#endif

// Launches made while a PdlOff object lives on this thread are plain (see jlm_batch_run for when).
inline thread_local int jlm_pdl_off_depth = 0;
struct PdlOff {
  bool on;
  explicit PdlOff(bool off) : on(off) { if (on) ++jlm_pdl_off_depth; }
  ~PdlOff() { if (on) --jlm_pdl_off_depth; }
  PdlOff(const PdlOff&) = delete;
  PdlOff& operator=(const PdlOff&) = delete;
};

static inline bool jlm_pdl_enabled() {
  static const int v = [] {
    const char* e = getenv("JLM_PDL");
    return e ? atoi(e) : 1;
  }();
  return v != 0 && jlm_pdl_off_depth == 0;
}

// kernel<<<grid, block, smem, st>>>(args...) with programmatic stream serialization (JLM_PDL=0: a plain launch)
template <typename... KArgs, typename... Args>
static inline cudaError_t jlm_launch_opt(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                         cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && jlm_pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename... KArgs, typename... Args>
static inline cudaError_t jlm_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = jlm_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

constexpr int JLM_KALIGN = 64;  // every GEMM K extent on the device is padded to this (zero filled)

// Grow-only device / pinned-host buffers, reused across calls so the steady state allocates nothing.
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int32_t reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    JLM_CUDA(cudaMalloc(&p, want));
    cap = want;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

struct HostBuf {
  void* p = nullptr;
  size_t cap = 0;
  int32_t reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    JLM_CUDA(cudaMallocHost(&p, want));
    cap = want;
    return 0;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

// Bump allocator over one DevBuf: plan sizes first (dry run), reserve once, then hand out pointers.
struct Arena {
  DevBuf buf;
  size_t off = 0;
  bool dry = true;
  void begin_plan() { off = 0; dry = true; }
  int32_t commit() {
    JLM_TRY(buf.reserve(off));
    off = 0;
    dry = false;
    return 0;
  }
  template <class T>
  T* take(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* r = dry ? nullptr : reinterpret_cast<T*>(static_cast<char*>(buf.p) + off);
    off += n * sizeof(T);
    return r;
  }
};

struct SegDev {
  int width;   // e_i: K of the segment's output GEMM
  int kpad;    // width rounded up to JLM_KALIGN
  int koff;    // column offset of the segment's slice inside the stage-1 output T (padded layout)
  int start, end;
  const float* W;   // [end-start, kpad] K-major, zero padded (exact back end + needed-word dots)
  const uint8_t* Wq = nullptr;   // optional 8-bit codes, same layout (pad columns hold code 0; A is zero there)
  const float* cb = nullptr;     // its 256-entry codebook (unused entries zero)
};

struct TcWeights;  // tensor-core operand copies (jlm_tc.cu)

struct jlm_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t side_stream = nullptr;   // needed-word dot products run here while the output GEMM holds `stream`
  cudaStream_t copy_stream[2] = {};     // [0] plan H2D, [1] n-best D2H of a batch: beside the previous / next batch's kernels
  int sm_count = 148;
  jlm_config cfg{};
  int V = 0, H = 0, E = 0;
  int Hp = 0, Ep = 0;   // padded to JLM_KALIGN
  int Kg = 0;           // gate GEMM K = Hp + Ep
  int Kt = 0;           // stage-1 output width (sum of kpad); untied: Hp (T aliases h)
  int n_seg = 0;
  SegDev seg[JLM_MAX_SEGMENTS];
  bool untied = false;
  // device weights (exact layouts)
  float* Wg = nullptr;      // [4H, Kg]  row n = gate*H + j ; k < Hp: HM, k >= Hp: IM
  float* bg = nullptr;      // [4H]
  float* b2 = nullptr;      // [V]
  float* LM_in = nullptr;   // [V, Ep] zero padded
  double* P1 = nullptr;     // [Kt, Hp] stage-1 weight, K-major, float64 (NULL when untied)
  float* Wseg_store[JLM_MAX_SEGMENTS] = {};
  uint8_t* Wq_store[JLM_MAX_SEGMENTS] = {};
  float* cb_store[JLM_MAX_SEGMENTS] = {};
  int q8_policy = -1;   // -1 auto, 0 never, 1 always (JLM_Q8)
  double guard_eps = JLM_GUARD_EPS_DEFAULT;   // near-tie guard bound of tensor-core batches (jlm_set_guard, JLM_GUARD_EPS)
  bool guard_all = false;                     // JLM_GUARD_ALL / jlm_set_guard_scope: every rank decision of every frame
  bool guard_verify = true;                   // tier 1 (re-score the near-tied pairs) before tier 2 (re-decode the sentence)
  jlm_pool* guard_pool = nullptr;             // float64 state pool the guard re-scores near-tied paths with
  cudaStream_t guard_stream = nullptr;        // ... on its own stream, beside the next batch's kernels
  int batches_unfetched = 0;                  // tensor-core batches run but not yet fetched (a streaming caller keeps several)
  int tier2_recent = 0;                       // > 0: one of the last batches needed a float64 re-decode (launch chaining policy)
  void* plan_scratch = nullptr;   // host vectors of the last batch plan, reused by the next upload (jlm_beam.cu)
  TcWeights* tc = nullptr;
  // scratch for the model-level API and the batch engine
  DevBuf scratch[8];
  // device arenas of destroyed batches, reused by later uploads (several batches can be in flight when
  // a call is pipelined in chunks; the steady state allocates nothing)
  std::vector<DevBuf> batch_cache;
  HostBuf pinned[4];
  std::vector<HostBuf> out_pool;   // pinned n-best landing buffers of finished batches
  // pinned staging ring for the plan upload: a slot is reused only after its H2D copy completed
  static constexpr int N_STAGE = 4;
  HostBuf stage[N_STAGE];
  cudaEvent_t stage_ev[N_STAGE] = {};
  bool stage_busy[N_STAGE] = {};
  int stage_next = 0;
  cudaEvent_t ev[4] = {};
};

// Segment table passed by value to kernels that look a word's output block up.
struct SegTable {
  int n;
  int start[JLM_MAX_SEGMENTS], end[JLM_MAX_SEGMENTS], koff[JLM_MAX_SEGMENTS], kpad[JLM_MAX_SEGMENTS];
  const float* W[JLM_MAX_SEGMENTS];
};
static inline SegTable make_seg_table(const jlm_handle* h) {
  SegTable t{};
  t.n = h->n_seg;
  for (int i = 0; i < h->n_seg; ++i) {
    t.start[i] = h->seg[i].start;
    t.end[i] = h->seg[i].end;
    t.koff[i] = h->seg[i].koff;
    t.kpad[i] = h->seg[i].kpad;
    t.W[i] = h->seg[i].W;
  }
  return t;
}

// ---------------------------------------------------------------- exact back end (jlm_exact.cu)
// C[M,N] (+bias) = A[M,K] . B[N,K]^T, float64 accumulation.  K % 32 == 0, lda/ldb % 4 == 0.
// part != nullptr: also emit per-(row, 64-column tile) (max, sum exp) partials, tile index offset
// part_tile0, row stride part_ld (in tiles).  C may be nullptr when only partials are wanted.
int32_t exact_gemm_f32w(cudaStream_t st, const double* A, int lda, const float* B, int ldb, const float* bias,
                        double* C, int64_t ldc, int M, int N, int K, double2* part, int part_ld,
                        int part_tile0);
// true when the 8-bit code path can and should serve this call: codes attached, shape supported by the
// streaming kernel, and (policy) the float32 block is too large to stay L2-resident between LM steps
// (JLM_Q8=1 forces the codes, JLM_Q8=0 disables them; default: blocks > 96 MB)
bool exact_use_q8(const jlm_handle* h, const SegDev& s, int M);
// same contraction with the weights given as 8-bit codes + codebook; M <= 16 only (weight streaming)
int32_t exact_gemm_q8w(cudaStream_t st, const double* A, int lda, const uint8_t* Bq, int ldb, const float* codebook,
                       const float* bias, double* C, int64_t ldc, int M, int N, int K, double2* part, int part_ld,
                       int part_tile0);
int32_t exact_gemm_f64w(cudaStream_t st, const double* A, int lda, const double* B, int ldb,
                        double* C, int64_t ldc, int M, int N, int K);
int exact_tiles_n(int N);
// lse[m] = log sum exp over the partial tiles of row m
int32_t exact_lse_merge(cudaStream_t st, const double2* part, int part_ld, int n_tiles, int M, double* lse,
                        int self_norm);
// lse[m] = log sum exp of the dense row y[m, 0..N)
int32_t exact_rows_lse(cudaStream_t st, const double* y, int64_t ld, int M, int N, double* lse);
// pred = exp(y - lse[m]) (softmax) ; lse == nullptr -> exp(y)
int32_t exact_softmax_rows(cudaStream_t st, const double* y, int64_t ld, int M, int N, const double* lse,
                           double* pred);
// A[m] = [ h_src[parent[m]] (Hp) | LM_in[word[m]] (Ep) ] as float64; parent < 0 -> zero state.
int32_t exact_gather_gate_input(cudaStream_t st, const jlm_handle* h, const double* h_src, const int32_t* parent,
                                const int32_t* word, int M, double* A);
// gates -> (h,c): c' = c*f + g*i ; h' = tanh(c')*o.  c_src rows gathered through parent (<0 -> 0).
int32_t exact_lstm_pointwise(cudaStream_t st, const jlm_handle* h, const double* gates, const double* c_src,
                             const int32_t* parent, int M, double* h_out, double* c_out);

// y[r, j] for selected words, row-major per job: out[job.out0 + r*job.ncols + j]
struct SubsetJob {
  int64_t row0;      // first T row
  int32_t rows;
  int64_t col0;      // first entry of cols/bias_idx
  int32_t ncols;
  int64_t out0;
};
template <typename TT>
int32_t subset_logits(cudaStream_t st, const jlm_handle* h, const TT* T, int64_t ldt, const SubsetJob* jobs,
                      int n_jobs, int max_cols, const int32_t* cols, const int32_t* bias_idx, double* out);

void beam_free_plan_scratch(jlm_handle* h);   // jlm_beam.cu
void beam_free_guard(jlm_handle* h);          // jlm_beam.cu: the near-tie verifier's state pool and stream

// ---------------------------------------------------------------- LM state pool internals (jlm_pool.cu)
int32_t pool_step_rows(jlm_pool* p, int32_t n, const int32_t* src, const int32_t* index, int64_t* first_slot, bool with_lse);
int32_t pool_lse_rows(jlm_pool* p, int64_t first, int64_t count);
int32_t pool_lse_slots(jlm_pool* p, const int32_t* slots, int32_t n);
int32_t pool_lse_subsets(jlm_pool* p, int32_t n, const int32_t* slots, const int64_t* col_ptr, const int32_t* cols, double* out);

// ---------------------------------------------------------------- tensor-core back end (jlm_tc.cu)
int32_t tc_prepare_weights(jlm_handle* h);
void tc_free_weights(jlm_handle* h);
