// LM state pool: LSTM_Model.predict_with_context (decoder/model.py:195-198) with the (hidden, cell) pairs kept on
// the device.  A host-driven search - the reference's CharRNNDecoder (decoder/decoder.py:244-341), whose control
// flow is string-keyed and per sentence - names states by slot: a step reads the states in `src` slots, feeds one
// word index per row and leaves the new states, their stage-1 projection rows and their softmax log-normaliser in
// fresh consecutive slots; jlm_pool_nll then returns -log p(word | state) for (slot, word) pairs.  Per call only
// indices go down and one float64 per asked pair comes back - no [B,H] states, no [B,V] distributions cross PCIe,
// which is what a predict_with_context round trip costs.  Float64 kernels of the exact back end throughout.
#include "jlm_common.cuh"

struct jlm_pool {
  jlm_handle* h = nullptr;
  int64_t cap = 0, used = 0;
  double* hx = nullptr;    // [cap, Hp]
  double* cx = nullptr;    // [cap, Hp]
  double* T = nullptr;     // [cap, Kt] stage-1 rows h.PM (tied) ; untied: aliases hx
  double* lse = nullptr;   // [cap]
  int ldt = 0;
  int tiles = 0;
  DevBuf A, G, part, idx, out;
  HostBuf out_host;
};

namespace {

// -log p(col | slot): y = T[slot] . W[col] + b2[col] in float64 (one warp per pair), lse[slot] - y
__global__ void __launch_bounds__(256)
k_pool_nll(SegTable seg, const double* __restrict__ T, int64_t ldt, const double* __restrict__ lse,
           const float* __restrict__ b2, const int32_t* __restrict__ slot, const int32_t* __restrict__ col, int n,
           int use_lse, double* __restrict__ out) {
  pdl_enter();
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const int w = col[i];
  const int64_t sl = slot[i];
  int s = 0;
#pragma unroll
  for (int k = 1; k < JLM_MAX_SEGMENTS; ++k)
    if (k < seg.n && w >= seg.start[k]) s = k;
  const int kpad = seg.kpad[s];
  const float* wrow = seg.W[s] + (int64_t)(w - seg.start[s]) * kpad;
  const double* trow = T + sl * ldt + seg.koff[s];
  double acc = 0.0;
  for (int k = lane; k < kpad; k += 32) acc = fma(trow[k], (double)wrow[k], acc);
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    const double y = acc + (double)b2[w];
    out[i] = use_lse ? lse[sl] - y : -y;
  }
}

__global__ void k_pool_gather_T(const double* __restrict__ T, int64_t ldt, const int32_t* __restrict__ slots, int n,
                                double* __restrict__ out) {
  pdl_enter();
  const int64_t total = (int64_t)n * ldt;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = T[(int64_t)slots[i / ldt] * ldt + i % ldt];
}

__global__ void k_pool_scatter_lse(const double* __restrict__ tmp, const int32_t* __restrict__ slots, int n,
                                   double* __restrict__ lse) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) lse[slots[i]] = tmp[i];
}

// log-sum-exp of a state's logits over an explicit word list (vocabulary-selection modes: the softmax runs over the
// sentence's lattice_vocab, decoder.py:137-151 / decoder_dynamic.py:93-148).  One CTA per request; a warp per word
// (float64 dot against the stage-1 row), online (max, sum) per warp, merged through shared memory.
__global__ void __launch_bounds__(128)
k_pool_lse_subset(SegTable seg, const double* __restrict__ T, int64_t ldt, const float* __restrict__ b2,
                  const int32_t* __restrict__ slots, const int64_t* __restrict__ col_ptr, const int32_t* __restrict__ cols,
                  double* __restrict__ out) {
  pdl_enter();
  __shared__ double wm[4], ws[4];
  const int r = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t c0 = col_ptr[r], c1 = col_ptr[r + 1];
  const double* trow0 = T + (int64_t)slots[r] * ldt;
  double mx = -INFINITY, sm = 0.0;
  for (int64_t c = c0 + warp; c < c1; c += 4) {
    const int w = cols[c];
    int sg = 0;
#pragma unroll
    for (int k = 1; k < JLM_MAX_SEGMENTS; ++k)
      if (k < seg.n && w >= seg.start[k]) sg = k;
    const int kpad = seg.kpad[sg];
    const float* wrow = seg.W[sg] + (int64_t)(w - seg.start[sg]) * kpad;
    const double* trow = trow0 + seg.koff[sg];
    double acc = 0.0;
    for (int k = lane; k < kpad; k += 32) acc = fma(trow[k], (double)wrow[k], acc);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    const double y = acc + (double)b2[w];
    const double nm = fmax(mx, y);
    sm = sm * exp(mx - nm) + exp(y - nm);
    mx = nm;
  }
  if (lane == 0) {
    wm[warp] = mx;
    ws[warp] = sm;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = fmax(fmax(wm[0], wm[1]), fmax(wm[2], wm[3]));
    double t = 0.0;
    for (int k = 0; k < 4; ++k)
      if (wm[k] > -INFINITY) t += ws[k] * exp(wm[k] - m);
    out[r] = m + log(t);
  }
}

}  // namespace

extern "C" int32_t jlm_pool_create(jlm_handle* h, int64_t capacity, jlm_pool** out) {
  JLM_REQUIRE(h && out && capacity > 0, "jlm_pool_create: bad argument");
  JLM_CUDA(cudaSetDevice(h->device));
  *out = nullptr;
  jlm_pool* p = new jlm_pool();
  p->h = h;
  p->cap = capacity;
  p->ldt = h->untied ? h->Hp : h->Kt;
  for (int i = 0; i < h->n_seg; ++i) p->tiles += exact_tiles_n(h->seg[i].end - h->seg[i].start);
  const size_t n = (size_t)capacity;
  cudaError_t e = cudaMalloc(&p->hx, n * h->Hp * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&p->cx, n * h->Hp * sizeof(double));
  if (e == cudaSuccess && !h->untied) e = cudaMalloc(&p->T, n * h->Kt * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&p->lse, n * sizeof(double));
  if (e != cudaSuccess) {
    jlm_set_error("jlm_pool_create: %lld states do not fit: %s", (long long)capacity, cudaGetErrorString(e));
    cudaGetLastError();
    jlm_pool_destroy(p);
    return 1;
  }
  if (h->untied) p->T = p->hx;
  *out = p;
  return 0;
}

extern "C" int32_t jlm_pool_destroy(jlm_pool* p) {
  if (!p) return 0;
  cudaSetDevice(p->h->device);
  cudaStreamSynchronize(p->h->stream);
  if (p->T && p->T != p->hx) cudaFree(p->T);
  cudaFree(p->hx);
  cudaFree(p->cx);
  cudaFree(p->lse);
  p->A.release();
  p->G.release();
  p->part.release();
  p->idx.release();
  p->out.release();
  p->out_host.release();
  delete p;
  return 0;
}

extern "C" int32_t jlm_pool_reset(jlm_pool* p) {
  JLM_REQUIRE(p, "jlm_pool_reset: null pool");
  p->used = 0;
  return 0;
}

// log-normalisers of the states in slots [first, first + count): full-vocabulary logits reduced to their log-sum-exp
int32_t pool_lse_rows(jlm_pool* p, int64_t first, int64_t count) {
  jlm_handle* h = p->h;
  if (h->cfg.self_norm || count <= 0) return 0;
  cudaStream_t st = h->stream;
  const double* Trow = p->T + first * p->ldt;
  const int n = (int)count;
  JLM_TRY(p->part.reserve(sizeof(double2) * (size_t)n * p->tiles));
  double2* part = p->part.as<double2>();
  int tile0 = 0;
  for (int i = 0; i < h->n_seg; ++i) {
    const SegDev& s = h->seg[i];
    const int Vi = s.end - s.start;
    JLM_TRY(exact_gemm_f32w(st, Trow + s.koff, p->ldt, s.W, s.kpad, h->b2 + s.start, nullptr, 0, n, Vi, s.kpad, part,
                            p->tiles, tile0));
    tile0 += exact_tiles_n(Vi);
  }
  JLM_TRY(exact_lse_merge(st, part, p->tiles, p->tiles, n, p->lse + first, 0));
  return 0;
}

// the same for a scattered list of slots (host array): their stage-1 rows are gathered into a dense block first
int32_t pool_lse_slots(jlm_pool* p, const int32_t* slots, int32_t n) {
  jlm_handle* h = p->h;
  if (h->cfg.self_norm || n <= 0) return 0;
  cudaStream_t st = h->stream;
  JLM_TRY(p->idx.reserve(sizeof(int32_t) * (size_t)n));
  JLM_TRY(p->G.reserve(sizeof(double) * ((size_t)n * p->ldt + (size_t)n)));
  JLM_TRY(p->part.reserve(sizeof(double2) * (size_t)n * p->tiles));
  int32_t* d_slots = p->idx.as<int32_t>();
  double* Tg = p->G.as<double>();
  double* lse_tmp = Tg + (size_t)n * p->ldt;
  JLM_CUDA(cudaMemcpyAsync(d_slots, slots, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
  JLM_CUDA(jlm_launch(k_pool_gather_T, dim3(ceil_div((int64_t)n * p->ldt, 256)), dim3(256), 0, st, p->T, p->ldt, d_slots, n, Tg));
  JLM_CUDA(cudaGetLastError());
  double2* part = p->part.as<double2>();
  int tile0 = 0;
  for (int i = 0; i < h->n_seg; ++i) {
    const SegDev& s = h->seg[i];
    const int Vi = s.end - s.start;
    JLM_TRY(exact_gemm_f32w(st, Tg + s.koff, p->ldt, s.W, s.kpad, h->b2 + s.start, nullptr, 0, n, Vi, s.kpad, part,
                            p->tiles, tile0));
    tile0 += exact_tiles_n(Vi);
  }
  JLM_TRY(exact_lse_merge(st, part, p->tiles, p->tiles, n, lse_tmp, 0));
  JLM_CUDA(jlm_launch(k_pool_scatter_lse, dim3(ceil_div(n, 256)), dim3(256), 0, st, lse_tmp, d_slots, n, p->lse));
  JLM_CUDA(cudaGetLastError());
  return 0;      // asynchronous: later users of p->idx / p->G enqueue behind these kernels on the same stream
}

// out[r] = log-sum-exp over the words cols[col_ptr[r] .. col_ptr[r+1]) of the logits of the state in slots[r] (host arrays)
int32_t pool_lse_subsets(jlm_pool* p, int32_t n, const int32_t* slots, const int64_t* col_ptr, const int32_t* cols, double* out) {
  jlm_handle* h = p->h;
  if (n <= 0) return 0;
  cudaStream_t st = h->stream;
  const int64_t nc = col_ptr[n];
  JLM_TRY(p->G.reserve(sizeof(int64_t) * (size_t)(n + 1) + sizeof(int32_t) * (size_t)(n + nc) + sizeof(double) * (size_t)n + 64));
  char* base = p->G.as<char>();
  int64_t* d_ptr = reinterpret_cast<int64_t*>(base);
  double* d_out = reinterpret_cast<double*>(base + sizeof(int64_t) * (size_t)(n + 1));
  int32_t* d_slots = reinterpret_cast<int32_t*>(d_out + n);
  int32_t* d_cols = d_slots + n;
  JLM_CUDA(cudaMemcpyAsync(d_ptr, col_ptr, sizeof(int64_t) * (size_t)(n + 1), cudaMemcpyHostToDevice, st));
  JLM_CUDA(cudaMemcpyAsync(d_slots, slots, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, st));
  JLM_CUDA(cudaMemcpyAsync(d_cols, cols, sizeof(int32_t) * (size_t)nc, cudaMemcpyHostToDevice, st));
  JLM_CUDA(jlm_launch(k_pool_lse_subset, dim3(n), dim3(128), 0, st, make_seg_table(h), p->T, p->ldt, h->b2, d_slots, d_ptr, d_cols, d_out));
  JLM_CUDA(cudaGetLastError());
  JLM_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
  JLM_CUDA(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int32_t jlm_pool_step(jlm_pool* p, int32_t n, const int32_t* src, const int32_t* index, int64_t* first_slot) {
  return pool_step_rows(p, n, src, index, first_slot, true);
}

// with_lse = false leaves the log-normalisers to a later pool_lse_rows call over a range of slots (one GEMM over
// many rows instead of one per step: the near-tie verifier, jlm_beam.cu)
int32_t pool_step_rows(jlm_pool* p, int32_t n, const int32_t* src, const int32_t* index, int64_t* first_slot, bool with_lse) {
  JLM_REQUIRE(p && src && index && first_slot && n > 0, "jlm_pool_step: bad argument");
  jlm_handle* h = p->h;
  JLM_REQUIRE(p->used + n <= p->cap, "jlm_pool_step: pool full (%lld used + %d > %lld)", (long long)p->used, n,
              (long long)p->cap);
  for (int i = 0; i < n; ++i) {
    JLM_REQUIRE(index[i] >= 0 && index[i] < h->V, "jlm_pool_step: index %d out of range", index[i]);
    JLM_REQUIRE(src[i] >= -1 && src[i] < p->used, "jlm_pool_step: source slot %d does not hold a state", src[i]);
  }
  JLM_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  JLM_TRY(p->idx.reserve(sizeof(int32_t) * 2 * (size_t)n));
  JLM_TRY(p->A.reserve(sizeof(double) * (size_t)n * h->Kg));
  JLM_TRY(p->G.reserve(sizeof(double) * (size_t)n * 4 * h->H));
  int32_t* d_src = p->idx.as<int32_t>();
  int32_t* d_idx = d_src + n;
  JLM_CUDA(cudaMemcpyAsync(d_src, src, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
  JLM_CUDA(cudaMemcpyAsync(d_idx, index, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
  const int64_t s0 = p->used;
  double* hrow = p->hx + s0 * h->Hp;
  double* crow = p->cx + s0 * h->Hp;
  // _lstm_cell (model.py:125-139): gather [h[src] | emb[index]], one float64 GEMM, pointwise gates
  JLM_TRY(exact_gather_gate_input(st, h, p->hx, d_src, d_idx, n, p->A.as<double>()));
  JLM_TRY(exact_gemm_f32w(st, p->A.as<double>(), h->Kg, h->Wg, h->Kg, h->bg, p->G.as<double>(), 4 * h->H, n, 4 * h->H,
                          h->Kg, nullptr, 0, 0));
  JLM_TRY(exact_lstm_pointwise(st, h, p->G.as<double>(), p->cx, d_src, n, hrow, crow));
  // project (model.py:141-193): stage-1 rows are kept (needed-word logits are dots against them), the full
  // vocabulary is only reduced to its log-sum-exp
  if (!h->untied) JLM_TRY(exact_gemm_f64w(st, hrow, h->Hp, h->P1, h->Hp, p->T + s0 * h->Kt, h->Kt, n, h->Kt, h->Hp));
  if (with_lse) JLM_TRY(pool_lse_rows(p, s0, n));
  else JLM_CUDA(cudaMemsetAsync(p->lse + s0, 0, sizeof(double) * (size_t)n, st));   // "no normaliser": jlm_pool_nll returns -y
  // Public entry: return with the step finished.  The verifier's chain of steps (with_lse = false) stays asynchronous:
  // the index copies above are in stream order behind the kernels that read the previous indices, the host arrays are
  // pageable (staged before cudaMemcpyAsync returns), and jlm_pool_nll synchronises at the end of the chain.
  if (with_lse) JLM_CUDA(cudaStreamSynchronize(st));
  p->used += n;
  *first_slot = s0;
  return 0;
}

extern "C" int32_t jlm_pool_nll(jlm_pool* p, int32_t n, const int32_t* slot, const int32_t* col, double* out) {
  JLM_REQUIRE(p && slot && col && out && n > 0, "jlm_pool_nll: bad argument");
  jlm_handle* h = p->h;
  for (int i = 0; i < n; ++i) {
    JLM_REQUIRE(slot[i] >= 0 && slot[i] < p->used, "jlm_pool_nll: slot %d does not hold a state", slot[i]);
    JLM_REQUIRE(col[i] >= 0 && col[i] < h->V, "jlm_pool_nll: word %d out of range", col[i]);
  }
  JLM_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  JLM_TRY(p->idx.reserve(sizeof(int32_t) * 2 * (size_t)n));
  JLM_TRY(p->out.reserve(sizeof(double) * (size_t)n));
  JLM_TRY(p->out_host.reserve(sizeof(double) * (size_t)n));
  int32_t* d_slot = p->idx.as<int32_t>();
  int32_t* d_col = d_slot + n;
  JLM_CUDA(cudaMemcpyAsync(d_slot, slot, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
  JLM_CUDA(cudaMemcpyAsync(d_col, col, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
  JLM_CUDA(jlm_launch(k_pool_nll, dim3(ceil_div((int64_t)n * 32, 256)), dim3(256), 0, st, make_seg_table(h), p->T, p->ldt, p->lse, h->b2, d_slot, d_col, n, h->cfg.self_norm ? 0 : 1, p->out.as<double>()));
  JLM_CUDA(cudaGetLastError());
  JLM_CUDA(cudaMemcpyAsync(p->out_host.p, p->out.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  JLM_CUDA(cudaStreamSynchronize(st));
  memcpy(out, p->out_host.p, sizeof(double) * (size_t)n);
  return 0;
}

extern "C" int32_t jlm_pool_get_state(jlm_pool* p, int64_t slot, int32_t count, double* h_out, double* c_out) {
  JLM_REQUIRE(p && count > 0 && slot >= 0 && slot + count <= p->used, "jlm_pool_get_state: slots out of range");
  jlm_handle* h = p->h;
  JLM_CUDA(cudaSetDevice(h->device));
  JLM_CUDA(cudaStreamSynchronize(h->stream));
  for (int which = 0; which < 2; ++which) {
    double* dst = which ? c_out : h_out;
    if (!dst) continue;
    const double* src = (which ? p->cx : p->hx) + slot * h->Hp;
    JLM_CUDA(cudaMemcpy2D(dst, sizeof(double) * h->H, src, sizeof(double) * h->Hp, sizeof(double) * h->H, count,
                          cudaMemcpyDeviceToHost));
  }
  return 0;
}
