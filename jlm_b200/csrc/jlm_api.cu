// C-ABI entry points: handle life cycle, weight re-layout, and the model-level API
// (LSTM_Model._lstm_cell / project / predict, decoder/model.py:106-198) on the exact back end.
#include <cstdarg>
#include <cmath>

#include <algorithm>

#include "jlm_common.cuh"

static thread_local char g_err[1024] = "";

void jlm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* jlm_last_error(void) { return g_err; }
extern "C" int32_t jlm_abi_version(void) { return JLM_ABI_VERSION; }

namespace {

template <class T>
int32_t upload(T** dst, const std::vector<T>& src) {
  JLM_CUDA(cudaMalloc(dst, src.size() * sizeof(T)));
  JLM_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

int32_t build_weights(jlm_handle* h, const jlm_weights* w) {
  const int H = h->H, E = h->E, V = h->V, Hp = h->Hp, Ep = h->Ep, Kg = h->Kg;
  for (int g = 0; g < 4; ++g)
    JLM_REQUIRE(w->HM[g] && w->IM[g] && w->b[g], "jlm_create: missing gate weights for gate %d", g);
  JLM_REQUIRE(w->b2 && w->LM_in, "jlm_create: b2 / LM_in missing");
  {  // gate weight [4H, Kg], K-major: decoder/model.py:128-131 as one contraction
    std::vector<float> Wg((size_t)4 * H * Kg, 0.f), bg((size_t)4 * H);
    for (int g = 0; g < 4; ++g)
      for (int j = 0; j < H; ++j) {
        float* row = &Wg[((size_t)g * H + j) * Kg];
        for (int k = 0; k < H; ++k) row[k] = w->HM[g][(size_t)k * H + j];
        for (int e = 0; e < E; ++e) row[Hp + e] = w->IM[g][(size_t)e * H + j];
        bg[(size_t)g * H + j] = w->b[g][j];
      }
    JLM_TRY(upload(&h->Wg, Wg));
    JLM_TRY(upload(&h->bg, bg));
  }
  {
    std::vector<float> b2(w->b2, w->b2 + V);
    JLM_TRY(upload(&h->b2, b2));
    std::vector<float> LM((size_t)V * Ep, 0.f);
    for (int v = 0; v < V; ++v) memcpy(&LM[(size_t)v * Ep], w->LM_in + (size_t)v * E, sizeof(float) * E);
    JLM_TRY(upload(&h->LM_in, LM));
  }
  // output blocks [V_i, kpad_i]
  for (int i = 0; i < h->n_seg; ++i) {
    SegDev& s = h->seg[i];
    const int Vi = s.end - s.start;
    std::vector<float> W((size_t)Vi * s.kpad, 0.f);
    if (h->untied) {
      JLM_REQUIRE(w->UM, "jlm_create: UM missing for the untied projection");
      for (int k = 0; k < H; ++k) {
        const float* src = w->UM + (size_t)k * V;
        for (int v = 0; v < V; ++v) W[(size_t)v * s.kpad + k] = src[v];
      }
    } else {
      JLM_REQUIRE(w->seg_LM[i], "jlm_create: seg_LM[%d] missing", i);
      for (int v = 0; v < Vi; ++v) memcpy(&W[(size_t)v * s.kpad], w->seg_LM[i] + (size_t)v * s.width, sizeof(float) * s.width);
    }
    JLM_TRY(upload(&h->Wseg_store[i], W));
    s.W = h->Wseg_store[i];
  }
  if (!h->untied) {
    // stage-1 weight P1 [Kt, Hp] in float64.  D-softmax*: (h.PM).VT_i^T == h.(PM.VT_i^T), the product
    // is formed here once in float64 (decoder/model.py:171-172 evaluates it per call in float64).
    JLM_REQUIRE(w->PM, "jlm_create: PM missing for a tied projection");
    std::vector<double> P1((size_t)h->Kt * Hp, 0.0);
    const int mode = h->cfg.proj_mode;
    int pm_cols = 0;
    if (mode == JLM_PROJ_DSOFTMAX) {
      for (int i = 0; i < h->n_seg; ++i) pm_cols += h->seg[i].width;
    } else {
      pm_cols = h->seg[0].width;
    }
    int col = 0;
    for (int i = 0; i < h->n_seg; ++i) {
      const SegDev& s = h->seg[i];
      for (int e = 0; e < s.width; ++e) {
        double* dst = &P1[(size_t)(s.koff + e) * Hp];
        if (mode == JLM_PROJ_DSOFTMAX_STAR && i > 0) {
          JLM_REQUIRE(w->seg_VT[i], "jlm_create: seg_VT[%d] missing", i);
          const float* vt = w->seg_VT[i] + (size_t)e * pm_cols;
          for (int k = 0; k < H; ++k) {
            const float* pm = w->PM + (size_t)k * pm_cols;
            double a = 0.0;
            for (int q = 0; q < pm_cols; ++q) a += (double)pm[q] * (double)vt[q];
            dst[k] = a;
          }
        } else {
          const int c = (mode == JLM_PROJ_DSOFTMAX) ? col + e : e;
          for (int k = 0; k < H; ++k) dst[k] = (double)w->PM[(size_t)k * pm_cols + c];
        }
      }
      col += s.width;
    }
    JLM_TRY(upload(&h->P1, P1));
  }
  return 0;
}

}  // namespace

extern "C" int32_t jlm_create(const jlm_config* cfg, const jlm_weights* w, int32_t device, jlm_handle** out) {
  JLM_REQUIRE(cfg && w && out, "jlm_create: null argument");
  *out = nullptr;
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  JLM_REQUIRE(e == cudaSuccess && n_dev > 0,
              "jlm_create: no usable CUDA device (%s); libjlm_b200 has no CPU fallback",
              e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  JLM_REQUIRE(device >= 0 && device < n_dev, "jlm_create: device %d out of range (%d devices)", device, n_dev);
  JLM_REQUIRE(cfg->vocab_size > 0 && cfg->hidden_size > 0 && cfg->input_embed > 0, "jlm_create: bad sizes");
  JLM_REQUIRE(cfg->n_seg >= 1 && cfg->n_seg <= JLM_MAX_SEGMENTS, "jlm_create: n_seg %d out of range", cfg->n_seg);
  JLM_REQUIRE(cfg->proj_mode >= 0 && cfg->proj_mode <= 3, "jlm_create: bad proj_mode");
  JLM_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  JLM_CUDA(cudaGetDeviceProperties(&prop, device));
  JLM_REQUIRE(prop.major == 10, "jlm_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only",
              device, prop.major, prop.minor);
  jlm_handle* h = new jlm_handle();
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->cfg = *cfg;
  h->V = cfg->vocab_size;
  h->H = cfg->hidden_size;
  h->E = cfg->input_embed;
  h->Hp = (int)round_up64(h->H, JLM_KALIGN);
  h->Ep = (int)round_up64(h->E, JLM_KALIGN);
  h->Kg = h->Hp + h->Ep;
  h->untied = cfg->proj_mode == JLM_PROJ_UNTIED;
  h->n_seg = cfg->n_seg;
  int koff = 0, expect = 0;
  for (int i = 0; i < h->n_seg; ++i) {
    SegDev& s = h->seg[i];
    s.start = cfg->seg_start[i];
    s.end = cfg->seg_end[i];
    s.width = h->untied ? h->H : cfg->seg_width[i];
    s.kpad = (int)round_up64(s.width, JLM_KALIGN);
    s.koff = koff;
    koff += s.kpad;
    if (s.start != expect || s.end <= s.start) {
      jlm_set_error("jlm_create: segments must tile [0,V) in order (segment %d is [%d,%d))", i, s.start, s.end);
      delete h;
      return 1;
    }
    expect = s.end;
  }
  if (expect != h->V || (h->untied && h->n_seg != 1)) {
    jlm_set_error("jlm_create: segments cover [0,%d) but V=%d", expect, h->V);
    delete h;
    return 1;
  }
  h->Kt = koff;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    jlm_set_error("jlm_create: cudaStreamCreate failed");
    delete h;
    return 1;
  }
  h->own_stream = true;
  if (const char* e = getenv("JLM_Q8")) h->q8_policy = atoi(e) ? 1 : 0;
  if (const char* e = getenv("JLM_GUARD_EPS")) h->guard_eps = std::max(0.0, atof(e));
  if (const char* e = getenv("JLM_GUARD_VERIFY")) h->guard_verify = atoi(e) != 0;
  if (const char* e = getenv("JLM_GUARD_ALL")) h->guard_all = atoi(e) != 0;
  for (auto& ev : h->ev) cudaEventCreate(&ev);
  if (build_weights(h, w)) {
    jlm_destroy(h);
    return 1;
  }
  *out = h;
  return 0;
}

extern "C" int32_t jlm_destroy(jlm_handle* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (auto& cs : h->copy_stream)
    if (cs) {
      cudaStreamSynchronize(cs);
      cudaStreamDestroy(cs);
    }
  tc_free_weights(h);
  beam_free_plan_scratch(h);
  beam_free_guard(h);
  cudaFree(h->Wg);
  cudaFree(h->bg);
  cudaFree(h->b2);
  cudaFree(h->LM_in);
  cudaFree(h->P1);
  for (auto& p : h->Wseg_store) cudaFree(p);
  for (auto& p : h->Wq_store) cudaFree(p);
  for (auto& p : h->cb_store) cudaFree(p);
  for (auto& b : h->scratch) b.release();
  for (auto& b : h->batch_cache) b.release();
  for (auto& b : h->out_pool) b.release();
  for (auto& b : h->pinned) b.release();
  for (auto& b : h->stage) b.release();
  for (auto& ev : h->stage_ev)
    if (ev) cudaEventDestroy(ev);
  for (auto& ev : h->ev)
    if (ev) cudaEventDestroy(ev);
  if (h->side_stream) cudaStreamDestroy(h->side_stream);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

extern "C" int32_t jlm_set_quantized_block(jlm_handle* h, int32_t segment, const uint8_t* code, const float* codebook,
                                           int32_t n_codes) {
  JLM_REQUIRE(h && code && codebook, "jlm_set_quantized_block: null argument");
  JLM_REQUIRE(segment >= 0 && segment < h->n_seg, "jlm_set_quantized_block: segment %d out of range", segment);
  JLM_REQUIRE(n_codes >= 1 && n_codes <= 256, "jlm_set_quantized_block: n_codes %d not in [1,256]", n_codes);
  JLM_CUDA(cudaSetDevice(h->device));
  SegDev& s = h->seg[segment];
  const int64_t Vi = s.end - s.start;
  // the codes must decode to exactly the float32 block already on the device
  std::vector<float> W((size_t)Vi * s.kpad);
  JLM_CUDA(cudaMemcpy(W.data(), s.W, W.size() * sizeof(float), cudaMemcpyDeviceToHost));
  std::vector<uint8_t> q((size_t)Vi * s.kpad, 0);
  for (int64_t v = 0; v < Vi; ++v)
    for (int k = 0; k < s.width; ++k) {
      const uint8_t c = code[v * s.width + k];
      JLM_REQUIRE(c < n_codes, "jlm_set_quantized_block: code %d >= n_codes %d", (int)c, n_codes);
      JLM_REQUIRE(memcmp(&codebook[c], &W[(size_t)v * s.kpad + k], sizeof(float)) == 0,
                  "jlm_set_quantized_block: codebook[code] differs from the float32 weight at row %lld col %d",
                  (long long)v, k);
      q[(size_t)v * s.kpad + k] = c;
    }
  std::vector<float> cb(256, 0.f);
  memcpy(cb.data(), codebook, sizeof(float) * n_codes);
  if (h->Wq_store[segment]) cudaFree(h->Wq_store[segment]);
  if (h->cb_store[segment]) cudaFree(h->cb_store[segment]);
  h->Wq_store[segment] = nullptr;
  h->cb_store[segment] = nullptr;
  JLM_TRY(upload(&h->Wq_store[segment], q));
  JLM_TRY(upload(&h->cb_store[segment], cb));
  s.Wq = h->Wq_store[segment];
  s.cb = h->cb_store[segment];
  return 0;
}

extern "C" int32_t jlm_set_stream(jlm_handle* h, void* cuda_stream) {
  JLM_REQUIRE(h, "jlm_set_stream: null handle");
  if (h->own_stream && h->stream) {
    cudaStreamSynchronize(h->stream);
    cudaStreamDestroy(h->stream);
  }
  h->stream = static_cast<cudaStream_t>(cuda_stream);
  h->own_stream = false;
  return 0;
}

extern "C" int32_t jlm_set_guard(jlm_handle* h, double eps) {
  JLM_REQUIRE(h, "jlm_set_guard: null handle");
  JLM_REQUIRE(eps == eps, "jlm_set_guard: eps is NaN");
  if (eps < 0.0) {
    const char* e = getenv("JLM_GUARD_EPS");
    eps = e ? std::max(0.0, atof(e)) : JLM_GUARD_EPS_DEFAULT;
  }
  h->guard_eps = eps;
  return 0;
}

extern "C" int32_t jlm_set_guard_scope(jlm_handle* h, int32_t all_decisions) {
  JLM_REQUIRE(h, "jlm_set_guard_scope: null handle");
  h->guard_all = all_decisions != 0;
  return 0;
}

extern "C" int32_t jlm_set_guard_verify(jlm_handle* h, int32_t on) {
  JLM_REQUIRE(h, "jlm_set_guard_verify: null handle");
  h->guard_verify = on != 0;
  return 0;
}

extern "C" int32_t jlm_synchronize(jlm_handle* h) {
  JLM_REQUIRE(h, "jlm_synchronize: null handle");
  JLM_CUDA(cudaSetDevice(h->device));
  JLM_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// model-level API (exact back end)
// ------------------------------------------------------------------------------------------------
namespace {

// scratch slots
enum { S_A = 0, S_GATES, S_STATE_IN, S_STATE_OUT, S_T, S_Y, S_PART, S_IDX };

int32_t h2d_padded(jlm_handle* h, double* dst, const double* src, int B, int W, int Wp) {
  JLM_CUDA(cudaMemsetAsync(dst, 0, sizeof(double) * (size_t)B * Wp, h->stream));
  JLM_CUDA(cudaMemcpy2DAsync(dst, sizeof(double) * Wp, src, sizeof(double) * W, sizeof(double) * W, B,
                             cudaMemcpyHostToDevice, h->stream));
  return 0;
}

int32_t d2h_padded(jlm_handle* h, double* dst, const double* src, int B, int W, int Wp) {
  JLM_CUDA(cudaMemcpy2DAsync(dst, sizeof(double) * W, src, sizeof(double) * Wp, sizeof(double) * W, B,
                             cudaMemcpyDeviceToHost, h->stream));
  return 0;
}

// device-side _lstm_cell: state_in/out are [2][B,Hp] (h then c)
int32_t dev_lstm_step(jlm_handle* h, const int32_t* d_index, const double* d_hin, const double* d_cin, int B,
                      double* d_hout, double* d_cout) {
  JLM_TRY(h->scratch[S_A].reserve(sizeof(double) * (size_t)B * h->Kg));
  JLM_TRY(h->scratch[S_GATES].reserve(sizeof(double) * (size_t)B * 4 * h->H));
  double* A = h->scratch[S_A].as<double>();
  double* G = h->scratch[S_GATES].as<double>();
  JLM_TRY(exact_gather_gate_input(h->stream, h, d_hin, nullptr, d_index, B, A));
  JLM_TRY(exact_gemm_f32w(h->stream, A, h->Kg, h->Wg, h->Kg, h->bg, G, 4 * h->H, B, 4 * h->H, h->Kg, nullptr, 0, 0));
  JLM_TRY(exact_lstm_pointwise(h->stream, h, G, d_cin, nullptr, B, d_hout, d_cout));
  return 0;
}

// device-side project: d_hidden [B,Hp] -> d_y [B,N] (+ optional lse [B])
int32_t dev_project(jlm_handle* h, const double* d_hidden, int B, const int32_t* d_cols, const int32_t* d_bias,
                    int n_cols, double* d_y, double* d_lse) {
  const double* T = d_hidden;
  int ldt = h->Hp;
  if (!h->untied) {
    JLM_TRY(h->scratch[S_T].reserve(sizeof(double) * (size_t)B * h->Kt));
    double* Tb = h->scratch[S_T].as<double>();
    JLM_TRY(exact_gemm_f64w(h->stream, d_hidden, h->Hp, h->P1, h->Hp, Tb, h->Kt, B, h->Kt, h->Hp));
    T = Tb;
    ldt = h->Kt;
  }
  if (!d_cols) {
    const int V = h->V;
    int tiles = 0;
    for (int i = 0; i < h->n_seg; ++i) tiles += exact_tiles_n(h->seg[i].end - h->seg[i].start);
    double2* part = nullptr;
    if (d_lse) {
      JLM_TRY(h->scratch[S_PART].reserve(sizeof(double2) * (size_t)B * tiles));
      part = h->scratch[S_PART].as<double2>();
    }
    int tile0 = 0;
    for (int i = 0; i < h->n_seg; ++i) {
      const SegDev& s = h->seg[i];
      const int Vi = s.end - s.start;
      if (exact_use_q8(h, s, B))
        JLM_TRY(exact_gemm_q8w(h->stream, T + s.koff, ldt, s.Wq, s.kpad, s.cb, h->b2 + s.start,
                               d_y ? d_y + s.start : nullptr, V, B, Vi, s.kpad, part, tiles, tile0));
      else
        JLM_TRY(exact_gemm_f32w(h->stream, T + s.koff, ldt, s.W, s.kpad, h->b2 + s.start, d_y ? d_y + s.start : nullptr, V,
                                B, Vi, s.kpad, part, tiles, tile0));
      tile0 += exact_tiles_n(Vi);
    }
    if (d_lse) JLM_TRY(exact_lse_merge(h->stream, part, tiles, tiles, B, d_lse, 0));
  } else {
    SubsetJob job{0, B, 0, n_cols, 0};
    JLM_TRY(h->scratch[S_PART].reserve(sizeof(SubsetJob)));
    SubsetJob* d_job = h->scratch[S_PART].as<SubsetJob>();
    JLM_CUDA(cudaMemcpyAsync(d_job, &job, sizeof(job), cudaMemcpyHostToDevice, h->stream));
    JLM_TRY(subset_logits<double>(h->stream, h, T, ldt, d_job, 1, n_cols, d_cols, d_bias, d_y));
    if (d_lse) JLM_TRY(exact_rows_lse(h->stream, d_y, n_cols, B, n_cols, d_lse));
  }
  return 0;
}

int32_t upload_cols(jlm_handle* h, const int32_t* cols, const int32_t* bias_idx, int n_cols, const int32_t** d_cols,
                    const int32_t** d_bias, size_t extra_ints) {
  JLM_REQUIRE(n_cols > 0, "vocab subset must not be empty");
  for (int j = 0; j < n_cols; ++j) {
    JLM_REQUIRE(cols[j] >= 0 && cols[j] < h->V, "vocab id %d out of range", cols[j]);
    if (bias_idx) JLM_REQUIRE(bias_idx[j] >= 0 && bias_idx[j] < h->V, "vocab id %d out of range", bias_idx[j]);
  }
  JLM_TRY(h->scratch[S_IDX].reserve(sizeof(int32_t) * ((size_t)2 * n_cols + extra_ints)));
  int32_t* d = h->scratch[S_IDX].as<int32_t>();
  JLM_CUDA(cudaMemcpyAsync(d, cols, sizeof(int32_t) * n_cols, cudaMemcpyHostToDevice, h->stream));
  *d_cols = d;
  *d_bias = nullptr;
  if (bias_idx) {
    JLM_CUDA(cudaMemcpyAsync(d + n_cols, bias_idx, sizeof(int32_t) * n_cols, cudaMemcpyHostToDevice, h->stream));
    *d_bias = d + n_cols;
  }
  return 0;
}

}  // namespace

extern "C" int32_t jlm_lstm_step(jlm_handle* h, const int32_t* index, const double* h_in, const double* c_in,
                                 int32_t B, double* h_out, double* c_out) {
  JLM_REQUIRE(h && index && h_in && c_in && h_out && c_out && B > 0, "jlm_lstm_step: bad argument");
  JLM_CUDA(cudaSetDevice(h->device));
  for (int i = 0; i < B; ++i) JLM_REQUIRE(index[i] >= 0 && index[i] < h->V, "jlm_lstm_step: index %d out of range", index[i]);
  const size_t sb = (size_t)B * h->Hp;
  JLM_TRY(h->scratch[S_STATE_IN].reserve(sizeof(double) * 2 * sb));
  JLM_TRY(h->scratch[S_STATE_OUT].reserve(sizeof(double) * 2 * sb));
  JLM_TRY(h->scratch[S_IDX].reserve(sizeof(int32_t) * (size_t)B));
  double* sin = h->scratch[S_STATE_IN].as<double>();
  double* sout = h->scratch[S_STATE_OUT].as<double>();
  int32_t* d_idx = h->scratch[S_IDX].as<int32_t>();
  JLM_CUDA(cudaMemcpyAsync(d_idx, index, sizeof(int32_t) * B, cudaMemcpyHostToDevice, h->stream));
  JLM_TRY(h2d_padded(h, sin, h_in, B, h->H, h->Hp));
  JLM_TRY(h2d_padded(h, sin + sb, c_in, B, h->H, h->Hp));
  JLM_TRY(dev_lstm_step(h, d_idx, sin, sin + sb, B, sout, sout + sb));
  JLM_TRY(d2h_padded(h, h_out, sout, B, h->H, h->Hp));
  JLM_TRY(d2h_padded(h, c_out, sout + sb, B, h->H, h->Hp));
  JLM_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int32_t jlm_project(jlm_handle* h, const double* hidden, int32_t B, const int32_t* cols,
                               const int32_t* bias_idx, int32_t n_cols, double* y_out) {
  JLM_REQUIRE(h && hidden && y_out && B > 0, "jlm_project: bad argument");
  JLM_CUDA(cudaSetDevice(h->device));
  const int N = cols ? n_cols : h->V;
  const size_t sb = (size_t)B * h->Hp;
  JLM_TRY(h->scratch[S_STATE_IN].reserve(sizeof(double) * sb));
  JLM_TRY(h->scratch[S_Y].reserve(sizeof(double) * (size_t)B * N));
  double* d_h = h->scratch[S_STATE_IN].as<double>();
  double* d_y = h->scratch[S_Y].as<double>();
  const int32_t *d_cols = nullptr, *d_bias = nullptr;
  if (cols) JLM_TRY(upload_cols(h, cols, bias_idx, n_cols, &d_cols, &d_bias, 0));
  JLM_TRY(h2d_padded(h, d_h, hidden, B, h->H, h->Hp));
  JLM_TRY(dev_project(h, d_h, B, d_cols, d_bias, n_cols, d_y, nullptr));
  JLM_CUDA(cudaMemcpyAsync(y_out, d_y, sizeof(double) * (size_t)B * N, cudaMemcpyDeviceToHost, h->stream));
  JLM_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int32_t jlm_predict(jlm_handle* h, const int32_t* index, const double* h_in, const double* c_in, int32_t B,
                               const int32_t* cols, const int32_t* bias_idx, int32_t n_cols, double* pred_out,
                               double* y_out, double* h_out, double* c_out, float* ms_lstm, float* ms_softmax) {
  JLM_REQUIRE(h && index && h_in && c_in && pred_out && y_out && h_out && c_out && B > 0, "jlm_predict: bad argument");
  JLM_CUDA(cudaSetDevice(h->device));
  for (int i = 0; i < B; ++i) JLM_REQUIRE(index[i] >= 0 && index[i] < h->V, "jlm_predict: index %d out of range", index[i]);
  const int N = cols ? n_cols : h->V;
  const size_t sb = (size_t)B * h->Hp;
  JLM_TRY(h->scratch[S_STATE_IN].reserve(sizeof(double) * 2 * sb));
  JLM_TRY(h->scratch[S_STATE_OUT].reserve(sizeof(double) * 2 * sb));
  JLM_TRY(h->scratch[S_Y].reserve(sizeof(double) * ((size_t)2 * B * N + B)));
  double* sin = h->scratch[S_STATE_IN].as<double>();
  double* sout = h->scratch[S_STATE_OUT].as<double>();
  double* d_y = h->scratch[S_Y].as<double>();
  double* d_pred = d_y + (size_t)B * N;
  double* d_lse = d_pred + (size_t)B * N;
  const int32_t *d_cols = nullptr, *d_bias = nullptr;
  int32_t* d_idx = nullptr;
  if (cols) {
    JLM_TRY(upload_cols(h, cols, bias_idx, n_cols, &d_cols, &d_bias, B));
    d_idx = h->scratch[S_IDX].as<int32_t>() + 2 * (size_t)n_cols;
  } else {
    JLM_TRY(h->scratch[S_IDX].reserve(sizeof(int32_t) * (size_t)B));
    d_idx = h->scratch[S_IDX].as<int32_t>();
  }
  JLM_CUDA(cudaMemcpyAsync(d_idx, index, sizeof(int32_t) * B, cudaMemcpyHostToDevice, h->stream));
  JLM_TRY(h2d_padded(h, sin, h_in, B, h->H, h->Hp));
  JLM_TRY(h2d_padded(h, sin + sb, c_in, B, h->H, h->Hp));
  JLM_CUDA(cudaEventRecord(h->ev[0], h->stream));
  JLM_TRY(dev_lstm_step(h, d_idx, sin, sin + sb, B, sout, sout + sb));
  JLM_CUDA(cudaEventRecord(h->ev[1], h->stream));
  const bool sn = h->cfg.self_norm != 0;
  JLM_TRY(dev_project(h, sout, B, d_cols, d_bias, n_cols, d_y, sn ? nullptr : d_lse));
  JLM_TRY(exact_softmax_rows(h->stream, d_y, N, B, N, sn ? nullptr : d_lse, d_pred));
  JLM_CUDA(cudaEventRecord(h->ev[2], h->stream));
  JLM_CUDA(cudaMemcpyAsync(y_out, d_y, sizeof(double) * (size_t)B * N, cudaMemcpyDeviceToHost, h->stream));
  JLM_CUDA(cudaMemcpyAsync(pred_out, d_pred, sizeof(double) * (size_t)B * N, cudaMemcpyDeviceToHost, h->stream));
  JLM_TRY(d2h_padded(h, h_out, sout, B, h->H, h->Hp));
  JLM_TRY(d2h_padded(h, c_out, sout + sb, B, h->H, h->Hp));
  JLM_CUDA(cudaStreamSynchronize(h->stream));
  float a = 0.f, b = 0.f;
  cudaEventElapsedTime(&a, h->ev[0], h->ev[1]);
  cudaEventElapsedTime(&b, h->ev[1], h->ev[2]);
  if (ms_lstm) *ms_lstm = a;
  if (ms_softmax) *ms_softmax = b;
  return 0;
}
