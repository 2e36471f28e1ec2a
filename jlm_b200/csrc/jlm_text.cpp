// Decoder.decode for a batch of kana strings in ONE call (decoder/decoder.py:220-241,
// decoder/decoder_dynamic.py:177-194): lattice build -> plan -> host-to-device -> all frames on the device
// -> n-best back on the host.  The batch is cut into chunks that flow through those stages as a
// software pipeline: while the device decodes chunk c, the host builds the lattices and the plan of
// chunk c+1 and enqueues its kernels, so only the first chunk's host work is exposed.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/jlm_b200.h"

void jlm_set_error(const char* fmt, ...);

namespace {
struct Chunk {
  int32_t s0 = 0, n = 0;
  jlm_lattice* lat = nullptr;
  jlm_batch* batch = nullptr;
};
}  // namespace

struct jlm_text_job {
  jlm_handle* h = nullptr;
  int32_t n_sent = 0, top = 0;
  bool timers = false;
  std::vector<Chunk> chunks;
};

static void job_free(jlm_text_job* job) {
  if (!job) return;
  for (auto& ck : job->chunks) {
    if (ck.batch) jlm_batch_destroy(ck.batch);
    if (ck.lat) jlm_lattice_destroy(ck.lat);
  }
  delete job;
}

// First half of jlm_decode_texts: everything up to and including the enqueue of the n-best D2H copy.
// Returns as soon as the work is queued; the text buffers are not referenced after the call.
extern "C" int32_t jlm_decode_texts_submit(jlm_handle* h, const jlm_lexicon* lex, int32_t n_sent, const int64_t* text_ptr,
                                           const uint32_t* text, int32_t beam_width, int32_t top_n, int32_t mode,
                                           int32_t n_extra, const int32_t* extra_ids, int32_t backend, int32_t n_chunks,
                                           int32_t timers, jlm_text_job** out_job) {
  if (!h || !lex || !text_ptr || !out_job || n_sent <= 0) {
    jlm_set_error("jlm_decode_texts_submit: bad argument");
    return 1;
  }
  *out_job = nullptr;
  const int32_t top = beam_width == JLM_BEAM_UNLIMITED ? std::max(1, top_n) : std::max(1, std::min(top_n, beam_width));
  // Chunk boundaries.  Measured on B200 (1024 sentences, cfg 2): equal splits cost more device time
  // (smaller GEMMs, per-frame launch overheads) than the host work they hide (1.15 M chars/s with 1 chunk,
  // 1.11 M with 2, 0.74 M with 8), and a small head chunk that gets the device busy while the host prepares
  // the rest gains < 3 %.  The automatic choice is therefore one lock-step batch, cut only into
  // 4096-sentence pieces to bound memory; JLM_HEAD_CHUNK=n asks for a head chunk of n sentences.
  std::vector<int32_t> bounds{0};
  if (n_chunks <= 0) {
    const char* e = getenv("JLM_HEAD_CHUNK");
    const int32_t head = e ? atoi(e) : 0;
    if (head > 0 && head < n_sent) bounds.push_back(head);
    while (n_sent - bounds.back() > 4096) bounds.push_back(bounds.back() + 4096);
    bounds.push_back(n_sent);
  } else {
    n_chunks = std::min(n_chunks, n_sent);
    for (int c = 1; c <= n_chunks; ++c) bounds.push_back((int32_t)((int64_t)n_sent * c / n_chunks));
  }
  n_chunks = (int32_t)bounds.size() - 1;
  jlm_text_job* job = new jlm_text_job();
  job->h = h;
  job->n_sent = n_sent;
  job->top = top;
  job->timers = timers != 0;
  job->chunks.resize(n_chunks);
  std::vector<Chunk>& chunks = job->chunks;
  int32_t rc = 0;
  // stage 1..3 for every chunk: build, upload, enqueue (all asynchronous with respect to the device)
  for (int c = 0; c < n_chunks && !rc; ++c) {
    Chunk& ck = chunks[c];
    ck.s0 = bounds[c];
    ck.n = bounds[c + 1] - bounds[c];
    rc = jlm_lattice_build(lex, ck.n, text_ptr + ck.s0, text, mode, n_extra,
                           n_extra ? extra_ids + (int64_t)ck.s0 * n_extra : nullptr, &ck.lat);
    jlm_lattice_batch view;
    if (!rc) rc = jlm_lattice_view(ck.lat, &view, nullptr, nullptr);
    if (!rc) rc = jlm_batch_upload(h, &view, beam_width, top, mode, backend, &ck.batch);
    if (!rc && timers) rc = jlm_batch_enable_timers(ck.batch, 1);
    if (!rc) rc = jlm_batch_run(ck.batch);
    if (!rc) rc = jlm_batch_fetch_async(ck.batch);
  }
  if (rc) {
    job_free(job);
    return rc;
  }
  *out_job = job;
  return 0;
}

// Second half: waits for the job's own batches (not for work submitted after it), translates node
// indices into (lexicon entry, start frame), releases the job.  The job is consumed even on error.
extern "C" int32_t jlm_decode_texts_collect(jlm_text_job* job, jlm_text_nbest* out, jlm_batch_info* info) {
  if (!job || !out) {
    jlm_set_error("jlm_decode_texts_collect: bad argument");
    job_free(job);
    return 1;
  }
  if (!out->scores || !out->n_paths || !out->path_len || !out->path_entry || !out->path_start) {
    jlm_set_error("jlm_decode_texts_collect: null output arrays");
    job_free(job);
    return 1;
  }
  const int32_t top = job->top;
  if (out->top_n < top) {
    jlm_set_error("jlm_decode_texts_collect: output top_n %d < %d", out->top_n, top);
    job_free(job);
    return 1;
  }
  std::vector<Chunk>& chunks = job->chunks;
  const int32_t n_chunks = (int32_t)chunks.size();
  int32_t rc = 0;
  // stage 4: fetch in order; translate node indices into (lexicon entry, start frame)
  std::vector<double> sc;
  std::vector<int32_t> np_, ln, nd;
  if (info) memset(info, 0, sizeof(*info));
  for (int c = 0; c < n_chunks && !rc; ++c) {
    Chunk& ck = chunks[c];
    jlm_lattice_batch view;
    const int32_t* node_entry = nullptr;
    int64_t n_nodes = 0;
    rc = jlm_lattice_view(ck.lat, &view, &node_entry, &n_nodes);
    if (rc) break;
    int32_t max_len = 1;
    for (int32_t s = 0; s < ck.n; ++s) max_len = std::max(max_len, view.sent_len[s] + 1);
    if (out->max_len < max_len) {
      jlm_set_error("jlm_decode_texts: output max_len %d < %d", out->max_len, max_len);
      rc = 1;
      break;
    }
    sc.resize((size_t)ck.n * top);
    np_.resize(ck.n);
    ln.resize((size_t)ck.n * top);
    nd.resize((size_t)ck.n * top * max_len);
    jlm_nbest nb{top, max_len, sc.data(), np_.data(), ln.data(), nd.data()};
    rc = jlm_batch_fetch(ck.batch, &nb);
    if (rc) break;
    if (info) {   // totals over the chunks; n_steps is the longest chunk's (they run one after the other)
      jlm_batch_info bi;
      rc = jlm_batch_get_info(ck.batch, &bi);
      if (rc) break;
      info->n_slots += bi.n_slots;
      info->n_candidates += bi.n_candidates;
      info->n_nodes += bi.n_nodes;
      info->n_steps = std::max(info->n_steps, bi.n_steps);
      info->backend = bi.backend;
      info->kernel_launches += bi.kernel_launches;
      info->h2d_bytes += bi.h2d_bytes;
      info->d2h_bytes += bi.d2h_bytes;
      info->ms_lstm += bi.ms_lstm;
      info->ms_softmax += bi.ms_softmax;
      info->ms_beam += bi.ms_beam;
      info->ms_gate_gemm += bi.ms_gate_gemm;
      info->ms_proj_gemm += bi.ms_proj_gemm;
      info->n_gate_launches += bi.n_gate_launches;
      info->n_proj_launches += bi.n_proj_launches;
      info->beam_width = std::max(info->beam_width, bi.beam_width);
      info->n_guard_flagged += bi.n_guard_flagged;
      info->n_guard_pairs += bi.n_guard_pairs;
      info->n_guard_rerun += bi.n_guard_rerun;
      info->guard_eps = bi.guard_eps;
      info->guard_min_gap = (c == 0) ? bi.guard_min_gap : std::min(info->guard_min_gap, bi.guard_min_gap);
    }
    for (int32_t s = 0; s < ck.n; ++s) {
      const int64_t gs = ck.s0 + s;
      out->n_paths[gs] = np_[s];
      for (int32_t k = 0; k < out->top_n; ++k) {
        const size_t o = (size_t)gs * out->top_n + k;
        if (k >= top) {
          out->scores[o] = INFINITY;
          out->path_len[o] = 0;
          continue;
        }
        const size_t i = (size_t)s * top + k;
        out->scores[o] = sc[i];
        out->path_len[o] = ln[i];
        for (int32_t q = 0; q < ln[i] && q < max_len; ++q) {
          const int32_t node = nd[i * max_len + q];
          out->path_entry[o * out->max_len + q] = node_entry[node];
          out->path_start[o * out->max_len + q] = view.node_start[node];
        }
      }
    }
  }
  job_free(job);
  return rc;
}

extern "C" int32_t jlm_decode_texts_cancel(jlm_text_job* job) {
  job_free(job);
  return 0;
}

extern "C" int32_t jlm_decode_texts(jlm_handle* h, const jlm_lexicon* lex, int32_t n_sent, const int64_t* text_ptr,
                                    const uint32_t* text, int32_t beam_width, int32_t top_n, int32_t mode,
                                    int32_t n_extra, const int32_t* extra_ids, int32_t backend, int32_t n_chunks,
                                    jlm_text_nbest* out, jlm_batch_info* info) {
  if (!out) {
    jlm_set_error("jlm_decode_texts: bad argument");
    return 1;
  }
  jlm_text_job* job = nullptr;
  int32_t rc = jlm_decode_texts_submit(h, lex, n_sent, text_ptr, text, beam_width, top_n, mode, n_extra, extra_ids,
                                       backend, n_chunks, info != nullptr, &job);
  if (rc) return rc;
  return jlm_decode_texts_collect(job, out, info);
}
