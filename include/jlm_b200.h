/*
 * jlm_b200.h - C ABI of libjlm_b200.so: the B200 (sm_100a) implementation of JLM's numpy
 * inference hot path (decoder/model.py LSTM step + projection + softmax, decoder/decoder.py and
 * decoder/decoder_dynamic.py beam search).
 *
 * The reference has no FFI of its own: its boundary is the Python class surface
 * LSTM_Model.predict()/project() and Decoder.decode().  Each entry point
 * below names the reference interface it replaces; the Python mirror in jlm_b200/model.py,
 * decoder.py and decoder_dynamic.py binds them with ctypes (see INTEGRATION.md for the stub a
 * maintainer of the reference would add).
 *
 * Conventions: every function returns 0 on success, non-zero on failure with a message available
 * from jlm_last_error().  Handles are opaque, caller-owned, bound to one CUDA device and one
 * stream, and NOT thread-safe (the reference objects are not re-entrant either).  Host buffers
 * are caller-owned, dense row-major, little-endian.  Device memory is owned by the handle.
 * No call ever falls back to a CPU implementation: without a CUDA device jlm_create fails.
 */
#ifndef JLM_B200_H
#define JLM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JLM_ABI_VERSION 1
#define JLM_MAX_SEGMENTS 8
#define JLM_MAX_BEAM 128
/* beam_width value for the reference's beam_width=None (decoder.py:227-229 not executed): every candidate
 * of a frame is kept, in enumeration order (node order, then parent rank); nothing is sorted.  The number of
 * paths is exponential in the sentence length; a batch that would keep more than JLM_MAX_UNPRUNED_PATHS
 * paths is refused with an error. */
#define JLM_BEAM_UNLIMITED 0
#define JLM_MAX_UNPRUNED_PATHS (4ll << 20)

/* projection modes, decoder/model.py:141-193 */
enum {
  JLM_PROJ_UNTIED = 0,        /* share_embedding=False: h.UM + b2                 (model.py:187-191) */
  JLM_PROJ_TIED = 1,          /* (h.PM).LM^T + b2                                 (model.py:182-186) */
  JLM_PROJ_DSOFTMAX = 2,      /* per segment (h.PM)[:,cols_i].LM_i^T              (model.py:144-160) */
  JLM_PROJ_DSOFTMAX_STAR = 3  /* per segment ((h.PM).VT_i^T).LM_i^T               (model.py:161-181) */
};

/* arithmetic back ends */
enum {
  JLM_BACKEND_AUTO = 0,   /* exact for few rows in flight, tensor cores for lock-step batches */
  JLM_BACKEND_EXACT = 1,  /* CUDA-core kernels, float64 accumulation over float32 weights (mirrors the
                             reference's float64 state, model.py:44-45,128-139) */
  JLM_BACKEND_TC = 2      /* tcgen05 tensor-core GEMMs on 2-term fp16 splits of fp32 operands
                             (3 MMAs per product, fp32 accumulate in TMEM), fp32 state */
};

/* decode modes */
enum {
  JLM_DECODE_FULL = 0,          /* softmax over the whole vocabulary  (decoder.py:220-241)          */
  JLM_DECODE_STATIC_VOCAB = 1,  /* vocab_select=True: one sorted list per sentence (decoder.py:137-151) */
  JLM_DECODE_DYNAMIC = 2        /* DynamicDecoder: per-frame cumulative vocab (decoder_dynamic.py)   */
};

typedef struct jlm_handle jlm_handle;
typedef struct jlm_batch jlm_batch;
typedef struct jlm_lexicon jlm_lexicon;
typedef struct jlm_lattice jlm_lattice;
struct jlm_batch_info_s;

/* config.json keys the hot path reads (model.py:41-56,117) after normalisation by the host. */
typedef struct {
  int32_t vocab_size;    /* V */
  int32_t hidden_size;   /* H */
  int32_t input_embed;   /* width of the LSTM input embedding rows (model.py:43,49,56-71) */
  int32_t proj_mode;     /* JLM_PROJ_* */
  int32_t self_norm;     /* pred = exp(y) instead of softmax(y) (model.py:117-118) */
  int32_t n_seg;         /* 1 for untied / tied */
  int32_t seg_width[JLM_MAX_SEGMENTS];  /* e_i, config['embedding_seg'][i][0] */
  int32_t seg_start[JLM_MAX_SEGMENTS];  /* first word id of the segment */
  int32_t seg_end[JLM_MAX_SEGMENTS];    /* one past the last word id (None -> V) */
} jlm_config;

/* float32 weights in the layout of lstm_weights.pkl (train/weights.py:30-58; shapes from
 * train/model.py:137-150,188-193,54-62).  Gate order i,f,o,g. */
typedef struct {
  const float* HM[4];   /* [H,H]                                                         */
  const float* IM[4];   /* [input_embed,H]                                               */
  const float* b[4];    /* [H]                                                           */
  const float* b2;      /* [V]                                                           */
  const float* LM_in;   /* [V,input_embed] LSTM input table as materialised by model.py:47-71 */
  const float* PM;      /* tied modes: [H, sum(seg_width)] (dsoftmax) or [H, seg_width[0]] ; untied: NULL */
  const float* UM;      /* untied: [H,V]; else NULL                                      */
  const float* seg_LM[JLM_MAX_SEGMENTS]; /* output blocks [seg_end-seg_start, seg_width] (tied: LM) */
  const float* seg_VT[JLM_MAX_SEGMENTS]; /* dsoftmax*: VT_i [seg_width[i], seg_width[0]] for i>=1 */
} jlm_weights;

const char* jlm_last_error(void);
int32_t jlm_abi_version(void);

/* LSTM_Model.__init__ / _load_model (model.py:37-104): uploads and re-lays-out the weights on
 * `device`.  Fails (no fallback) when no CUDA device is usable. */
int32_t jlm_create(const jlm_config* cfg, const jlm_weights* w, int32_t device, jlm_handle** out);
int32_t jlm_destroy(jlm_handle* h);
/* Optional 8-bit form of one output block (train/comp.py:20-48,69-80: k-means code + codebook, the
 * reference's `comp_N/lstm_weights_comp_dump.pkl`): code is [seg_end-seg_start, seg_width] row-major,
 * codebook has n_codes (<= 256) float32 centroids.  codebook[code] must reproduce the float32 block
 * given to jlm_create bit for bit (checked), so results do not change; the few-rows-in-flight path then
 * streams 1 byte per weight instead of 4 (decoder/model.py:74-76 only ever sees the decoded floats). */
int32_t jlm_set_quantized_block(jlm_handle* h, int32_t segment, const uint8_t* code, const float* codebook,
                                int32_t n_codes);
/* Run all work of this handle on an existing CUDA stream (cudaStream_t passed as void*), e.g.
 * torch.cuda.current_stream().cuda_stream so torch.cuda.Event timing sees the kernels. */
int32_t jlm_set_stream(jlm_handle* h, void* cuda_stream);
/* Near-tie guard of the tensor-core back end.  Its path scores carry an error of a few 1e-6 per transition
 * (fp32 state, fp32-accumulated logits), so where the reference's sort (decoder.py:227-229) separates two candidates
 * by less than that, the tensor-core ranking is not certain to agree.  With eps > 0 every sentence in which some rank
 * decision of some frame - two adjacent kept paths, or the last kept path against the best rejected candidate -
 * rests on a gap < eps is checked inside jlm_batch_fetch / jlm_decode_*: the two paths of the decision are re-scored
 * in float64 (jlm_pool kernels); if that contradicts the tensor-core ranking - or the decode mode normalises over
 * per-frame word lists - the sentence is re-decoded on the float64 back end and that result is returned.
 * eps = 0 switches the guard off; eps < 0 restores the default (JLM_GUARD_EPS_DEFAULT, or the JLM_GUARD_EPS
 * environment variable read by jlm_create). */
#define JLM_GUARD_EPS_DEFAULT 1e-4
int32_t jlm_set_guard(jlm_handle* h, double eps);
/* Which rank decisions are guarded.  0 (default): those that can change what decode() returns - the kept/rejected
 * boundary of every frame and the order of the last frame's paths.  1: every adjacent pair of every frame, so the
 * per-frame rank order reported by jlm_batch_get_beams is certified too (JLM_GUARD_ALL). */
int32_t jlm_set_guard_scope(jlm_handle* h, int32_t all_decisions);
/* on = 0: skip the pair re-scoring and re-decode every flagged sentence in float64 (default on; JLM_GUARD_VERIFY) */
int32_t jlm_set_guard_verify(jlm_handle* h, int32_t on);
int32_t jlm_synchronize(jlm_handle* h);

/* LSTM_Model._lstm_cell (model.py:125-139) for B rows; state in/out float64 [B,H]. */
int32_t jlm_lstm_step(jlm_handle* h, const int32_t* index, const double* h_in, const double* c_in,
                      int32_t B, double* h_out, double* c_out);

/* LSTM_Model.project (model.py:141-193).  cols==NULL: y_out is [B,V].  Otherwise y_out is
 * [B,n_cols] with y[:,j] = logit_without_bias(cols[j]) + b2[bias_idx[j]]; passing two lists lets
 * the host reproduce the reference's column order for segmented models. */
int32_t jlm_project(jlm_handle* h, const double* hidden, int32_t B, const int32_t* cols,
                    const int32_t* bias_idx, int32_t n_cols, double* y_out);

/* LSTM_Model.predict / predict_with_context (model.py:106-123,195-198): step + project + softmax
 * (or exp when self_norm).  pred_out / y_out are [B,N], N = V or n_cols.  ms_lstm / ms_softmax
 * receive CUDA-event milliseconds for the two buckets the reference times (model.py:111-121). */
int32_t jlm_predict(jlm_handle* h, const int32_t* index, const double* h_in, const double* c_in,
                    int32_t B, const int32_t* cols, const int32_t* bias_idx, int32_t n_cols,
                    double* pred_out, double* y_out, double* h_out, double* c_out,
                    float* ms_lstm, float* ms_softmax);

/* A batch of kana lattices (Decoder._build_lattice output, decoder.py:79-135) in CSR form.
 * Sentence s has frames 0..sent_len[s]; frame_ptr holds, per sentence, sent_len[s]+2 offsets
 * (starting at frame_ptr_off[s]) into node_start/node_word: the nodes ENDING at each frame, in
 * the reference's order.  Frame 0 of every sentence holds exactly the <eos> node (start -1). */
typedef struct {
  int32_t n_sent;
  const int32_t* sent_len;       /* [n_sent] kana per sentence (T) */
  const int64_t* frame_ptr_off;  /* [n_sent] */
  const int64_t* frame_ptr;      /* [sum(T+2)] absolute node indices */
  const int32_t* node_start;     /* [n_nodes] start frame, -1 for <eos> */
  const int32_t* node_word;      /* [n_nodes] word id */
  /* vocabulary lists, only for JLM_DECODE_STATIC_VOCAB / JLM_DECODE_DYNAMIC:
   * STATIC : vocab_ids[vocab_ptr[s]..vocab_ptr[s+1]) is lattice_vocab (decoder.py:142-151).
   * DYNAMIC: the same range lists the sentence's words ordered by the first frame whose
   *          lattice_vocab contains them (decoder_dynamic.py:30-46); vocab_frame_ptr holds, per
   *          sentence, sent_len[s]+2 offsets (at frame_ptr_off[s]) relative to vocab_ptr[s]: entries
   *          [0, vocab_frame_ptr[i+1]) form lattice_vocab[i].  dup_ids (per sentence, dup_ptr) are the
   *          extra duplicate entries the reference keeps in lattice_vocab[0] only. */
  const int64_t* vocab_ptr;        /* [n_sent+1] or NULL */
  const int32_t* vocab_ids;
  const int32_t* vocab_frame_ptr;  /* DYNAMIC only */
  const int64_t* dup_ptr;          /* DYNAMIC only, [n_sent+1] */
  const int32_t* dup_ids;
} jlm_lattice_batch;

/* ------------------------------------------------------------------------------------------------
 * Native lattice builder (host code): Decoder._build_lattice (decoder.py:79-135) and the vocabulary
 * selection lists (decoder.py:137-151, decoder_dynamic.py:30-46) for a batch of sentences, emitted as
 * the CSR arrays above.  Needs no GPU.
 *
 * jlm_lexicon_create takes the dictionary side as the reference prepares it in Decoder.__init__
 * (decoder.py:55-74): for every key of reading_dict.pkl its UTF-32 code points and the ids (w2i) of its
 * in-vocabulary words in `sorted(lexicon ids)` order (decoder.py:97-103; out-of-vocabulary words
 * already skipped).  Entry e of word_ids is "lexicon entry e"; node_entry reports it per lattice node so
 * the host can recover the word string (-1: '<eos>', -2: the '<unk>' fallback whose word is the kana
 * at node_start, decoder.py:129-130). */
int32_t jlm_lexicon_create(int32_t n_readings, const int64_t* reading_ptr /* [n+1] */,
                           const uint32_t* reading_chars, const int64_t* word_ptr /* [n+1] */,
                           const int32_t* word_ids, int32_t eos_id, int32_t unk_id, jlm_lexicon** out);
int32_t jlm_lexicon_destroy(jlm_lexicon* lex);
/* Sentence s is text[text_ptr[s] .. text_ptr[s+1]) (UTF-32).  mode = JLM_DECODE_*: which vocabulary
 * lists to emit.  extra_ids [n_sent, n_extra] are the `samples` word ids of each sentence
 * (top_sampling: 0..samples-1; random_sampling: the host's np.random.randint draw). */
int32_t jlm_lattice_build(const jlm_lexicon* lex, int32_t n_sent, const int64_t* text_ptr, const uint32_t* text,
                          int32_t mode, int32_t n_extra, const int32_t* extra_ids, jlm_lattice** out);
/* Borrowed pointers into the lattice object (valid until jlm_lattice_destroy). */
int32_t jlm_lattice_view(const jlm_lattice* lat, jlm_lattice_batch* view, const int32_t** node_entry,
                         int64_t* n_nodes);
int32_t jlm_lattice_destroy(jlm_lattice* lat);

/* n-best output of Decoder.decode (decoder.py:237-241) for every sentence. */
typedef struct {
  int32_t top_n;        /* capacity per sentence */
  int32_t max_len;      /* capacity of one path, >= max(sent_len)+1 */
  double* scores;       /* [n_sent, top_n] neg_log_prob, ascending */
  int32_t* n_paths;     /* [n_sent] min(topN, paths in the last frame) */
  int32_t* path_len;    /* [n_sent, top_n] nodes per path, <eos> node included */
  int32_t* path_nodes;  /* [n_sent, top_n, max_len] absolute node indices, first to last */
} jlm_nbest;

/* Decoder.decode / DynamicDecoder.decode (decoder.py:220-241, decoder_dynamic.py:177-194) for a
 * batch of independent sentences decoded in lock-step: plan + host->device copy + all frames +
 * device->host copy of the n-best lists.  1 <= beam_width <= JLM_MAX_BEAM, or JLM_BEAM_UNLIMITED.
 * A batch of ONE sentence on the float64 back end (JLM_BACKEND_EXACT, or AUTO below 512 rows per frame) with
 * beam_width <= 64, full softmax and a tied / segmented projection runs its whole frame loop in one cooperative
 * kernel (2 launches per call instead of ~8 per frame) unless timers are enabled on the batch. */
int32_t jlm_decode_batch(jlm_handle* h, const jlm_lattice_batch* lat, int32_t beam_width, int32_t top_n,
                         int32_t mode, int32_t backend, jlm_nbest* out);

/* Decoder.decode / DynamicDecoder.decode for a batch of kana strings in one call: jlm_lattice_build +
 * jlm_decode_batch, pipelined over n_chunks sub-batches (0 = automatic) so that the host work of chunk
 * c+1 (lattice, plan, kernel enqueue) overlaps the device work of chunk c.  Paths come back as lexicon
 * entries (see jlm_lexicon_create; -1 '<eos>', -2 '<unk>' whose word is the kana at path_start) so no
 * lattice object outlives the call.  Arguments as jlm_lattice_build / jlm_decode_batch. */
typedef struct {
  int32_t top_n;        /* capacity per sentence */
  int32_t max_len;      /* capacity of one path, >= max(kana per sentence)+1 */
  double* scores;       /* [n_sent, top_n] */
  int32_t* n_paths;     /* [n_sent] */
  int32_t* path_len;    /* [n_sent, top_n] */
  int32_t* path_entry;  /* [n_sent, top_n, max_len] */
  int32_t* path_start;  /* [n_sent, top_n, max_len] start frame of each node (-1 for '<eos>') */
} jlm_text_nbest;
int32_t jlm_decode_texts(jlm_handle* h, const jlm_lexicon* lex, int32_t n_sent, const int64_t* text_ptr,
                         const uint32_t* text, int32_t beam_width, int32_t top_n, int32_t mode, int32_t n_extra,
                         const int32_t* extra_ids, int32_t backend, int32_t n_chunks, jlm_text_nbest* out,
                         struct jlm_batch_info_s* info /* nullable: totals over the chunks, enables the timers */);

/* LM state pool: LSTM_Model.predict_with_context (decoder/model.py:195-198) for host-driven searches such as
 * CharRNNDecoder (decoder/decoder.py:244-341), with the (hidden, cell) pairs resident on the device and named by slot.
 * jlm_pool_step: row k reads the state in slot src[k] (-1 = the zero state of Path.__init__, decoder.py:35-36),
 * feeds word index[k] (_lstm_cell, model.py:125-139) and leaves the new state, its projection and the log-normaliser
 * of its softmax (project + softmax, model.py:141-193,15-20) in slot *first_slot + k.  jlm_pool_nll: -log p(col[k] |
 * state slot[k]), i.e. Path.append_node's -math.log(transition_probs[0][idx]) (decoder.py:43-49); -y for self-normalised
 * models.  Only indices go down and one float64 per pair comes back.  jlm_pool_reset forgets every state. */
typedef struct jlm_pool jlm_pool;
int32_t jlm_pool_create(jlm_handle* h, int64_t capacity /* states */, jlm_pool** out);
int32_t jlm_pool_destroy(jlm_pool* p);
int32_t jlm_pool_reset(jlm_pool* p);
int32_t jlm_pool_step(jlm_pool* p, int32_t n, const int32_t* src, const int32_t* index, int64_t* first_slot);
int32_t jlm_pool_nll(jlm_pool* p, int32_t n, const int32_t* slot, const int32_t* col, double* out);
int32_t jlm_pool_get_state(jlm_pool* p, int64_t slot, int32_t count, double* h_out /* [count,H] nullable */,
                           double* c_out /* nullable */);

/* jlm_decode_texts split in two for streaming callers (a server decoding batch after batch): submit
 * builds the lattices and the plan, copies the plan to the device and enqueues every frame plus the
 * n-best device->host copy, then returns without waiting; collect waits for THAT job only, fills `out`
 * and frees the job.  Submitting job k+1 before collecting job k hides all host work behind the device
 * work of job k (one handle, one stream, jobs complete in submission order).  The text buffers may be
 * reused as soon as submit returns; `timers` != 0 enables the CUDA-event buckets reported by collect.
 * collect consumes the job even when it fails; cancel waits for and discards a job. */
typedef struct jlm_text_job jlm_text_job;
int32_t jlm_decode_texts_submit(jlm_handle* h, const jlm_lexicon* lex, int32_t n_sent, const int64_t* text_ptr,
                                const uint32_t* text, int32_t beam_width, int32_t top_n, int32_t mode,
                                int32_t n_extra, const int32_t* extra_ids, int32_t backend, int32_t n_chunks,
                                int32_t timers, jlm_text_job** job);
int32_t jlm_decode_texts_collect(jlm_text_job* job, jlm_text_nbest* out, struct jlm_batch_info_s* info /* nullable */);
int32_t jlm_decode_texts_cancel(jlm_text_job* job);

/* The same call split in three so a benchmark can time the device part with the lattices already
 * resident in HBM: upload (plan + H2D), run (enqueue every frame, asynchronous), fetch (D2H). */
int32_t jlm_batch_upload(jlm_handle* h, const jlm_lattice_batch* lat, int32_t beam_width, int32_t top_n,
                         int32_t mode, int32_t backend, jlm_batch** out);
int32_t jlm_batch_run(jlm_batch* b);
int32_t jlm_batch_fetch(jlm_batch* b, jlm_nbest* out);
/* enqueue the n-best device->host copy behind the batch's kernels without waiting; jlm_batch_fetch and
 * jlm_batch_destroy then wait for this batch's completion event only, not for the whole stream */
int32_t jlm_batch_fetch_async(jlm_batch* b);
int32_t jlm_batch_destroy(jlm_batch* b);

/* Introspection used by the parity tests and the benchmark. */
typedef struct jlm_batch_info_s {
  int64_t n_slots;        /* beam entries over all sentences and frames (= LM rows stepped) */
  int64_t n_candidates;   /* expanded candidates scored over all frames */
  int64_t n_nodes;
  int32_t n_steps;        /* lock-step frames */
  int32_t backend;        /* resolved JLM_BACKEND_* */
  int64_t kernel_launches;/* kernels enqueued by the last jlm_batch_run */
  int64_t h2d_bytes;      /* bytes copied by jlm_batch_upload */
  int64_t d2h_bytes;      /* bytes copied by jlm_batch_fetch */
  float ms_lstm;          /* CUDA-event totals of the last run when timers are enabled */
  float ms_softmax;
  float ms_beam;
  /* tensor-core back end, timers enabled: CUDA-event time spent inside the two GEMM kernels alone
   * (gate GEMM + LSTM epilogue; output projection + online-LSE epilogue), their launch counts and
   * the LM rows they processed, for roofline accounting. */
  float ms_gate_gemm;
  float ms_proj_gemm;
  int32_t n_gate_launches;
  int32_t n_proj_launches;
  int32_t beam_width;     /* paths per frame the per-frame arrays of jlm_batch_get_beams are strided by: the
                             beam_width of the call, or the widest frame for JLM_BEAM_UNLIMITED */
  int32_t n_guard_flagged;/* near-tie guard (after jlm_batch_fetch): sentences with a rank decision below the bound */
  double guard_min_gap;   /* smallest gap any rank decision of the batch rested on (guard enabled), else 0 */
  double guard_eps;       /* the bound in force for this batch (0: guard off) */
  int32_t n_guard_pairs;  /* near-tied rank decisions whose two paths were re-scored in float64 */
  int32_t n_guard_rerun;  /* sentences re-decoded on the float64 back end (a re-scored pair contradicted the
                             tensor-core ranking, mass ties, or a vocabulary-selection mode) */
} jlm_batch_info;
int32_t jlm_batch_get_info(jlm_batch* b, jlm_batch_info* info);
/* CUDA events between the kernels of every frame (ms_lstm / ms_softmax / ms_beam buckets).  A one-sentence batch then
 * takes the per-frame launch path instead of the cooperative single-sentence kernel. */
int32_t jlm_batch_enable_timers(jlm_batch* b, int32_t on);

/* Per-frame beams of one sentence after jlm_batch_run (+ synchronize): for frame t the entries
 * [t*beam_width, t*beam_width + count[t]) (beam_width as reported by jlm_batch_get_info).  parent_rank/parent_frame identify the back-pointer
 * (-1 for the <eos> path); node is the absolute node index; lse is the row's log-sum-exp.
 * h_out/c_out (nullable) receive the float64 state rows [ (T+1)*beam_width, H ]. */
int32_t jlm_batch_get_beams(jlm_batch* b, int32_t sentence, int32_t* count, double* score,
                            int32_t* parent_frame, int32_t* parent_rank, int32_t* node, double* lse,
                            double* h_out, double* c_out);

/* Tensor-core GEMM self-test hook (tests only): C[M,N] = A[M,K].B[N,K]^T through the split-fp16
 * tcgen05 kernel, float32 in/out on the host. */
int32_t jlm_tc_gemm_selftest(jlm_handle* h, const float* A, const float* B, int32_t M, int32_t N,
                             int32_t K, float* C, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* JLM_B200_H */
