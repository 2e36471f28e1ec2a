// Batch decode engine: host-side plan + device beam state shared by both arithmetic back ends.
#pragma once
#include "jlm_common.cuh"

struct StepPlan {
  int64_t row0;       // first beam slot of this lock-step frame (slots are frame-major)
  int rows_step;      // rows that take an LM step (sentences that continue past this frame)
  int rows_all;       // all beam entries of the frame
  int nact;           // sentences (sorted positions 0..nact) that have this frame
  int nstep;          // sentences that continue (prefix of the active ones)
  int64_t item0;      // first entry of start_items: lattice nodes that START at this frame
  int n_items;
  int64_t job0;       // first vocabulary SubsetJob of the step (one per stepped sentence)
  int max_vocab_cols;
  int64_t yv_elems;   // elements of the per-step vocab-logit scratch
};

struct DynJobInfo {
  int64_t vfp_off;    // into d_vfp (sentence's vocab_frame_ptr, T+2 entries)
  int32_t nv, nd, T, pad;
};

// Work item of the scoring kernel: everything a warp needs about one lattice node, gathered once per run
// by k_build_items so the scoring warps start their weight / T-row loads after ONE dependent load
// instead of three (start_items -> node_word, node_pfid, cand_pos -> bc, slot0).
struct __align__(16) ScoreItem {
  int32_t word;     // node_word
  int32_t rows;     // kept paths of the node's start frame (bc)
  int32_t node;
  int32_t pad;      // dynamic: frames the node spans (node_span), else 0
  int64_t ps0;      // first slot of the start frame
  int64_t cpos;     // first candidate slot of the node
};

struct BeamDev {
  // ---- plan (read-only on the device) ----
  int32_t* node_word = nullptr;
  int32_t* node_pfid = nullptr;     // frame id of the node's start frame, -1 for <eos>
  int32_t* node_span = nullptr;     // dynamic only: frames the node spans (end frame - start frame)
  int64_t* cand_pos = nullptr;      // first candidate slot of the node (candidates are END-frame major)
  int32_t* frame_lo = nullptr;      // per frame id: nodes ending here [lo, hi)
  int32_t* frame_hi = nullptr;
  int64_t* frame_cand_lo = nullptr; // per frame id: first candidate
  int32_t* frame_ncand = nullptr;   // per frame id: candidates (sum over its nodes of the parents' beam counts)
  int32_t* frame_minpf = nullptr;   // smallest parent frame id (dynamic chain sums)
  int32_t* bc = nullptr;            // beam count per frame id
  int64_t* slot0 = nullptr;         // first slot per frame id
  int64_t* fbase = nullptr;         // first frame id per sorted sentence
  int32_t* sent_T = nullptr;
  int32_t* start_items = nullptr;   // node ids grouped by the lock-step frame they start at
  ScoreItem* items = nullptr;       // start_items expanded on the device (k_build_items)
  int64_t n_items = 0;
  SubsetJob* vocab_jobs = nullptr;
  DynJobInfo* dyn_info = nullptr;
  int32_t* vocab_cols = nullptr;
  int32_t* vfp = nullptr;
  // ---- beam state ----
  double* slot_score = nullptr;
  double* slot_lse = nullptr;
  double* slot_cumy = nullptr;      // dynamic: sum of the path's logits
  double* dyn_lse = nullptr;        // dynamic: [slot, Tmax+1] LSE over lattice_vocab[i]
  double* dyn_chain = nullptr;      // dynamic: [slot, Tmax+1] sum of the path's LSEs over lattice_vocab[i]
  int32_t* slot_parent = nullptr;   // global slot of the parent, -1 for <eos>
  int32_t* slot_node = nullptr;
  int32_t* slot_word = nullptr;
  // candidates, contiguous per (sentence, frame) in the reference's enumeration order
  // (node order, then parent rank): candidate cand_pos[n] + r extends parent rank r with node n
  double* cand_val = nullptr;       // static: full path score ; dynamic: the transition's logit
  int32_t* cand_parent = nullptr;   // dynamic only: global slot of the candidate's parent path
  double* cand_score = nullptr;     // dynamic only: the candidate's full score under its END frame's vocabulary
  // ---- near-tie guard (tensor-core back end): per sorted sentence, written by the prune kernel ----
  double* guard_gap = nullptr;      // smallest gap between adjacent ranks 0..W of any frame (kept vs kept, kept vs best rejected)
  int32_t* guard_flag = nullptr;    // bit 0: some gap fell below the bound (records queued); bit 1: re-decode in float64
  int32_t* guard_n = nullptr;       // [1] near-tie records queued by the prune kernel
  int4* guard_rec = nullptr;        // [guard_cap] {sorted sentence, frame, candidate ordinal ranked first, ... second}
  int32_t* guard_paths = nullptr;   // [guard_cap][2][max_len + 1] {length, node ids first..last} of both candidates
  int32_t guard_cap = 0;
  // ---- n-best output ----
  double* out_score = nullptr;
  int32_t* out_npaths = nullptr;
  int32_t* out_len = nullptr;
  int32_t* out_nodes = nullptr;
};

struct TcBatchState;  // jlm_tc.cu

struct jlm_batch {
  jlm_handle* h = nullptr;
  int S = 0, W = 0, topN = 0, mode = 0, backend = 0, Tmax = 0, n_steps = 0;
  bool use_lse = true;
  bool dynamic = false;
  bool unlimited = false;        // beam_width=None: no sort, no prune (W = widest frame of the plan)
  bool counted = false;          // this batch is in jlm_handle::batches_unfetched
  // Near-tie guard.  The tensor-core back end carries fp32 state and fp32-accumulated logits, so two candidates
  // whose float64 scores differ by less than its error can come out in the other order.  With guard_eps > 0 the
  // prune kernel flags every sentence in which a rank decision (adjacent kept paths, or the last kept path against
  // the best rejected candidate) rests on a gap below guard_eps; jlm_batch_fetch re-decodes exactly those sentences
  // on the float64 back end and splices their results in.
  double guard_eps = 0.0;
  bool guard_all = false;        // guard every rank decision, not only those that can change the returned n-best
  struct GuardLattice* guard_lat = nullptr;   // host copy of the lattice arrays (caller's buffers may be gone by fetch)
  jlm_batch* rerun = nullptr;                 // the float64 batch of the flagged sentences (kept for jlm_batch_get_beams)
  std::vector<int> rerun_index;               // caller's sentence index -> sentence of `rerun`, -1 = not flagged
  int n_flagged = 0;             // sentences with at least one near-tie
  int n_pairs = 0;               // near-tied rank decisions re-scored in float64 (jlm_pool path scores)
  int n_lse_rows = 0;            // float64 log-sum-exp rows the re-scoring needed
  int n_rerun = 0;               // sentences re-decoded in float64 (a re-scored pair contradicted the ranking, or no cheap check applies)
  double min_gap = 0.0;
  std::vector<int> order;        // sorted position -> caller's sentence index
  std::vector<int> sent_T;       // by sorted position
  std::vector<int64_t> fbase;
  std::vector<int> bc;
  std::vector<int64_t> slot0;
  std::vector<StepPlan> steps;
  int64_t F = 0, N = 0, n_slots = 0, n_cand = 0, n_jobs = 0;
  int max_rows_step = 0;
  int64_t max_yv = 0;
  // Vocabulary-selection modes: the first n_shared entries of EVERY sentence's word list are the same ids (the
  // `samples` top words of top_sampling, decoder.py:144-149 / decoder_dynamic.py:37-43): their logits come from one
  // dense tensor-core GEMM per step (y0, row stride ldy0, float32, bias included) instead of a gather per sentence.
  int n_shared = 0;
  int64_t n_vocab_cols = 0;      // entries of all sentences' word lists (vocabulary-selection modes)
  const float* y0 = nullptr;
  int ldy0 = 0;
  int max_len = 0;
  DevBuf mem;
  BeamDev d;
  // exact back end state
  double* hx = nullptr;          // [n_slots, Hp]
  double* cx = nullptr;
  double* A = nullptr;           // [max_rows_step, Kg]
  double* G = nullptr;           // [max_rows_step, 4H]
  double* T = nullptr;           // [max_rows_step, Kt]
  double2* part = nullptr;
  int part_tiles = 0;
  double2* spart = nullptr;      // [sm_count, 16] per-CTA (max, sum exp) of the single-sentence kernel (k_single_f64)
  double* yv = nullptr;          // vocab logits scratch
  TcBatchState* tc = nullptr;
  // bookkeeping
  int64_t launches = 0, h2d_bytes = 0, d2h_bytes = 0;
  bool timers = false;
  bool ran = false;
  // completion: `done` is recorded after the batch's last enqueued operation (run, then the n-best D2H),
  // so fetch / destroy wait for THIS batch only and later batches on the stream keep the device busy
  cudaEvent_t done = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;   // side-stream scoring under the output GEMM
  HostBuf out_host;              // pinned n-best landing buffer, taken from / returned to h->out_pool
  bool d2h_queued = false;
  std::vector<cudaEvent_t> events;
  float ms_lstm = 0.f, ms_softmax = 0.f, ms_beam = 0.f;
  std::vector<cudaEvent_t> kev;  // tensor-core back end: 4 per step around the gate / projection GEMMs
  float ms_gate_gemm = 0.f, ms_proj_gemm = 0.f;
  int n_gate_launches = 0, n_proj_launches = 0;
};

// tensor-core back end hooks (jlm_tc.cu)
int32_t tc_batch_plan(jlm_batch* b, Arena& a);                 // called twice: dry run + real
// gate GEMM .. full-vocabulary LSE for step t; returns the fp32 stage-1 rows for the needed-word dots
int32_t tc_batch_lm_state(jlm_batch* b, int t, const float** T_out, int* ldt_out);   // gather, gate GEMM, stage-1
int32_t tc_batch_lm_lse(jlm_batch* b, int t);                                         // output GEMMs + LSE merge
int32_t tc_batch_get_state(jlm_batch* b, int64_t slot, int count, double* h_out, double* c_out);
int32_t tc_vocab_logits(jlm_batch* b, int t, double* out);   // 0 = done, 1 = error, 2 = shape not supported: use the float64 kernel
void tc_batch_free(jlm_batch* b);
