// Native kana lattice builder: Decoder._build_lattice (decoder/decoder.py:79-135) and the
// vocabulary-selection lists (decoder/decoder.py:137-151, decoder/decoder_dynamic.py:30-46) for a
// batch of sentences, emitted directly as the CSR arrays jlm_decode_batch consumes.
//
// The reference scans every substring input[i:i+j+1] against reading_dict and creates one Node per
// in-vocabulary word of the reading, in `sorted(lexicon ids)` order, appended to the frame the reading
// ends at; a frame that is still empty after the j==0 probe gets the '<unk>' fallback node.  The
// dictionary side (reading -> in-vocabulary word ids in lexicon-id order) is prepared once by the host
// (jlm_lexicon_create); here it is an open-addressing hash over UTF-32 code points with an
// incrementally extended FNV-1a hash per start position, so a sentence costs O(T * max_reading) probes.
#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/jlm_b200.h"

void jlm_set_error(const char* fmt, ...);

#define LAT_REQUIRE(cond, ...)      \
  do {                              \
    if (!(cond)) {                  \
      jlm_set_error(__VA_ARGS__);   \
      return 1;                     \
    }                               \
  } while (0)

struct jlm_lexicon {
  int32_t n_readings = 0;
  int32_t eos_id = 0, unk_id = 0;
  int32_t max_reading = 0;
  std::vector<int64_t> reading_ptr;   // [n+1] into chars
  std::vector<uint32_t> chars;
  std::vector<int64_t> word_ptr;      // [n+1] into word_ids; entry index == position in word_ids
  std::vector<int32_t> word_ids;
  struct Slot { uint32_t tag; int32_t idx; };   // idx -1 = empty; tag = high hash bits of the slot's reading
  std::vector<Slot> table;            // open addressing; one cache line access per probe
  std::vector<uint64_t> bloom;        // 1 bit per hash value (cache resident): most substrings are not readings
  uint64_t mask = 0, bloom_mask = 0;
};

struct LatPart;   // per-thread scratch of the build, kept with the object so its buffers are reused

struct jlm_lattice {
  std::vector<LatPart>* parts = nullptr;
  int32_t n_sent = 0;
  std::vector<int32_t> sent_len;
  std::vector<int64_t> frame_ptr_off, frame_ptr;
  std::vector<int32_t> node_start, node_word, node_entry;
  std::vector<int64_t> vocab_ptr, dup_ptr;
  std::vector<int32_t> vocab_ids, vocab_frame_ptr, dup_ids;
  int32_t mode = 0;
};

namespace {
constexpr uint64_t FNV_OFF = 1469598103934665603ull, FNV_PRIME = 1099511628211ull;
inline uint64_t fnv_step(uint64_t h, uint32_t c) {
  h ^= (uint64_t)c;
  h *= FNV_PRIME;
  return h ^ (h >> 29);
}
}  // namespace

extern "C" int32_t jlm_lexicon_create(int32_t n_readings, const int64_t* reading_ptr, const uint32_t* reading_chars,
                                      const int64_t* word_ptr, const int32_t* word_ids, int32_t eos_id,
                                      int32_t unk_id, jlm_lexicon** out) {
  LAT_REQUIRE(out && n_readings >= 0 && (n_readings == 0 || (reading_ptr && reading_chars && word_ptr)),
              "jlm_lexicon_create: bad argument");
  *out = nullptr;
  jlm_lexicon* L = new jlm_lexicon();
  L->n_readings = n_readings;
  L->eos_id = eos_id;
  L->unk_id = unk_id;
  if (n_readings) {
    L->reading_ptr.assign(reading_ptr, reading_ptr + n_readings + 1);
    L->chars.assign(reading_chars, reading_chars + reading_ptr[n_readings]);
    L->word_ptr.assign(word_ptr, word_ptr + n_readings + 1);
    if (word_ptr[n_readings]) L->word_ids.assign(word_ids, word_ids + word_ptr[n_readings]);
  } else {
    L->reading_ptr.assign(1, 0);
    L->word_ptr.assign(1, 0);
  }
  uint64_t cap = 16;
  while (cap < (uint64_t)n_readings * 2 + 2) cap <<= 1;
  L->mask = cap - 1;
  L->table.assign(cap, jlm_lexicon::Slot{0u, -1});
  uint64_t bits = (uint64_t)1 << 15;
  while (bits < (uint64_t)n_readings * 16 && bits < ((uint64_t)1 << 21)) bits <<= 1;
  L->bloom.assign(bits / 64, 0);
  L->bloom_mask = bits - 1;
  for (int32_t r = 0; r < n_readings; ++r) {
    const int64_t a = L->reading_ptr[r], b = L->reading_ptr[r + 1];
    if (b < a || L->word_ptr[r + 1] < L->word_ptr[r]) {
      delete L;
      jlm_set_error("jlm_lexicon_create: offsets of reading %d are not monotone", r);
      return 1;
    }
    L->max_reading = std::max<int32_t>(L->max_reading, (int32_t)(b - a));
    uint64_t h = FNV_OFF;
    for (int64_t k = a; k < b; ++k) h = fnv_step(h, L->chars[k]);
    uint64_t s = h & L->mask;
    while (L->table[s].idx >= 0) {
      const int32_t q = L->table[s].idx;
      const int64_t qa = L->reading_ptr[q], qb = L->reading_ptr[q + 1];
      if (qb - qa == b - a && (b == a || memcmp(&L->chars[qa], &L->chars[a], sizeof(uint32_t) * (b - a)) == 0)) {
        delete L;
        jlm_set_error("jlm_lexicon_create: duplicate reading at index %d", r);
        return 1;
      }
      s = (s + 1) & L->mask;
    }
    L->table[s] = jlm_lexicon::Slot{(uint32_t)(h >> 32), r};
    const uint64_t bb = (h >> 20) & L->bloom_mask;
    L->bloom[bb >> 6] |= (uint64_t)1 << (bb & 63);
  }
  *out = L;
  return 0;
}

extern "C" int32_t jlm_lexicon_destroy(jlm_lexicon* lex) {
  delete lex;
  return 0;
}

namespace {

struct Match { int32_t i, end, reading; };

}  // namespace

// per-thread scratch of the two-pass build
struct LatPart {
  std::vector<Match> matches;
  std::vector<int64_t> match_ptr;     // per sentence of the range (+1)
  std::vector<int32_t> fcount;        // per frame of every sentence of the range: nodes ending there
  std::vector<int32_t> vocab_ids, vocab_frame_ptr, dup_ids;
  std::vector<int64_t> vocab_len, dup_len;
  int32_t rc = 0;
  char err[200] = "";
  void reset() {
    matches.clear(); match_ptr.clear(); fcount.clear(); vocab_ids.clear(); vocab_frame_ptr.clear(); dup_ids.clear();
    vocab_len.clear(); dup_len.clear();
    rc = 0;
  }
};
typedef LatPart Part;

namespace {

// finished lattice objects are recycled: their (page-touched) buffers make the next build cheaper
std::mutex g_pool_mu;
std::vector<jlm_lattice*> g_pool;

// pass 1: probe every substring, remember the matches and how many nodes end at each frame
void scan_range(const jlm_lexicon* L, int32_t s_lo, int32_t s_hi, const int64_t* text_ptr, const uint32_t* text,
                Part* P, int64_t* node_total) {
  P->match_ptr.assign(1, 0);
  for (int32_t s = s_lo; s < s_hi; ++s) {
    const int64_t t0 = text_ptr[s], t1 = text_ptr[s + 1];
    if (t1 < t0 || t1 - t0 > (int64_t)1 << 24) {
      snprintf(P->err, sizeof(P->err), "jlm_lattice_build: bad text offsets for sentence %d", s);
      P->rc = 1;
      return;
    }
    const int32_t T = (int32_t)(t1 - t0);
    const uint32_t* tx = text + t0;
    const size_t f0 = P->fcount.size();
    P->fcount.resize(f0 + T + 1, 0);
    int32_t* fc = P->fcount.data() + f0;
    fc[0] = 1;                                                       // '<eos>', decoder.py:89-90
    for (int32_t i = 0; i < T; ++i) {
      uint64_t h = FNV_OFF;
      const int32_t jmax = std::min(T - i, L->max_reading);
      for (int32_t j = 0; j < jmax; ++j) {
        h = fnv_step(h, tx[i + j]);
        const uint64_t bb = (h >> 20) & L->bloom_mask;
        if (!((L->bloom[bb >> 6] >> (bb & 63)) & 1)) continue;
        uint64_t slot = h & L->mask;
        const uint32_t tg = (uint32_t)(h >> 32);
        while (L->table[slot].idx >= 0) {
          if (L->table[slot].tag == tg) {
            const int32_t q = L->table[slot].idx;
            const int64_t qa = L->reading_ptr[q];
            if (L->reading_ptr[q + 1] - qa == j + 1 && memcmp(&L->chars[qa], tx + i, sizeof(uint32_t) * (j + 1)) == 0) {
              const int32_t nw = (int32_t)(L->word_ptr[q + 1] - L->word_ptr[q]);
              if (nw) {
                P->matches.push_back({i, i + j + 1, q});
                fc[i + j + 1] += nw;
              }
              break;
            }
          }
          slot = (slot + 1) & L->mask;
        }
      }
    }
    int64_t total = 0;
    for (int32_t t = 0; t <= T; ++t) {
      if (fc[t] == 0) fc[t] = -1;          // empty frame -> the '<unk>' fallback node (decoder.py:129-130)
      total += fc[t] < 0 ? 1 : fc[t];
    }
    node_total[s] = total;
    P->match_ptr.push_back((int64_t)P->matches.size());
  }
}

// pass 2: place the nodes (frame-major; inside a frame start ascending, then lexicon-id order) and
// derive the vocabulary lists of the sentence from its finished CSR slice
void fill_range(const jlm_lexicon* L, int32_t s_lo, int32_t s_hi, const int64_t* text_ptr, int32_t mode, int32_t n_extra,
                const int32_t* extra_ids, const int64_t* node_off, jlm_lattice* lat, Part* P) {
  std::vector<int64_t> cursor;
  std::vector<int32_t> scratch, seen_sorted, fresh;
  size_t f0 = 0;
  for (int32_t s = s_lo; s < s_hi; ++s) {
    const int32_t T = (int32_t)(text_ptr[s + 1] - text_ptr[s]);
    const int32_t* fc = P->fcount.data() + f0;
    f0 += T + 1;
    lat->sent_len[s] = T;
    int64_t* fp = lat->frame_ptr.data() + lat->frame_ptr_off[s];
    cursor.resize(T + 1);
    int64_t pos = node_off[s];
    for (int32_t t = 0; t <= T; ++t) {
      fp[t] = pos;
      cursor[t] = pos;
      pos += fc[t] < 0 ? 1 : fc[t];
    }
    fp[T + 1] = pos;
    lat->node_start[fp[0]] = -1;
    lat->node_word[fp[0]] = L->eos_id;
    lat->node_entry[fp[0]] = -1;
    for (int32_t t = 1; t <= T; ++t)
      if (fc[t] < 0) {
        lat->node_start[fp[t]] = t - 1;
        lat->node_word[fp[t]] = L->unk_id;
        lat->node_entry[fp[t]] = -2;
      }
    const int32_t ls = s - s_lo;
    for (int64_t m = P->match_ptr[ls]; m < P->match_ptr[ls + 1]; ++m) {
      const Match& mt = P->matches[m];
      int64_t c = cursor[mt.end];
      for (int64_t e = L->word_ptr[mt.reading]; e < L->word_ptr[mt.reading + 1]; ++e, ++c) {
        lat->node_start[c] = mt.i;
        lat->node_word[c] = L->word_ids[e];
        lat->node_entry[c] = (int32_t)e;
      }
      cursor[mt.end] = c;
    }
    if (mode == JLM_DECODE_FULL) continue;
    const int32_t* extra = n_extra ? extra_ids + (int64_t)s * n_extra : nullptr;
    const int32_t* words = lat->node_word.data();
    if (mode == JLM_DECODE_STATIC_VOCAB) {
      // decoder.py:142-151: sorted(set(all node words)) (+ samples, re-sorted, de-duplicated)
      scratch.assign(words + fp[0], words + fp[T + 1]);
      for (int32_t k = 0; k < n_extra; ++k) scratch.push_back(extra[k]);
      std::sort(scratch.begin(), scratch.end());
      scratch.erase(std::unique(scratch.begin(), scratch.end()), scratch.end());
      P->vocab_ids.insert(P->vocab_ids.end(), scratch.begin(), scratch.end());
      P->vocab_len.push_back((int64_t)scratch.size());
    } else {
      // decoder_dynamic.py:30-46: lattice_vocab[0] = sorted(frame-0 words) + samples (duplicates kept),
      // lattice_vocab[i] = lattice_vocab[i-1] | words ending at i.  Emitted as columns ordered by first
      // appearance + per-frame boundaries + the duplicate entries of frame 0.
      scratch.assign(words + fp[0], words + fp[1]);
      std::sort(scratch.begin(), scratch.end());
      for (int32_t k = 0; k < n_extra; ++k) scratch.push_back(extra[k]);
      seen_sorted = scratch;
      std::sort(seen_sorted.begin(), seen_sorted.end());
      seen_sorted.erase(std::unique(seen_sorted.begin(), seen_sorted.end()), seen_sorted.end());
      size_t nd = 0;
      {  // duplicates: every occurrence beyond the first, in list order
        std::vector<char> used(seen_sorted.size(), 0);
        for (int32_t v : scratch) {
          const size_t k = std::lower_bound(seen_sorted.begin(), seen_sorted.end(), v) - seen_sorted.begin();
          if (used[k]) {
            P->dup_ids.push_back(v);
            ++nd;
          } else {
            used[k] = 1;
          }
        }
      }
      const size_t base = P->vocab_ids.size();
      P->vocab_ids.insert(P->vocab_ids.end(), seen_sorted.begin(), seen_sorted.end());
      P->vocab_frame_ptr.push_back(0);
      P->vocab_frame_ptr.push_back((int32_t)seen_sorted.size());
      for (int32_t t = 1; t <= T; ++t) {
        fresh.clear();
        for (int64_t n = fp[t]; n < fp[t + 1]; ++n)
          if (!std::binary_search(seen_sorted.begin(), seen_sorted.end(), words[n])) fresh.push_back(words[n]);
        std::sort(fresh.begin(), fresh.end());
        fresh.erase(std::unique(fresh.begin(), fresh.end()), fresh.end());
        P->vocab_ids.insert(P->vocab_ids.end(), fresh.begin(), fresh.end());
        const size_t mid = seen_sorted.size();
        seen_sorted.insert(seen_sorted.end(), fresh.begin(), fresh.end());
        std::inplace_merge(seen_sorted.begin(), seen_sorted.begin() + mid, seen_sorted.end());
        P->vocab_frame_ptr.push_back((int32_t)(P->vocab_ids.size() - base));
      }
      P->vocab_len.push_back((int64_t)(P->vocab_ids.size() - base));
      P->dup_len.push_back((int64_t)nd);
    }
  }
}

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

}  // namespace

extern "C" int32_t jlm_lattice_build(const jlm_lexicon* L, int32_t n_sent, const int64_t* text_ptr,
                                     const uint32_t* text, int32_t mode, int32_t n_extra, const int32_t* extra_ids,
                                     jlm_lattice** out) {
  LAT_REQUIRE(L && out && n_sent > 0 && text_ptr, "jlm_lattice_build: bad argument");
  LAT_REQUIRE(mode >= JLM_DECODE_FULL && mode <= JLM_DECODE_DYNAMIC, "jlm_lattice_build: bad mode %d", mode);
  LAT_REQUIRE(n_extra >= 0 && (n_extra == 0 || extra_ids), "jlm_lattice_build: extra ids missing");
  *out = nullptr;
  // sentences are independent: two passes over contiguous ranges on host threads (scan + count, then
  // fill at the offsets of a prefix sum), so the arrays are written once, in place
  int nthreads = (int)std::thread::hardware_concurrency();
  if (const char* e = getenv("JLM_HOST_THREADS")) nthreads = atoi(e);
  nthreads = std::max(1, std::min(std::min(nthreads, 8), n_sent / 128));
  static const bool dbg = getenv("JLM_DEBUG_TIMING") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](auto a, auto c) { return std::chrono::duration<double, std::milli>(c - a).count(); };
  const auto t_0 = now();
  jlm_lattice* lat = nullptr;
  {
    std::lock_guard<std::mutex> g(g_pool_mu);
    if (!g_pool.empty()) {
      lat = g_pool.back();
      g_pool.pop_back();
    }
  }
  if (!lat) lat = new jlm_lattice();
  if (!lat->parts) lat->parts = new std::vector<LatPart>();
  std::vector<Part>& part = *lat->parts;
  if ((int)part.size() < nthreads) part.resize(nthreads);
  for (auto& p : part) p.reset();
  std::vector<int64_t> node_off((size_t)n_sent + 1, 0);
  auto lo = [&](int k) { return (int32_t)((int64_t)n_sent * k / nthreads); };
  run_threads(nthreads, [&](int k) { scan_range(L, lo(k), lo(k + 1), text_ptr, text, &part[k], node_off.data() + 1); });
  for (int k = 0; k < nthreads; ++k)
    if (part[k].rc) {
      jlm_set_error("%s", part[k].err);
      jlm_lattice_destroy(lat);
      return 1;
    }
  const auto t_1 = now();
  lat->vocab_ptr.clear(); lat->vocab_ids.clear(); lat->vocab_frame_ptr.clear(); lat->dup_ptr.clear(); lat->dup_ids.clear();
  lat->n_sent = n_sent;
  lat->mode = mode;
  lat->sent_len.resize(n_sent);
  lat->frame_ptr_off.resize(n_sent);
  int64_t fpo = 0;
  for (int32_t s = 0; s < n_sent; ++s) {
    node_off[s + 1] += node_off[s];
    lat->frame_ptr_off[s] = fpo;
    fpo += (text_ptr[s + 1] - text_ptr[s]) + 2;
  }
  lat->frame_ptr.resize(fpo);
  lat->node_start.resize(node_off[n_sent]);
  lat->node_word.resize(node_off[n_sent]);
  lat->node_entry.resize(node_off[n_sent]);
  const auto t_2 = now();
  run_threads(nthreads, [&](int k) {
    fill_range(L, lo(k), lo(k + 1), text_ptr, mode, n_extra, extra_ids, node_off.data(), lat, &part[k]);
  });
  if (dbg)
    fprintf(stderr, "[jlm] lattice_build: %d threads, scan %.3f ms, alloc %.3f ms, fill %.3f ms\n", nthreads, ms(t_0, t_1),
            ms(t_1, t_2), ms(t_2, now()));
  if (mode != JLM_DECODE_FULL) {
    lat->vocab_ptr.assign(1, 0);
    if (mode == JLM_DECODE_DYNAMIC) lat->dup_ptr.assign(1, 0);
    for (int k = 0; k < nthreads; ++k) {
      const Part& p = part[k];
      lat->vocab_ids.insert(lat->vocab_ids.end(), p.vocab_ids.begin(), p.vocab_ids.end());
      for (int64_t n : p.vocab_len) lat->vocab_ptr.push_back(lat->vocab_ptr.back() + n);
      if (mode == JLM_DECODE_DYNAMIC) {
        lat->vocab_frame_ptr.insert(lat->vocab_frame_ptr.end(), p.vocab_frame_ptr.begin(), p.vocab_frame_ptr.end());
        lat->dup_ids.insert(lat->dup_ids.end(), p.dup_ids.begin(), p.dup_ids.end());
        for (int64_t n : p.dup_len) lat->dup_ptr.push_back(lat->dup_ptr.back() + n);
      }
    }
    if (mode == JLM_DECODE_DYNAMIC && lat->dup_ids.empty()) lat->dup_ids.push_back(0);   // keep the pointer non-null
  }
  *out = lat;
  return 0;
}

extern "C" int32_t jlm_lattice_view(const jlm_lattice* lat, jlm_lattice_batch* view, const int32_t** node_entry,
                                    int64_t* n_nodes) {
  LAT_REQUIRE(lat && view, "jlm_lattice_view: bad argument");
  memset(view, 0, sizeof(*view));
  view->n_sent = lat->n_sent;
  view->sent_len = lat->sent_len.data();
  view->frame_ptr_off = lat->frame_ptr_off.data();
  view->frame_ptr = lat->frame_ptr.data();
  view->node_start = lat->node_start.data();
  view->node_word = lat->node_word.data();
  if (lat->mode != JLM_DECODE_FULL) {
    view->vocab_ptr = lat->vocab_ptr.data();
    view->vocab_ids = lat->vocab_ids.data();
  }
  if (lat->mode == JLM_DECODE_DYNAMIC) {
    view->vocab_frame_ptr = lat->vocab_frame_ptr.data();
    view->dup_ptr = lat->dup_ptr.data();
    view->dup_ids = lat->dup_ids.data();
  }
  if (node_entry) *node_entry = lat->node_entry.data();
  if (n_nodes) *n_nodes = (int64_t)lat->node_word.size();
  return 0;
}

extern "C" int32_t jlm_lattice_destroy(jlm_lattice* lat) {
  if (!lat) return 0;
  {
    std::lock_guard<std::mutex> g(g_pool_mu);
    if (g_pool.size() < 4) {
      g_pool.push_back(lat);
      return 0;
    }
  }
  delete lat->parts;
  delete lat;
  return 0;
}
