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
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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
  std::vector<int32_t> table;         // open addressing, -1 = empty, else reading index
  std::vector<uint32_t> tag;          // high hash bits of the slot's reading: rejects most misses without
                                      // touching the reading itself
  uint64_t mask = 0;
};

struct jlm_lattice {
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
  L->table.assign(cap, -1);
  L->tag.assign(cap, 0);
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
    while (L->table[s] >= 0) {
      const int32_t q = L->table[s];
      const int64_t qa = L->reading_ptr[q], qb = L->reading_ptr[q + 1];
      if (qb - qa == b - a && (b == a || memcmp(&L->chars[qa], &L->chars[a], sizeof(uint32_t) * (b - a)) == 0)) {
        delete L;
        jlm_set_error("jlm_lexicon_create: duplicate reading at index %d", r);
        return 1;
      }
      s = (s + 1) & L->mask;
    }
    L->table[s] = r;
    L->tag[s] = (uint32_t)(h >> 32);
  }
  *out = L;
  return 0;
}

extern "C" int32_t jlm_lexicon_destroy(jlm_lexicon* lex) {
  delete lex;
  return 0;
}

namespace {

// Builds sentences [s_lo, s_hi) into `lat` with node / vocabulary offsets relative to this chunk.
int32_t build_range(const jlm_lexicon* L, int32_t s_lo, int32_t s_hi, const int64_t* text_ptr, const uint32_t* text,
                    int32_t mode, int32_t n_extra, const int32_t* extra_ids, jlm_lattice* lat, char* err, size_t errn) {
  struct Tmp { int32_t start, word, entry; };
  std::vector<std::vector<Tmp>> frames;
  std::vector<int32_t> scratch, seen_sorted, fresh;
  if (mode != JLM_DECODE_FULL) lat->vocab_ptr.assign(1, 0);
  if (mode == JLM_DECODE_DYNAMIC) lat->dup_ptr.assign(1, 0);
  {
    const size_t chars = (size_t)(text_ptr[s_hi] - text_ptr[s_lo]);
    lat->node_start.reserve(chars * 12);
    lat->node_word.reserve(chars * 12);
    lat->node_entry.reserve(chars * 12);
    lat->frame_ptr.reserve(chars + 2 * (size_t)(s_hi - s_lo));
  }
  for (int32_t s = s_lo; s < s_hi; ++s) {
    const int64_t t0 = text_ptr[s], t1 = text_ptr[s + 1];
    if (t1 < t0 || t1 - t0 > (int64_t)1 << 24) {
      snprintf(err, errn, "jlm_lattice_build: bad text offsets for sentence %d", s);
      return 1;
    }
    const int32_t T = (int32_t)(t1 - t0);
    const uint32_t* tx = text + t0;
    if ((int32_t)frames.size() < T + 1) frames.resize(T + 1);
    for (int32_t t = 0; t <= T; ++t) frames[t].clear();
    frames[0].push_back({-1, L->eos_id, -1});                       // decoder.py:89-90
    for (int32_t i = 0; i < T; ++i) {
      uint64_t h = FNV_OFF;
      const int32_t jmax = std::min(T - i, L->max_reading);
      for (int32_t j = 0; j < jmax; ++j) {
        h = fnv_step(h, tx[i + j]);
        uint64_t slot = h & L->mask;
        const uint32_t tg = (uint32_t)(h >> 32);
        while (L->table[slot] >= 0) {
          const int32_t q = L->table[slot];
          const int64_t qa = L->tag[slot] == tg ? L->reading_ptr[q] : 0;
          if (L->tag[slot] == tg && L->reading_ptr[q + 1] - qa == j + 1 &&
              memcmp(&L->chars[qa], tx + i, sizeof(uint32_t) * (j + 1)) == 0) {
            std::vector<Tmp>& end = frames[i + j + 1];
            for (int64_t e = L->word_ptr[q]; e < L->word_ptr[q + 1]; ++e)   // lexicon-id order, OOV already dropped
              end.push_back({i, L->word_ids[e], (int32_t)e});
            break;
          }
          slot = (slot + 1) & L->mask;
        }
        if (j == 0 && frames[i + 1].empty()) frames[i + 1].push_back({i, L->unk_id, -2});   // decoder.py:129-130
      }
      if (jmax == 0 && frames[i + 1].empty()) frames[i + 1].push_back({i, L->unk_id, -2});
    }
    // CSR
    lat->sent_len.push_back(T);
    lat->frame_ptr_off.push_back((int64_t)lat->frame_ptr.size());
    lat->frame_ptr.push_back((int64_t)lat->node_word.size());
    for (int32_t t = 0; t <= T; ++t) {
      for (const Tmp& n : frames[t]) {
        lat->node_start.push_back(n.start);
        lat->node_word.push_back(n.word);
        lat->node_entry.push_back(n.entry);
      }
      lat->frame_ptr.push_back((int64_t)lat->node_word.size());
    }
    const int32_t* extra = n_extra ? extra_ids + (int64_t)s * n_extra : nullptr;
    if (mode == JLM_DECODE_STATIC_VOCAB) {
      // decoder.py:142-151: sorted(set(all node words)) (+ samples, re-sorted, de-duplicated)
      scratch.clear();
      for (int32_t t = 0; t <= T; ++t)
        for (const Tmp& n : frames[t]) scratch.push_back(n.word);
      for (int32_t k = 0; k < n_extra; ++k) scratch.push_back(extra[k]);
      std::sort(scratch.begin(), scratch.end());
      scratch.erase(std::unique(scratch.begin(), scratch.end()), scratch.end());
      lat->vocab_ids.insert(lat->vocab_ids.end(), scratch.begin(), scratch.end());
      lat->vocab_ptr.push_back((int64_t)lat->vocab_ids.size());
    } else if (mode == JLM_DECODE_DYNAMIC) {
      // decoder_dynamic.py:30-46: lattice_vocab[0] = sorted(frame-0 words) + samples (duplicates kept),
      // lattice_vocab[i] = lattice_vocab[i-1] | words ending at i.  Emitted as columns ordered by first
      // appearance + per-frame boundaries + the duplicate entries of frame 0.
      scratch.clear();
      for (const Tmp& n : frames[0]) scratch.push_back(n.word);
      std::sort(scratch.begin(), scratch.end());
      for (int32_t k = 0; k < n_extra; ++k) scratch.push_back(extra[k]);
      seen_sorted = scratch;
      std::sort(seen_sorted.begin(), seen_sorted.end());
      seen_sorted.erase(std::unique(seen_sorted.begin(), seen_sorted.end()), seen_sorted.end());
      {  // duplicates: every occurrence beyond the first, in list order
        std::vector<char> used(seen_sorted.size(), 0);
        for (int32_t v : scratch) {
          const size_t k = std::lower_bound(seen_sorted.begin(), seen_sorted.end(), v) - seen_sorted.begin();
          if (used[k]) lat->dup_ids.push_back(v); else used[k] = 1;
        }
      }
      const int64_t base = (int64_t)lat->vocab_ids.size();
      lat->vocab_ids.insert(lat->vocab_ids.end(), seen_sorted.begin(), seen_sorted.end());
      lat->vocab_frame_ptr.push_back(0);
      lat->vocab_frame_ptr.push_back((int32_t)seen_sorted.size());
      for (int32_t t = 1; t <= T; ++t) {
        fresh.clear();
        for (const Tmp& n : frames[t])
          if (!std::binary_search(seen_sorted.begin(), seen_sorted.end(), n.word)) fresh.push_back(n.word);
        std::sort(fresh.begin(), fresh.end());
        fresh.erase(std::unique(fresh.begin(), fresh.end()), fresh.end());
        lat->vocab_ids.insert(lat->vocab_ids.end(), fresh.begin(), fresh.end());
        const size_t mid = seen_sorted.size();
        seen_sorted.insert(seen_sorted.end(), fresh.begin(), fresh.end());
        std::inplace_merge(seen_sorted.begin(), seen_sorted.begin() + mid, seen_sorted.end());
        lat->vocab_frame_ptr.push_back((int32_t)((int64_t)lat->vocab_ids.size() - base));
      }
      lat->vocab_ptr.push_back((int64_t)lat->vocab_ids.size());
      lat->dup_ptr.push_back((int64_t)lat->dup_ids.size());
    }
  }
  return 0;
}

template <class T>
void append(std::vector<T>& dst, const std::vector<T>& src, T add, size_t skip = 0) {
  const size_t o = dst.size();
  dst.resize(o + src.size() - skip);
  for (size_t i = skip; i < src.size(); ++i) dst[o + i - skip] = src[i] + add;
}

}  // namespace

extern "C" int32_t jlm_lattice_build(const jlm_lexicon* L, int32_t n_sent, const int64_t* text_ptr,
                                     const uint32_t* text, int32_t mode, int32_t n_extra, const int32_t* extra_ids,
                                     jlm_lattice** out) {
  LAT_REQUIRE(L && out && n_sent > 0 && text_ptr, "jlm_lattice_build: bad argument");
  LAT_REQUIRE(mode >= JLM_DECODE_FULL && mode <= JLM_DECODE_DYNAMIC, "jlm_lattice_build: bad mode %d", mode);
  LAT_REQUIRE(n_extra >= 0 && (n_extra == 0 || extra_ids), "jlm_lattice_build: extra ids missing");
  *out = nullptr;
  // sentences are independent: build contiguous chunks on host threads, then concatenate
  int nthreads = (int)std::thread::hardware_concurrency();
  if (const char* e = getenv("JLM_HOST_THREADS")) nthreads = atoi(e);
  nthreads = std::max(1, std::min(std::min(nthreads, 8), n_sent / 512));
  std::vector<jlm_lattice> part(nthreads);
  std::vector<int32_t> rc(nthreads, 0);
  std::vector<std::vector<char>> err(nthreads, std::vector<char>(256, 0));
  auto work = [&](int k) {
    const int32_t lo = (int32_t)((int64_t)n_sent * k / nthreads), hi = (int32_t)((int64_t)n_sent * (k + 1) / nthreads);
    rc[k] = build_range(L, lo, hi, text_ptr, text, mode, n_extra, extra_ids, &part[k], err[k].data(), err[k].size());
  };
  if (nthreads == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int k = 0; k < nthreads; ++k) th.emplace_back(work, k);
    for (auto& t : th) t.join();
  }
  for (int k = 0; k < nthreads; ++k)
    if (rc[k]) {
      jlm_set_error("%s", err[k].data());
      return 1;
    }
  jlm_lattice* lat = new jlm_lattice();
  if (nthreads == 1) {
    *lat = std::move(part[0]);
  } else {
    if (mode != JLM_DECODE_FULL) lat->vocab_ptr.assign(1, 0);
    if (mode == JLM_DECODE_DYNAMIC) lat->dup_ptr.assign(1, 0);
    for (int k = 0; k < nthreads; ++k) {
      const jlm_lattice& p = part[k];
      const int64_t node0 = (int64_t)lat->node_word.size(), fp0 = (int64_t)lat->frame_ptr.size();
      append(lat->sent_len, p.sent_len, 0);
      append(lat->frame_ptr_off, p.frame_ptr_off, fp0);
      append(lat->frame_ptr, p.frame_ptr, node0);
      append(lat->node_start, p.node_start, 0);
      append(lat->node_word, p.node_word, 0);
      append(lat->node_entry, p.node_entry, 0);
      if (mode != JLM_DECODE_FULL) {
        append(lat->vocab_ptr, p.vocab_ptr, (int64_t)lat->vocab_ids.size(), 1);
        append(lat->vocab_ids, p.vocab_ids, 0);
      }
      if (mode == JLM_DECODE_DYNAMIC) {
        append(lat->vocab_frame_ptr, p.vocab_frame_ptr, 0);
        append(lat->dup_ptr, p.dup_ptr, (int64_t)lat->dup_ids.size(), 1);
        append(lat->dup_ids, p.dup_ids, 0);
      }
    }
  }
  lat->n_sent = n_sent;
  lat->mode = mode;
  if (mode == JLM_DECODE_DYNAMIC && lat->dup_ids.empty()) lat->dup_ids.push_back(0);   // keep the pointer non-null
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
  delete lat;
  return 0;
}
