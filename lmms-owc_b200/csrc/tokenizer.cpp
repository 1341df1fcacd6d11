// tokenizer.cpp - host-side WordPiece tokenisation and packing.
//
// Tokenisation stays on the host (north star). The reference gets it from AutoTokenizer at
// src/data/pipelines/text/_text.py:193-195 (BertTokenizerFast: BertNormalizer -> BertPreTokenizer -> WordPiece ->
// "[CLS] A [SEP]" template, truncation to the model length). Once the device side scores millions of rows per
// second, that call - and the Python lists it returns - is the end-to-end bottleneck, so this file restates the
// same pipeline for the input that dominates the path (ASCII label strings) as multi-threaded C++ that writes
// the packed layout `owc_encode` consumes directly: ids back to back + cu_seqlens, no padding, no Python objects.
//
// Exactness contract: a row is tokenised here only if every byte is ASCII and it contains no literal special
// token ("[CLS]", "<s>", ...). Any other row is FLAGGED (length -1) and the Python host tokenises it with the
// `tokenizers` backend - Unicode normalisation (NFD accent stripping, CJK spacing, Unicode categories) is not
// re-implemented. For ASCII the BERT rules reduce to (transformers tokenization_bert.py BasicTokenizer /
// WordpieceTokenizer; tokenizers BertNormalizer / BertPreTokenizer):
//   clean_text   drop NUL and control characters other than \t \n \r; those three and ' ' are whitespace
//   lowercase    ASCII tolower (no accents to strip)
//   pre-tokenise split on whitespace; every ASCII punctuation byte (33-47, 58-64, 91-96, 123-126) is its own word
//   WordPiece    greedy longest-match-first, "##" continuation pieces, words of more than 100 characters or with
//                an unmatched remainder become ONE [UNK]
//   template     cls + pieces + sep, truncated to max_length tokens in total (pieces are cut, specials kept)
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/owc_b200.h"

namespace {

// open-addressing hash table from byte strings to token ids (FNV-1a, linear probing, power-of-two capacity)
struct PieceTable {
  struct Slot {
    uint32_t off = 0, len = 0;
    int32_t id = -1;
  };
  std::vector<Slot> slots;
  std::string pool;
  uint32_t mask = 0;
  int max_len = 0;

  static uint64_t hash(const char* p, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) {
      h ^= static_cast<unsigned char>(p[i]);
      h *= 1099511628211ull;
    }
    return h ^ (h >> 29);
  }
  void init(size_t n_items) {
    size_t cap = 16;
    while (cap < n_items * 2 + 2) cap <<= 1;
    slots.assign(cap, Slot{});
    mask = static_cast<uint32_t>(cap - 1);
  }
  void insert(const char* p, size_t n, int32_t id) {
    uint32_t i = static_cast<uint32_t>(hash(p, n)) & mask;
    while (slots[i].id >= 0) {
      if (slots[i].len == n && memcmp(pool.data() + slots[i].off, p, n) == 0) { slots[i].id = id; return; }   // last id wins, like {w: i for i, w in enumerate(vocab)}
      i = (i + 1) & mask;
    }
    slots[i].off = static_cast<uint32_t>(pool.size());
    slots[i].len = static_cast<uint32_t>(n);
    slots[i].id = id;
    pool.append(p, n);
    max_len = std::max(max_len, static_cast<int>(n));
  }
  int32_t find(const char* p, size_t n) const {
    uint32_t i = static_cast<uint32_t>(hash(p, n)) & mask;
    while (slots[i].id >= 0) {
      if (slots[i].len == n && memcmp(pool.data() + slots[i].off, p, n) == 0) return slots[i].id;
      i = (i + 1) & mask;
    }
    return -1;
  }
};

inline bool is_space(unsigned char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }
inline bool is_punct(unsigned char c) {
  return (c >= 33 && c <= 47) || (c >= 58 && c <= 64) || (c >= 91 && c <= 96) || (c >= 123 && c <= 126);
}
inline bool is_dropped(unsigned char c) { return c == 0 || (c < 32 && !is_space(c)) || c == 127; }

}  // namespace

struct owc_tokenizer {
  PieceTable first, cont;              // pieces that start a word / "##" continuation pieces (stored without the prefix)
  std::vector<std::string> specials;   // literal special tokens: a row containing one is left to the host library
  int32_t cls_id = -1, sep_id = -1, unk_id = -1;
  int32_t max_length = 512;

  // tokenises one row into out (cleared first); returns false if the row must go to the host library
  bool row(const char* s, size_t n, std::vector<int32_t>& out, std::string& word) const {
    out.clear();
    bool bracket = false;
    for (size_t i = 0; i < n; ++i) {
      const unsigned char c = static_cast<unsigned char>(s[i]);
      if (c >= 128) return false;
      bracket |= (c == '[' || c == '<');
    }
    if (bracket) {
      for (const std::string& sp : specials)
        if (sp.size() <= n && std::search(s, s + n, sp.begin(), sp.end()) != s + n) return false;
    }
    out.push_back(cls_id);
    const size_t budget = static_cast<size_t>(std::max(max_length - 1, 1));   // room for pieces before the closing special
    size_t i = 0;
    while (i < n && out.size() < budget) {
      const unsigned char c = static_cast<unsigned char>(s[i]);
      if (is_space(c) || is_dropped(c)) { ++i; continue; }
      word.clear();
      if (is_punct(c)) {
        word.push_back(static_cast<char>(c));
        ++i;
      } else {
        while (i < n) {
          const unsigned char d = static_cast<unsigned char>(s[i]);
          if (is_space(d) || is_punct(d)) break;
          if (!is_dropped(d)) word.push_back(static_cast<char>((d >= 'A' && d <= 'Z') ? d + 32 : d));
          ++i;
        }
        if (word.empty()) continue;
      }
      wordpiece(word, out);
    }
    if (out.size() > budget) out.resize(budget);
    out.push_back(sep_id);
    return true;
  }

  void wordpiece(const std::string& w, std::vector<int32_t>& out) const {
    const size_t n = w.size();
    if (n > 100) { out.push_back(unk_id); return; }
    const size_t mark = out.size();
    size_t start = 0;
    while (start < n) {
      const PieceTable& t = start == 0 ? first : cont;
      size_t end = std::min(n, start + static_cast<size_t>(t.max_len));
      int32_t id = -1;
      for (; end > start; --end) {
        id = t.find(w.data() + start, end - start);
        if (id >= 0) break;
      }
      if (id < 0) {            // an unmatched remainder turns the WHOLE word into one [UNK]
        out.resize(mark);
        out.push_back(unk_id);
        return;
      }
      out.push_back(id);
      start = end;
    }
  }
};

static thread_local char t_err[256] = "";
static int tfail(const char* msg) {
  strncpy(t_err, msg, sizeof(t_err) - 1);
  t_err[sizeof(t_err) - 1] = 0;
  return 1;
}

extern "C" const char* owc_tokenizer_last_error(void) { return t_err; }

extern "C" int32_t owc_tokenizer_create(const char* vocab_bytes, const int64_t* vocab_offsets, int32_t n_vocab, int32_t kind,
                                        int32_t max_length, owc_tokenizer** out) {
  if (!vocab_bytes || !vocab_offsets || !out || n_vocab <= 0) return tfail("owc_tokenizer_create: bad argument");
  if (kind != OWC_ARCH_BERT && kind != OWC_ARCH_MPNET) return tfail("owc_tokenizer_create: unknown kind");
  if (max_length < 3) return tfail("owc_tokenizer_create: max_length must be at least 3");
  *out = nullptr;
  owc_tokenizer* t = new (std::nothrow) owc_tokenizer();
  if (!t) return tfail("out of host memory");
  t->max_length = max_length;
  t->first.init(static_cast<size_t>(n_vocab));
  t->cont.init(static_cast<size_t>(n_vocab));
  const char* cls = kind == OWC_ARCH_BERT ? "[CLS]" : "<s>";
  const char* sep = kind == OWC_ARCH_BERT ? "[SEP]" : "</s>";
  if (kind == OWC_ARCH_BERT) t->specials = {"[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"};
  else t->specials = {"<s>", "<pad>", "</s>", "[UNK]", "<mask>", "<unk>"};
  for (int32_t i = 0; i < n_vocab; ++i) {
    const char* p = vocab_bytes + vocab_offsets[i];
    const size_t n = static_cast<size_t>(vocab_offsets[i + 1] - vocab_offsets[i]);
    if (n == strlen(cls) && memcmp(p, cls, n) == 0 && t->cls_id < 0) t->cls_id = i;
    if (n == strlen(sep) && memcmp(p, sep, n) == 0 && t->sep_id < 0) t->sep_id = i;
    if (n == 5 && memcmp(p, "[UNK]", 5) == 0 && t->unk_id < 0) t->unk_id = i;
    if (n > 2 && p[0] == '#' && p[1] == '#') t->cont.insert(p + 2, n - 2, i);
    else t->first.insert(p, n, i);
  }
  if (t->cls_id < 0 || t->sep_id < 0 || t->unk_id < 0) {
    delete t;
    return tfail("vocabulary lacks a special token (cls / sep / [UNK])");
  }
  *out = t;
  return 0;
}

extern "C" int32_t owc_tokenizer_destroy(owc_tokenizer* t) {
  delete t;
  return 0;
}

// texts: n rows as one byte blob + int64 offsets [n+1]. ids_out: capacity `ids_capacity` int32 (total bytes + 2n is
// always enough); lens_out: int32 [n] - tokens of the row, or -1 if the row must be tokenised by the host library
// (non-ASCII or a literal special token). Rows are laid out back to back in ids_out in row order, flagged rows
// contributing nothing. n_threads <= 0 picks the hardware concurrency (capped at 32).
extern "C" int32_t owc_tokenize_packed(const owc_tokenizer* t, const char* text_bytes, const int64_t* text_offsets, int64_t n,
                                       int32_t n_threads, int32_t* ids_out, int64_t ids_capacity, int32_t* lens_out,
                                       int64_t* n_ids_out) {
  if (!t || !lens_out || !n_ids_out || n < 0 || (n > 0 && (!text_bytes || !ids_out)))
    return tfail("owc_tokenize_packed: bad argument");
  *n_ids_out = 0;
  if (n == 0) return 0;
  // text_offsets == NULL: the rows are separated by single NUL bytes (n - 1 separators; the caller guarantees that no
  // row contains a NUL itself - the host joins the strings in one call instead of encoding them one by one)
  std::vector<int64_t> own_offsets;
  if (!text_offsets) {
    own_offsets.reserve(static_cast<size_t>(n) + 1);
    own_offsets.push_back(0);
    const char* p = text_bytes;
    for (int64_t r = 0; r + 1 < n; ++r) {
      const char* q = p + strlen(p);
      own_offsets.push_back(q - text_bytes);   // end of row r == position of its separator
      p = q + 1;
    }
    own_offsets.push_back((p - text_bytes) + static_cast<int64_t>(strlen(p)));
    // rows r >= 1 start one byte after the previous row's end: keep one offset array by treating the separator as
    // the first byte of the next row - NUL is a dropped character for the tokenizer, so the result is unchanged
    text_offsets = own_offsets.data();
  }
  int threads = n_threads > 0 ? n_threads : static_cast<int>(std::thread::hardware_concurrency());
  threads = std::max(1, std::min(threads, 32));
  if (n < 4096) threads = 1;
  threads = static_cast<int>(std::min<int64_t>(threads, n));
  struct Part {
    std::vector<int32_t> ids;
  };
  std::vector<Part> parts(static_cast<size_t>(threads));
  auto work = [&](int w) {
    const int64_t lo = n * w / threads, hi = n * (w + 1) / threads;
    std::vector<int32_t>& ids = parts[static_cast<size_t>(w)].ids;
    ids.reserve(static_cast<size_t>((text_offsets[hi] - text_offsets[lo]) / 3 + 2 * (hi - lo)));
    std::vector<int32_t> row;
    std::string word;
    for (int64_t r = lo; r < hi; ++r) {
      const char* s = text_bytes + text_offsets[r];
      const size_t len = static_cast<size_t>(text_offsets[r + 1] - text_offsets[r]);
      if (t->row(s, len, row, word)) {
        lens_out[r] = static_cast<int32_t>(row.size());
        ids.insert(ids.end(), row.begin(), row.end());
      } else {
        lens_out[r] = -1;
      }
    }
  };
  if (threads == 1) {
    work(0);
  } else {
    std::vector<std::thread> pool;
    pool.reserve(static_cast<size_t>(threads));
    for (int w = 0; w < threads; ++w) pool.emplace_back(work, w);
    for (std::thread& th : pool) th.join();
  }
  int64_t total = 0;
  for (const Part& p : parts) total += static_cast<int64_t>(p.ids.size());
  if (total > ids_capacity) return tfail("owc_tokenize_packed: ids_out too small");
  int64_t at = 0;
  for (const Part& p : parts) {
    if (!p.ids.empty()) memcpy(ids_out + at, p.ids.data(), p.ids.size() * sizeof(int32_t));
    at += static_cast<int64_t>(p.ids.size());
  }
  *n_ids_out = total;
  return 0;
}
