// Streaming reader / writer for the samples JSONL files of the scoring CLI.
//
// The reference parses every `*_samples_*.jsonl` file with `pd.read_json(..., lines=True)` and rewrites it with
// `df.to_json(..., lines=True, orient="records")` once per metric (`eval_metrics.py:62`, `:123`; the files are written
// by `src/engine/_tracker.py:314-341`, one `json.dumps` object per line). Only two members of each row feed the
// similarity metrics - `target` and `filtered_resps` - and only a handful of numeric members come back. This file
// does exactly that much work: one pass over the bytes that finds the top-level members of each line, decodes the
// one string the metric uses from each of the two members, and a second pass that copies every line through
// untouched while replacing / appending the score members.
//
// Host code only (no device work); part of libowc_b200.so so the Python side reaches it through the same C ABI.

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/owc_b200.h"

namespace {

thread_local std::string g_jsonl_err;

struct Fail {
    std::string msg;
};

[[noreturn]] void fail(size_t line, const char* what) {
    throw Fail{"jsonl: line " + std::to_string(line + 1) + ": " + what};
}

inline bool is_ws(unsigned char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n'; }

inline const char* skip_ws(const char* p, const char* e) {
    while (p < e && is_ws((unsigned char)*p)) ++p;
    return p;
}

// p points at the opening quote; returns the position after the closing quote.
inline const char* skip_string(const char* p, const char* e, size_t line) {
    ++p;
    while (p < e) {
        const char* q = (const char*)memchr(p, '"', (size_t)(e - p));
        if (!q) break;
        // count the backslashes in front of the quote: an odd count escapes it
        const char* b = q;
        while (b > p && b[-1] == '\\') --b;
        if (((q - b) & 1) == 0) return q + 1;
        p = q + 1;
    }
    fail(line, "unterminated string");
}

// Skips one JSON value of any type (containers by bracket depth, strings skipped whole).
const char* skip_value(const char* p, const char* e, size_t line) {
    if (p >= e) fail(line, "value expected");
    if (*p == '"') return skip_string(p, e, line);
    if (*p == '{' || *p == '[') {
        int depth = 0;
        while (p < e) {
            const char c = *p;
            if (c == '"') {
                p = skip_string(p, e, line);
                continue;
            }
            if (c == '{' || c == '[') ++depth;
            else if (c == '}' || c == ']') {
                if (--depth == 0) return p + 1;
            }
            ++p;
        }
        fail(line, "unterminated container");
    }
    const char* s = p;  // number / true / false / null
    while (p < e && *p != ',' && *p != '}' && *p != ']' && !is_ws((unsigned char)*p)) ++p;
    if (p == s) fail(line, "value expected");
    return p;
}

inline int hex4(const char* p, const char* e, size_t line) {
    if (e - p < 4) fail(line, "truncated \\u escape");
    int v = 0;
    for (int i = 0; i < 4; ++i) {
        const char c = p[i];
        int d;
        if (c >= '0' && c <= '9') d = c - '0';
        else if (c >= 'a' && c <= 'f') d = c - 'a' + 10;
        else if (c >= 'A' && c <= 'F') d = c - 'A' + 10;
        else fail(line, "bad \\u escape");
        v = (v << 4) | d;
    }
    return v;
}

inline void put_utf8(std::string& out, uint32_t cp) {
    if (cp < 0x80) out.push_back((char)cp);
    else if (cp < 0x800) {
        out.push_back((char)(0xC0 | (cp >> 6)));
        out.push_back((char)(0x80 | (cp & 0x3F)));
    } else if (cp < 0x10000) {
        out.push_back((char)(0xE0 | (cp >> 12)));
        out.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
        out.push_back((char)(0x80 | (cp & 0x3F)));
    } else {
        out.push_back((char)(0xF0 | (cp >> 18)));
        out.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
        out.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
        out.push_back((char)(0x80 | (cp & 0x3F)));
    }
}

// Decodes the JSON string [p, q) (quotes included) and appends it to `out`. Returns false when the text cannot be
// represented in the NUL-separated UTF-8 blob the caller builds (embedded NUL, lone surrogate).
bool decode_string(const char* p, const char* q, std::string& out, size_t line) {
    ++p;
    --q;
    // A raw control byte inside a string is not JSON (pandas / json.loads raise on it), and a raw NUL would split the
    // row in the NUL-separated blob: such a file is handed to the pandas path, which reports it like the reference does.
    for (const char* c = p; c < q; ++c)
        if ((unsigned char)*c < 0x20) return false;
    while (p < q) {
        const char* b = (const char*)memchr(p, '\\', (size_t)(q - p));
        if (!b) {
            out.append(p, (size_t)(q - p));
            break;
        }
        out.append(p, (size_t)(b - p));
        if (b + 1 >= q) fail(line, "dangling backslash");
        const char c = b[1];
        p = b + 2;
        switch (c) {
            case '"': out.push_back('"'); break;
            case '\\': out.push_back('\\'); break;
            case '/': out.push_back('/'); break;
            case 'b': out.push_back('\b'); break;
            case 'f': out.push_back('\f'); break;
            case 'n': out.push_back('\n'); break;
            case 'r': out.push_back('\r'); break;
            case 't': out.push_back('\t'); break;
            case 'u': {
                uint32_t cp = (uint32_t)hex4(p, q, line);
                p += 4;
                if (cp >= 0xD800 && cp <= 0xDBFF) {
                    if (q - p >= 6 && p[0] == '\\' && p[1] == 'u') {
                        const uint32_t lo = (uint32_t)hex4(p + 2, q, line);
                        if (lo >= 0xDC00 && lo <= 0xDFFF) {
                            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                            p += 6;
                        } else return false;
                    } else return false;
                } else if (cp >= 0xDC00 && cp <= 0xDFFF) return false;
                if (cp == 0) return false;
                put_utf8(out, cp);
                break;
            }
            default: fail(line, "bad escape");
        }
    }
    return true;
}

struct Member {
    const char* lead_b;   // first byte after the '{' or ',' in front of the member (whitespace before the key)
    const char* key_b;    // opening quote of the key
    const char* val_b;    // first byte of the value
    const char* val_e;    // one past the value
    const char* sep_e;    // one past the ',' that follows the value, or val_e when it is the last member
};

// Walks the top-level members of one line `{...}`; calls f(Member, key bytes without quotes, key length) for each.
// Keys are compared as raw bytes: a key spelled with escapes never matches (json.dumps does not write such keys).
template <class F>
void walk_members(const char* p, const char* e, size_t line, F&& f) {
    p = skip_ws(p, e);
    if (p >= e || *p != '{') fail(line, "object expected");
    ++p;
    const char* lead = p;
    p = skip_ws(p, e);
    if (p < e && *p == '}') {
        if (skip_ws(p + 1, e) != e) fail(line, "trailing data after the object");
        return;
    }
    for (;;) {
        p = skip_ws(p, e);
        if (p >= e || *p != '"') fail(line, "member name expected");
        Member m;
        m.lead_b = lead;
        m.key_b = p;
        const char* key_e = skip_string(p, e, line);
        p = skip_ws(key_e, e);
        if (p >= e || *p != ':') fail(line, "':' expected");
        p = skip_ws(p + 1, e);
        m.val_b = p;
        m.val_e = skip_value(p, e, line);
        p = skip_ws(m.val_e, e);
        if (p >= e) fail(line, "unterminated object");
        const bool last = *p == '}';
        if (!last && *p != ',') fail(line, "',' or '}' expected");
        m.sep_e = last ? m.val_e : p + 1;
        f(m, m.key_b + 1, (size_t)(key_e - m.key_b - 2));
        if (last) {
            if (skip_ws(p + 1, e) != e) fail(line, "trailing data after the object");   // `{...} junk`: pandas raises too
            return;
        }
        ++p;
        lead = p;
    }
}

// Calls f(line_begin, line_end, index) for every non-blank line (pandas drops blank lines the same way).
template <class F>
void for_lines(const char* buf, size_t len, F&& f) {
    const char* p = buf;
    const char* e = buf + len;
    size_t n = 0;
    while (p < e) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(e - p));
        const char* le = nl ? nl : e;
        const char* a = skip_ws(p, le);
        if (a < le) {
            const char* b = le;
            while (b > a && is_ws((unsigned char)b[-1])) --b;
            f(a, b, n++);
        }
        p = nl ? nl + 1 : e;
    }
}

// First and last element of the JSON list [b, e); false when the value is not a list or the list is empty.
struct ListEnds {
    const char* first_b;
    const char* first_e;
    const char* last_b;
    const char* last_e;
};

bool list_ends(const char* b, const char* e, size_t line, ListEnds* out) {
    if (b >= e || *b != '[') return false;
    const char* p = skip_ws(b + 1, e);
    if (p < e && *p == ']') return false;
    out->first_b = nullptr;
    for (;;) {
        p = skip_ws(p, e);
        const char* vb = p;
        const char* ve = skip_value(p, e, line);
        if (!out->first_b) {
            out->first_b = vb;
            out->first_e = ve;
        }
        out->last_b = vb;
        out->last_e = ve;
        p = skip_ws(ve, e);
        if (p >= e) fail(line, "unterminated list");
        if (*p == ']') return true;
        if (*p != ',') fail(line, "',' or ']' expected");
        ++p;
    }
}

inline bool take_string(const char* b, const char* e, size_t line, const char** sb, const char** se) {
    if (b >= e || *b != '"') return false;
    *sb = b;
    *se = skip_string(b, e, line);
    return true;
}

// The reference string of a row (`ref[0] if isinstance(ref, list) else ref`, `_group.py:518`):
//   "s" -> s ; ["s0", ...] -> s0. False for any other shape (the caller then reports the file as unsupported).
bool pick_reference(const char* b, const char* e, size_t line, const char** sb, const char** se) {
    if (take_string(b, e, line, sb, se)) return true;
    ListEnds le;
    return list_ends(b, e, line, &le) && take_string(le.first_b, le.first_e, line, sb, se);
}

// `eval_metrics.py:66-68` decides ONCE per file, on the first row, whether the responses are multi-round
// (`predictions[0][0]` is a list) and then takes `prediction[0]` of every row.
bool is_multi_round(const char* b, const char* e, size_t line) {
    ListEnds le;
    return list_ends(b, e, line, &le) && *le.first_b == '[';
}

// The scored prediction of a row (`pred[-1] if isinstance(pred, list) else pred`, `_group.py:519`, applied to the
// row's value, or to its first element in a multi-round file): the result must be a string.
bool pick_prediction(const char* b, const char* e, bool multi_round, size_t line, const char** sb, const char** se) {
    ListEnds le;
    if (multi_round) {                                          // pred = value[0]
        if (!list_ends(b, e, line, &le)) return false;          // (a bare string would be indexed by character)
        b = le.first_b;
        e = le.first_e;
    }
    if (take_string(b, e, line, sb, se)) return true;
    return list_ends(b, e, line, &le) && take_string(le.last_b, le.last_e, line, sb, se);
}

// pandas' ujson writer: fixed notation, 10 decimals, trailing zeros trimmed to one, NaN / inf -> null.
inline void put_double(std::string& out, double v) {
    if (!std::isfinite(v)) {
        out += "null";
        return;
    }
    if (v == 0.0) {                                             // -0.0 is written without the sign ...
        out += "0.0";
        return;
    }
    char tmp[64];
    int n = snprintf(tmp, sizeof tmp, "%.10f", v);               // ... but -1e-12 is written as -0.0
    while (n > 2 && tmp[n - 1] == '0' && tmp[n - 2] != '.') --n;
    out.append(tmp, (size_t)n);
}

// Copies every row with its other members byte for byte, drops members named in names[] and appends
// names[i]:values[i * n_rows + row]; hands the text to `sink` whenever `flush_at` bytes have accumulated and at the end.
// OWC_JSONL_RAW columns take their per-row text (a JSON fragment, written verbatim) from raws[i]: n_rows fragments, each
// followed by one NUL byte; values[] is ignored for them.
template <class Sink>
void render_columns(const char* buf, size_t len, const char* const* names, const int32_t* kinds, const double* values,
                    const char* const* raws, int n_cols, size_t n_rows, size_t flush_at, Sink&& sink) {
    std::vector<size_t> nl((size_t)n_cols);
    for (int i = 0; i < n_cols; ++i) nl[(size_t)i] = strlen(names[i]);
    std::vector<const char*> raw_pos((size_t)n_cols, nullptr);
    for (int i = 0; i < n_cols; ++i) {
        if (kinds[i] == OWC_JSONL_RAW) {
            if (!raws || !raws[i]) throw Fail{"jsonl: raw column without text"};
            raw_pos[(size_t)i] = raws[i];
        }
    }
    std::string dst;
    dst.reserve(flush_at == ~size_t(0) ? len + n_rows * 32 * (size_t)(n_cols + 1) : flush_at + (size_t(1) << 16));
    size_t rows = 0;
    for_lines(buf, len, [&](const char* a, const char* b, size_t line) {
        if (rows >= n_rows) fail(line, "more rows than values");
        dst.push_back('{');
        bool any = false;
        // kept members are copied with the source's own spacing and separators, as few runs as possible: a run ends
        // where a member is dropped (a row scored for the first time is one memcpy)
        const char* run_b = nullptr;
        const char* run_e = nullptr;
        walk_members(a, b, line, [&](const Member& m, const char* k, size_t kl) {
            for (int i = 0; i < n_cols; ++i)
                if (kl == nl[(size_t)i] && memcmp(k, names[i], kl) == 0) {
                    if (run_b) dst.append(run_b, (size_t)(run_e - run_b));
                    run_b = nullptr;
                    return;
                }
            if (!run_b) {
                if (any) dst.push_back(',');
                run_b = m.lead_b;
            }
            run_e = m.val_e;
            any = true;
        });
        if (run_b) dst.append(run_b, (size_t)(run_e - run_b));
        for (int i = 0; i < n_cols; ++i) {
            if (any) dst.push_back(',');
            dst.push_back('"');
            dst.append(names[i], nl[(size_t)i]);
            dst += "\":";
            if (kinds[i] == OWC_JSONL_RAW) {
                const char* frag = raw_pos[(size_t)i];
                const size_t fl = strlen(frag);
                dst.append(frag, fl);
                raw_pos[(size_t)i] = frag + fl + 1;
            } else {
                const double v = values[(size_t)i * n_rows + rows];
                if (kinds[i] == OWC_JSONL_INT && std::isfinite(v)) dst += std::to_string((long long)v);
                else put_double(dst, v);
            }
            any = true;
        }
        dst += "}\n";
        ++rows;
        if (dst.size() >= flush_at) {
            sink(dst);
            dst.clear();
        }
    });
    if (rows != n_rows) throw Fail{"jsonl: " + std::to_string(rows) + " rows in the file, " + std::to_string(n_rows) +
                                   " values per column"};
    if (!dst.empty()) sink(dst);
}

template <class F>
int guarded(F&& f) {
    try {
        return f();
    } catch (const Fail& x) {
        g_jsonl_err = x.msg;
    } catch (const std::exception& x) {
        g_jsonl_err = std::string("jsonl: ") + x.what();
    }
    return OWC_JSONL_ERROR;
}

}  // namespace

extern "C" {

const char* owc_jsonl_last_error(void) { return g_jsonl_err.c_str(); }

int owc_jsonl_count_rows(const char* buf, size_t len, size_t* n_rows) {
    return guarded([&] {
        size_t n = 0;
        for_lines(buf, len, [&](const char*, const char*, size_t) { ++n; });
        *n_rows = n;
        return OWC_JSONL_OK;
    });
}

int owc_jsonl_extract(const char* buf, size_t len, const char* key, int mode, char* out, size_t out_cap,
                      size_t* out_len, size_t* n_rows) {
    return guarded([&] {
        if (mode != OWC_JSONL_FIRST && mode != OWC_JSONL_LAST) throw Fail{"jsonl: unknown extract mode"};
        const size_t klen = strlen(key);
        std::string blob;
        blob.reserve(len / 8 + 64);
        bool fallback = false;
        bool multi_round = false;
        size_t rows = 0;
        for_lines(buf, len, [&](const char* a, const char* b, size_t line) {
            if (fallback) return;
            const char* vb = nullptr;
            const char* ve = nullptr;
            walk_members(a, b, line, [&](const Member& m, const char* k, size_t kl) {
                if (kl == klen && memcmp(k, key, klen) == 0) {  // duplicates: the last one wins, like json.loads
                    vb = m.val_b;
                    ve = m.val_e;
                }
            });
            const char* sb;
            const char* se;
            if (vb && mode == OWC_JSONL_LAST && rows == 0) multi_round = is_multi_round(vb, ve, line);
            const bool ok = vb && (mode == OWC_JSONL_FIRST ? pick_reference(vb, ve, line, &sb, &se)
                                                           : pick_prediction(vb, ve, multi_round, line, &sb, &se));
            if (!ok || !decode_string(sb, se, blob, line)) {
                fallback = true;
                return;
            }
            blob.push_back('\0');
            ++rows;
        });
        if (fallback) {
            g_jsonl_err = std::string("jsonl: member '") + key + "' is missing or has a shape the fast path does not cover";
            return OWC_JSONL_UNSUPPORTED;
        }
        if (blob.size() > out_cap) throw Fail{"jsonl: output buffer too small"};
        memcpy(out, blob.data(), blob.size());
        *out_len = blob.size();
        *n_rows = rows;
        return OWC_JSONL_OK;
    });
}

int owc_jsonl_extract_pairs(const char* buf, size_t len, const char* ref_key, const char* pred_key, char* out_refs,
                            size_t refs_cap, size_t* refs_len, char* out_preds, size_t preds_cap, size_t* preds_len,
                            size_t* n_rows) {
    return guarded([&] {
        const size_t rl = strlen(ref_key), pl = strlen(pred_key);
        std::string refs, preds;
        refs.reserve(len / 16 + 64);
        preds.reserve(len / 8 + 64);
        bool fallback = false;
        bool multi_round = false;
        size_t rows = 0;
        for_lines(buf, len, [&](const char* a, const char* b, size_t line) {
            if (fallback) return;
            const char *rb = nullptr, *re = nullptr, *pb = nullptr, *pe = nullptr;
            walk_members(a, b, line, [&](const Member& m, const char* k, size_t kl) {
                if (kl == rl && memcmp(k, ref_key, rl) == 0) {         // duplicates: the last one wins, like json.loads
                    rb = m.val_b;
                    re = m.val_e;
                } else if (kl == pl && memcmp(k, pred_key, pl) == 0) {
                    pb = m.val_b;
                    pe = m.val_e;
                }
            });
            const char* sb;
            const char* se;
            if (pb && rows == 0) multi_round = is_multi_round(pb, pe, line);
            if (!rb || !pb || !pick_reference(rb, re, line, &sb, &se) || !decode_string(sb, se, refs, line) ||
                !pick_prediction(pb, pe, multi_round, line, &sb, &se) || !decode_string(sb, se, preds, line)) {
                fallback = true;
                return;
            }
            refs.push_back('\0');
            preds.push_back('\0');
            ++rows;
        });
        if (fallback) {
            g_jsonl_err = std::string("jsonl: member '") + ref_key + "' or '" + pred_key +
                          "' is missing or has a shape the fast path does not cover";
            return OWC_JSONL_UNSUPPORTED;
        }
        if (refs.size() > refs_cap || preds.size() > preds_cap) throw Fail{"jsonl: output buffer too small"};
        memcpy(out_refs, refs.data(), refs.size());
        memcpy(out_preds, preds.data(), preds.size());
        *refs_len = refs.size();
        *preds_len = preds.size();
        *n_rows = rows;
        return OWC_JSONL_OK;
    });
}

int owc_blob_unique(const char* blob, size_t len, size_t n, int32_t* inverse, char* out, size_t out_cap, size_t* out_len,
                    size_t* n_unique) {
    return guarded([&] {
        // rows: n NUL-terminated strings back to back (what owc_jsonl_extract* write); first-occurrence order is kept
        size_t cap = 16;
        while (cap < 2 * n + 2) cap <<= 1;
        std::vector<int64_t> slot(cap, -1);               // offset of the unique string in `out`, -1 = empty
        std::vector<int32_t> slot_id(cap, 0);
        size_t used = 0, uniq = 0;
        const char* p = blob;
        const char* e = blob + len;
        for (size_t i = 0; i < n; ++i) {
            if (p >= e) throw Fail{"blob: fewer rows than announced"};
            const char* z = (const char*)memchr(p, 0, (size_t)(e - p));
            if (!z) throw Fail{"blob: missing terminator"};
            const size_t l = (size_t)(z - p);
            uint64_t h = 1469598103934665603ull;            // FNV-1a
            for (size_t j = 0; j < l; ++j) h = (h ^ (unsigned char)p[j]) * 1099511628211ull;
            size_t k = (size_t)(h ^ (h >> 29)) & (cap - 1);
            for (;;) {
                if (slot[k] < 0) {
                    if (used + l + 1 > out_cap) throw Fail{"blob: output buffer too small"};
                    memcpy(out + used, p, l + 1);
                    slot[k] = (int64_t)used;
                    slot_id[k] = (int32_t)uniq;
                    inverse[i] = (int32_t)uniq;
                    used += l + 1;
                    ++uniq;
                    break;
                }
                const char* q = out + slot[k];
                if (strlen(q) == l && memcmp(q, p, l) == 0) {
                    inverse[i] = slot_id[k];
                    break;
                }
                k = (k + 1) & (cap - 1);
            }
            p = z + 1;
        }
        *out_len = used;
        *n_unique = uniq;
        return OWC_JSONL_OK;
    });
}

int owc_jsonl_has_members(const char* buf, size_t len, const char* const* names, int n_names, int* all_rows) {
    return guarded([&] {
        std::vector<size_t> nl((size_t)n_names);
        for (int i = 0; i < n_names; ++i) nl[(size_t)i] = strlen(names[i]);
        bool all = true;
        size_t rows = 0;
        for_lines(buf, len, [&](const char* a, const char* b, size_t line) {
            if (!all) return;
            uint64_t seen = 0;
            walk_members(a, b, line, [&](const Member&, const char* k, size_t kl) {
                for (int i = 0; i < n_names && i < 64; ++i)
                    if (kl == nl[(size_t)i] && memcmp(k, names[i], kl) == 0) seen |= 1ull << i;
            });
            const uint64_t want = n_names >= 64 ? ~0ull : ((1ull << n_names) - 1);
            if (seen != want) all = false;
            ++rows;
        });
        *all_rows = (all && rows > 0) ? 1 : 0;
        return OWC_JSONL_OK;
    });
}

int owc_jsonl_set_columns(const char* buf, size_t len, const char* const* names, const int32_t* kinds,
                          const double* values, int n_cols, size_t n_rows, char* out, size_t out_cap,
                          size_t* out_len) {
    return guarded([&] {
        size_t used = 0;
        render_columns(buf, len, names, kinds, values, nullptr, n_cols, n_rows, ~size_t(0), [&](const std::string& chunk) {
            if (used + chunk.size() > out_cap) throw Fail{"jsonl: output buffer too small"};
            memcpy(out + used, chunk.data(), chunk.size());
            used += chunk.size();
        });
        *out_len = used;
        return OWC_JSONL_OK;
    });
}

static int write_columns_impl(const char* buf, size_t len, const char* const* names, const int32_t* kinds, const double* values,
                              const char* const* raws, int n_cols, size_t n_rows, const char* out_path) {
    return guarded([&] {
        FILE* f = fopen(out_path, "wb");
        if (!f) throw Fail{std::string("jsonl: cannot open ") + out_path + " for writing"};
        try {
            render_columns(buf, len, names, kinds, values, raws, n_cols, n_rows, size_t(4) << 20, [&](const std::string& chunk) {
                if (fwrite(chunk.data(), 1, chunk.size(), f) != chunk.size())
                    throw Fail{std::string("jsonl: short write to ") + out_path};
            });
        } catch (...) {
            fclose(f);
            remove(out_path);
            throw;
        }
        if (fclose(f) != 0) {
            remove(out_path);
            throw Fail{std::string("jsonl: cannot finish ") + out_path};
        }
        return OWC_JSONL_OK;
    });
}

int owc_jsonl_write_columns(const char* buf, size_t len, const char* const* names, const int32_t* kinds,
                            const double* values, int n_cols, size_t n_rows, const char* out_path) {
    return write_columns_impl(buf, len, names, kinds, values, nullptr, n_cols, n_rows, out_path);
}

int owc_jsonl_write_columns_raw(const char* buf, size_t len, const char* const* names, const int32_t* kinds,
                                const double* values, const char* const* raws, int n_cols, size_t n_rows,
                                const char* out_path) {
    return write_columns_impl(buf, len, names, kinds, values, raws, n_cols, n_rows, out_path);
}

int owc_jsonl_extract_raw(const char* buf, size_t len, const char* key, char* out, size_t out_cap, size_t* out_len,
                          size_t* n_rows) {
    return guarded([&] {
        const size_t klen = strlen(key);
        std::string blob;
        blob.reserve(len / 4 + 64);
        bool missing = false;
        size_t rows = 0;
        for_lines(buf, len, [&](const char* a, const char* b, size_t line) {
            if (missing) return;
            const char* vb = nullptr;
            const char* ve = nullptr;
            walk_members(a, b, line, [&](const Member& m, const char* k, size_t kl) {
                if (kl == klen && memcmp(k, key, klen) == 0) {
                    vb = m.val_b;
                    ve = m.val_e;
                }
            });
            if (!vb || memchr(vb, 0, (size_t)(ve - vb)) != nullptr) {
                missing = true;
                return;
            }
            blob.append(vb, (size_t)(ve - vb));
            blob.push_back('\0');
            ++rows;
        });
        if (missing) {
            g_jsonl_err = std::string("jsonl: member '") + key + "' is missing in a row";
            return OWC_JSONL_UNSUPPORTED;
        }
        if (blob.size() > out_cap) throw Fail{"jsonl: output buffer too small"};
        memcpy(out, blob.data(), blob.size());
        *out_len = blob.size();
        *n_rows = rows;
        return OWC_JSONL_OK;
    });
}

}  // extern "C"
