/*
 * owc_b200.h - C ABI of the B200-native text-similarity scoring path of altndrr/lmms-owc.
 *
 * The reference (pure Python, no FFI of its own - SURVEY.md F9) reaches this path through
 *   src/data/pipelines/text/_text.py:143-208   encode_sentence_bert(batch, rank, **kwargs)
 *   src/data/metrics/_group.py:488-544         semantic_similarity      (paired dot, :537-544)
 *   src/data/metrics/_group.py:392-458         mean_average_semantic_similarity
 *   src/data/metrics/_group.py:176-334         concept_semantic_similarity (encode + paired dot, :287-303)
 * so every entry point below cites the reference lines it replaces. The Python host side
 * (lmms-owc_b200/*.py) keeps the reference's operator/plugin signatures and binds these symbols
 * with ctypes; INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C: pointers, sizes, opaque handles; no torch / C++ types cross this boundary;
 *   - every function returns 0 on success, non-zero on failure; owc_last_error() (thread-local)
 *     describes the last failure;
 *   - one owc_ctx per (process, device); calls on a ctx are serialised by the caller; all device
 *     work is enqueued on the ctx's CUDA stream; functions that return HOST results synchronise
 *     that stream before returning, functions that leave results in DEVICE memory do not;
 *   - `space` arguments say where a caller buffer lives: OWC_MEM_HOST (pageable or pinned host
 *     memory - pinned makes the copies asynchronous) or OWC_MEM_DEVICE (memory of the ctx's device);
 *   - there is no CPU fallback: without an sm_100 device owc_ctx_create fails.
 */
#ifndef OWC_B200_H
#define OWC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OWC_ABI_VERSION 3

typedef struct owc_ctx owc_ctx;
typedef struct owc_encoder owc_encoder;

enum { OWC_ARCH_BERT = 0, OWC_ARCH_MPNET = 1 };

/* OWC_PREC_BF16: bf16 tensor-core operands (tcgen05), fp32 accumulation, fp32 residual stream,
 *                LayerNorm / softmax / pooling statistics in fp32 - the production mode.
 * OWC_PREC_FP32: "check mode" - every contraction in fp32 on the SIMT pipes (north star: cosine
 *                within 1e-4 of the reference's fp32 CPU path). Slow; for parity work only.      */
enum { OWC_PREC_BF16 = 0, OWC_PREC_FP32 = 1 };

enum { OWC_MEM_HOST = 0, OWC_MEM_DEVICE = 1 };

/* Architecture of the sentence encoder the reference loads at _text.py:158-163 (a BertModel for
 * all-MiniLM-L6-v2) or the MPNet variant BASELINE.json config 3 names.                           */
typedef struct owc_arch {
  int32_t kind;        /* OWC_ARCH_BERT | OWC_ARCH_MPNET */
  int32_t layers;
  int32_t hidden;      /* H; multiple of 64, <= 1024 */
  int32_t heads;       /* head_dim = hidden / heads must be 32 or 64 */
  int32_t inter;       /* FFN width; multiple of 64 */
  int32_t vocab;
  int32_t max_pos;     /* rows of the position table */
  int32_t type_vocab;  /* BERT: rows of the token-type table (row 0 is used); MPNet: 0 */
  int32_t pad_id;      /* MPNet position ids skip this id (modeling_mpnet.py:889-897) */
  float ln_eps;
} owc_arch;

/* Work counters of a ctx (bench.py reports gpu_launches from these). */
typedef struct owc_stats {
  int64_t kernel_launches;   /* kernels of this library launched since ctx creation */
  int64_t h2d_bytes;         /* bytes copied host->device by this library */
  int64_t d2h_bytes;         /* bytes copied device->host by this library */
  int64_t tokens_encoded;
  int64_t sequences_encoded;
} owc_stats;

/* Kernel categories for the event-bracketed profile (bench.py's roofline leg). */
enum {
  OWC_KCAT_EMBED = 0,    /* k_embed_ln  */
  OWC_KCAT_GEMM = 1,     /* k_gemm_tc (encoder contractions) / k_gemm_f32 in check mode */
  OWC_KCAT_ATTN = 2,     /* k_attn      */
  OWC_KCAT_LN = 3,       /* k_ln        */
  OWC_KCAT_POOL = 4,     /* k_pool_l2   */
  OWC_KCAT_SIM = 5,      /* k_gemm_tc with the top-k epilogue */
  OWC_KCAT_RESCORE = 6,  /* k_rescore / k_topk_rows / k_pair_dot */
  OWC_KCAT_MISC = 7,     /* conversions */
  /* per-shape breakdown of the wide encoder's contractions (k_gemm_wide); each launch is ALSO counted in OWC_KCAT_GEMM */
  OWC_KCAT_GEMM_QKV = 8,
  OWC_KCAT_GEMM_O = 9,
  OWC_KCAT_GEMM_FFN1 = 10,
  OWC_KCAT_GEMM_FFN2 = 11,
  OWC_KCAT_COUNT = 12
};
typedef struct owc_profile {
  double ms[OWC_KCAT_COUNT];         /* summed device time of the launches, CUDA events on the ctx stream */
  int64_t launches[OWC_KCAT_COUNT];
} owc_profile;

int32_t owc_abi_version(void);
const char* owc_last_error(void);

/* ctx: device ordinal + CUDA stream. `stream` is a cudaStream_t passed as void* (e.g.
 * torch.cuda.Stream().cuda_stream); NULL makes the library create and own a stream.
 * `chunk_tokens` caps the tokens encoded per pass (activation workspace size); 0 = default.      */
int32_t owc_ctx_create(int32_t device, void* stream, int32_t chunk_tokens, owc_ctx** out);
int32_t owc_ctx_destroy(owc_ctx* ctx);
int32_t owc_ctx_sync(owc_ctx* ctx);
int32_t owc_ctx_stats(owc_ctx* ctx, owc_stats* out);
/* Between begin and end every kernel launch of this ctx is bracketed by two CUDA events on the ctx
 * stream; end synchronises and returns per-category sums. Adds launch overhead: use it for the
 * roofline leg, never around a throughput measurement.                                           */
int32_t owc_ctx_profile_begin(owc_ctx* ctx);
int32_t owc_ctx_profile_end(owc_ctx* ctx, owc_profile* out);

/* Replaces AutoModel.from_pretrained + .to(device, dtype) (_text.py:158-170). `weights` is the
 * host fp32 blob laid out as documented in lmms-owc_b200/arch.py (pack_weights); n_floats must
 * match the architecture exactly. The library keeps fp32 masters and bf16 operand copies.        */
int32_t owc_encoder_load(owc_ctx* ctx, const owc_arch* arch, const float* weights, size_t n_floats,
                         owc_encoder** out);
int32_t owc_encoder_destroy(owc_encoder* enc);

/* Replaces _text.py:196-202: BertModel/MPNetModel forward, masked mean pooling (:175-189) and L2
 * normalisation (:202) for n_seq already-tokenised sequences.
 *   ids        int32 [cu_seqlens[n_seq]]  token ids, sequences back to back (no padding)
 *   cu_seqlens int32 [n_seq + 1]          HOST memory always; cu_seqlens[0] == 0, non-decreasing
 *   out        fp32  [n_seq, hidden]      unit-norm rows (NaN for an empty sequence, like :202)
 * Sequences longer than the position table are rejected (the host tokenizer truncates, :193-195). */
int32_t owc_encode(owc_encoder* enc, const int32_t* ids, int32_t ids_space, const int32_t* cu_seqlens,
                   int32_t n_seq, int32_t precision, float* out, int32_t out_space);

/* Replaces torch.bmm(refs_z.unsqueeze(1), preds_z.unsqueeze(2)) (_group.py:537-544, :301-303,
 * :441-456; eval_ranking.py:219-220): out[i] = sum_d A[i,d] * B[i,d], fp32, where row i of A is
 * a[a_index[i]] (or a[i] if a_index is NULL; then na == n) and likewise for B. The index form lets
 * the host encode each distinct string once (the reference re-encodes the class name per row, F5).
 *   a fp32 [na, dim], b fp32 [nb, dim], a_index/b_index int32 [n] or NULL, out fp32 [n];
 *   every buffer lives in `space`.                                                               */
int32_t owc_pair_dot(owc_ctx* ctx, const float* a, int64_t na, const int32_t* a_index, const float* b,
                     int64_t nb, const int32_t* b_index, int64_t n, int32_t dim, int32_t space, float* out);

/* North-star extension (no reference symbol): S = Zp Zc^T, row-wise
 * top-k, S itself never written to memory. OWC_PREC_BF16: bf16 tcgen05 GEMM whose epilogue keeps
 * per-row candidates in registers, then exact fp32 re-scoring of the candidates; OWC_PREC_FP32:
 * fp32 throughout. Ties go to the lower class index. Requires 1 <= k <= 16 and k <= c.
 *   zp fp32 [n, dim], zc fp32 [c, dim], scores fp32 [n, k] (descending), idx int32 [n, k]        */
int32_t owc_sim_topk(owc_ctx* ctx, const float* zp, int64_t n, const float* zc, int32_t c, int32_t dim,
                     int32_t k, int32_t precision, int32_t space, float* scores, int32_t* idx);

/* The whole hot path for one shard of prediction rows in one call (what bench.py times):
 * encode the predictions, match them against class embeddings zc (DEVICE memory, from a previous
 * owc_encode with out_space = OWC_MEM_DEVICE), return top-k and - if targets != NULL - the paired
 * score S[i, targets[i]] that semantic_similarity reports (_group.py:544).
 *   ids/ids_space/cu_seqlens/n_seq as in owc_encode; targets int32 [n_seq] in ids_space or NULL;
 *   out_scores fp32 [n_seq,k], out_idx int32 [n_seq,k], out_paired fp32 [n_seq] or NULL, all in
 *   out_space; out_embeds fp32 [n_seq, hidden] in out_space or NULL.                             */
int32_t owc_score(owc_encoder* enc, const int32_t* ids, int32_t ids_space, const int32_t* cu_seqlens,
                  int32_t n_seq, const float* zc, int32_t c, const int32_t* targets, int32_t k,
                  int32_t precision, float* out_scores, int32_t* out_idx, float* out_paired,
                  float* out_embeds, int32_t out_space);

/* The matching half of owc_score on its own, for prediction embeddings that are already in DEVICE memory (owc_encode with
 * out_space = OWC_MEM_DEVICE): top-k of zp zc^T plus the paired score S[i, targets[i]] (_group.py:544). Splitting the
 * call lets a multi-GPU host encode its prediction shard while the class embeddings are still in flight (broadcast /
 * all-gather over NCCL) and wait for them only here.
 *   zp fp32 [n, dim] DEVICE, zc fp32 [c, dim] DEVICE; targets int32 [n] in targets_space or NULL; outputs in out_space.
 * A target outside [0, c) fails the call (at its synchronisation for device outputs) instead of reading out of bounds. */
int32_t owc_match(owc_ctx* ctx, const float* zp, int64_t n, const float* zc, int32_t c, int32_t dim,
                  const int32_t* targets, int32_t targets_space, int32_t k, int32_t precision, float* out_scores,
                  int32_t* out_idx, float* out_paired, int32_t out_space);

/* ---- host-side tokenisation (tokenisation stays on the host) --------------------
 * Replaces the AutoTokenizer call of _text.py:193-195 for ASCII rows: BERT normaliser + pre-tokeniser + WordPiece
 * + "[CLS] A [SEP]" (MPNet: "<s> A </s>") + truncation, multi-threaded, written straight into the packed layout
 * owc_encode takes. Rows with a non-ASCII byte or a literal special token are FLAGGED (length -1) for the host
 * library instead of being approximated. No CUDA involved.
 *   vocab_bytes/vocab_offsets: n_vocab UTF-8 strings back to back + int64 offsets [n_vocab+1]; token i has id i.  */
typedef struct owc_tokenizer owc_tokenizer;
const char* owc_tokenizer_last_error(void);
int32_t owc_tokenizer_create(const char* vocab_bytes, const int64_t* vocab_offsets, int32_t n_vocab, int32_t kind,
                             int32_t max_length, owc_tokenizer** out);
int32_t owc_tokenizer_destroy(owc_tokenizer* tok);
/*   text_bytes/text_offsets: n rows as one byte blob + int64 offsets [n+1]; ids_out: int32 [ids_capacity]
 *   (total bytes + 2n always suffices), rows back to back in row order; lens_out: int32 [n], tokens per row or -1
 *   for a flagged row (which contributes nothing to ids_out); n_ids_out: ids written; n_threads <= 0: all cores.  */
int32_t owc_tokenize_packed(const owc_tokenizer* tok, const char* text_bytes, const int64_t* text_offsets, int64_t n,
                            int32_t n_threads, int32_t* ids_out, int64_t ids_capacity, int32_t* lens_out,
                            int64_t* n_ids_out);

/* ---- samples JSONL reader / writer (host only) -------------------------------------
 * Replaces, for the two paired-similarity metrics, the `pd.read_json(input_file, lines=True)` of eval_metrics.py:62
 * and the `df.to_json(input_file, lines=True, orient="records")` of eval_metrics.py:123 (files written one
 * json.dumps object per line by src/engine/_tracker.py:314-341). Blank lines are skipped like pandas skips them.
 * Return codes: OWC_JSONL_OK, OWC_JSONL_ERROR (malformed input or bad arguments; owc_jsonl_last_error() names the
 * line), OWC_JSONL_UNSUPPORTED (valid JSON whose rows the fast path does not cover: use the pandas path).          */
#define OWC_JSONL_OK 0
#define OWC_JSONL_ERROR 1
#define OWC_JSONL_UNSUPPORTED 2
#define OWC_JSONL_FIRST 0 /* "s" -> s, ["s0", ...] -> s0: the reference string, `ref[0]` of _group.py:518          */
#define OWC_JSONL_LAST 1  /* the scored prediction: `pred[-1]` of _group.py:519 ("s" -> s, [..., "sn"] -> sn) applied
                             to the row's value, or to its FIRST element when the file is multi-round, which
                             eval_metrics.py:66-68 decides once per file on the first row ([[...], ...])            */
#define OWC_JSONL_FLOAT 0 /* written as pandas writes doubles: "%.10f", trailing zeros trimmed, NaN/inf -> null    */
#define OWC_JSONL_INT 1
#define OWC_JSONL_RAW 2   /* a JSON fragment per row, written verbatim (list-valued columns of the concept metric)       */
const char* owc_jsonl_last_error(void);
int32_t owc_jsonl_count_rows(const char* buf, size_t len, size_t* n_rows);
/* One string per row from top-level member `key`, JSON escapes decoded to UTF-8, each followed by one NUL byte, into
 * out (out_cap >= len always suffices). Rows are in file order; *n_rows = rows written.                           */
int32_t owc_jsonl_extract(const char* buf, size_t len, const char* key, int32_t mode, char* out, size_t out_cap,
                          size_t* out_len, size_t* n_rows);
/* Both strings of every row in ONE pass: ref_key with OWC_JSONL_FIRST into out_refs, pred_key with OWC_JSONL_LAST into
 * out_preds (each cap >= len always suffices); what the CLI uses with ("target", "filtered_resps").              */
int32_t owc_jsonl_extract_pairs(const char* buf, size_t len, const char* ref_key, const char* pred_key, char* out_refs,
                                size_t refs_cap, size_t* refs_len, char* out_preds, size_t preds_cap,
                                size_t* preds_len, size_t* n_rows);
/* Distinct strings of a blob of n NUL-terminated rows (the format the two extractors write), in first-occurrence order,
 * into out (cap >= len suffices) + inverse int32 [n] (row -> index of its string): the dedup of text.encode_unique
 * without a Python object per row.                                                                              */
int32_t owc_blob_unique(const char* blob, size_t len, size_t n, int32_t* inverse, char* out, size_t out_cap,
                        size_t* out_len, size_t* n_unique);
/* *all_rows = 1 iff the buffer has at least one row and every row carries every member of names[].               */
int32_t owc_jsonl_has_members(const char* buf, size_t len, const char* const* names, int32_t n_names,
                              int32_t* all_rows);
/* Copies every row to out with its other members byte for byte, drops any existing member named in names[] and
 * appends names[i]:values[i*n_rows + row] (kinds[i] = OWC_JSONL_FLOAT | OWC_JSONL_INT) in order; one "\n" per row. */
int32_t owc_jsonl_set_columns(const char* buf, size_t len, const char* const* names, const int32_t* kinds,
                              const double* values, int32_t n_cols, size_t n_rows, char* out, size_t out_cap,
                              size_t* out_len);
/* The same text written straight to the file out_path (created / truncated; removed again on failure): the caller
 * renames it over the samples file, so a crash never leaves a truncated one.                                      */
int32_t owc_jsonl_write_columns(const char* buf, size_t len, const char* const* names, const int32_t* kinds,
                                const double* values, int32_t n_cols, size_t n_rows, const char* out_path);
/* The same with OWC_JSONL_RAW columns: raws[i] (NULL for numeric columns) holds n_rows JSON fragments, each followed by
 * one NUL byte, written verbatim as the value of names[i] - the list-valued `last_resp_concepts` /
 * `last_resp_concepts_similarities` columns the reference writes for concept_semantic_similarity (eval_metrics.py:93-104). */
int32_t owc_jsonl_write_columns_raw(const char* buf, size_t len, const char* const* names, const int32_t* kinds,
                                    const double* values, const char* const* raws, int32_t n_cols, size_t n_rows,
                                    const char* out_path);
/* The raw JSON text of top-level member `key` of every row, each followed by one NUL byte, into out (cap >= len + rows
 * suffices): how the CLI reads a list-valued column such as pre-extracted `prediction_concepts`.
 * OWC_JSONL_UNSUPPORTED when a row lacks the member.                                                                  */
int32_t owc_jsonl_extract_raw(const char* buf, size_t len, const char* key, char* out, size_t out_cap, size_t* out_len,
                              size_t* n_rows);

#ifdef __cplusplus
}
#endif
#endif /* OWC_B200_H */
