/*
 * orag.h -- C ABI of the B200-native hybrid-retrieval hot path.
 *
 * The reference (gabrielcheda/optimized-rag) is pure Python and has no FFI of
 * its own: the drop-in boundary is the duck-typed Python surface
 * (DocumentStore.search, HybridRetriever.retrieve/hybrid_search,
 * ReciprocalRankFusion.fuse -- optimized_rag_b200/*.py mirror those), and this
 * header is the plain-C layer directly below it that a maintainer binds with
 * ctypes/cffi.  Each entry point names the reference
 * arithmetic it replaces (file:line under the reference tree).
 *
 * Conventions
 *   - every function returns 0 on success, a negative ORAG_E* code on failure;
 *     orag_last_error() returns a thread-local message for the last failure.
 *   - all pointers named d_* are DEVICE pointers (sm_100a, current device);
 *     the library never allocates device memory (sole exception: orag_exchange_alloc):
 *     callers pass every buffer, including a workspace whose size they query first.
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work
 *     (no hidden synchronisation) unless stated otherwise.
 *   - ids are int64 global chunk ids = row_id_base + local row; missing
 *     entries (fewer than k results) are -1 with score 0.
 *   - ordering everywhere: score descending, ties -> ascending chunk id
 *     (Python's stable sorted(reverse=True) over index-ordered input,
 *     rag/retrieval.py:320), except RRF, whose tie rule is the reference's
 *     dict-insertion order (rag/reranker.py:250-261).
 *   - concurrency: entry points may be called from any thread, but calls that
 *     target the same device must not overlap in time (the library keeps one
 *     auxiliary stream, its fork/join events and the profiling events per
 *     process; a workspace belongs to the call that was handed it until the
 *     work enqueued by that call has finished).  One process per GPU, calls
 *     serialised by the host wrapper (optimized_rag_b200._ffi.GPU_LOCK), is
 *     the supported arrangement; different processes never share state.
 */
#ifndef ORAG_H
#define ORAG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORAG_OK 0
#define ORAG_EINVAL (-1)   /* bad argument */
#define ORAG_ECUDA (-2)    /* CUDA runtime / driver error */
#define ORAG_EWORKSPACE (-3) /* workspace too small */
#define ORAG_EUNSUPPORTED (-4)

/* cosine scan modes */
#define ORAG_COS_EXACT 0 /* fp64 CUDA-core scan of every row (anchor / fallback / small N) */
#define ORAG_COS_TF32 1  /* tcgen05 kind::tf32 first pass straight off the fp32 corpus + fp64 re-score */
#define ORAG_COS_BF16 2  /* tcgen05 kind::f16 (bf16) first pass over a bf16 shadow copy + fp64 re-score */
#define ORAG_COS_F16 3   /* tcgen05 kind::f16 (IEEE fp16) first pass over an fp16 shadow copy whose rows are scaled by
                            powers of two (orag_f32_to_f16_rows) + fp64 re-score: same bytes and tensor rate as bf16,
                            8x smaller rounding error -> 4-6x fewer first-pass candidates.  Preferred shadow mode. */

/* per-query status bits written by the *_topk entry points */
#define ORAG_STATUS_OK 0
#define ORAG_STATUS_OVERFLOW 1 /* candidate buffer overflowed: result for this query is NOT valid,
                                  caller must re-run the query with ORAG_COS_EXACT / dense BM25 */
#define ORAG_STATUS_EXCHANGE_TIMEOUT 2 /* sharded search: a peer's block did not arrive in time (orag_hybrid_wait) */
/* diagnostic detail, set together with ORAG_STATUS_OVERFLOW by the BM25 first pass: which of its buffers was too small */
#define ORAG_STATUS_WHERE_TILE 16        /* a tile sub-range marked more docs than the per-warp accumulator holds */
#define ORAG_STATUS_WHERE_CANDIDATES 32  /* more first-pass candidates than slots */
#define ORAG_STATUS_WHERE_SURVIVORS 64   /* more candidates above the final threshold than re-score slots (2048) */

int orag_version(void);
const char *orag_last_error(void);
/* sm count / compute capability of the current device */
int orag_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* Measurement hooks (bench.py): number of kernels this library has launched so far, and CUDA-event
 * brackets (recorded on the caller's stream) around the two dominant kernels of the most recent
 * calls -- the cosine main scan and the BM25 tile kernel.  orag_profile_read synchronises on those
 * events; a slot that has not run since orag_profile_enable reports -1. */
unsigned long long orag_launch_count(void);
int orag_profile_enable(int on);
int orag_profile_read(float *scan_ms, float *bm25_ms);
/* Every bracket of one slot (0 = cosine main scan, 1 = BM25 first pass) recorded since orag_profile_enable(1), oldest
 * first, at most `cap` (the library keeps the last 256): returns the number of durations written to ms[], or a
 * negative ORAG_E* code.  Synchronises on the events it reads. */
int orag_profile_read_all(int slot, float *ms, int cap);
/* Timeline of the hybrid step (diagnostic, scripts/timeline.py): after orag_timeline_enable(1) every tagged launch of
 * the path (tags: csrc/common.cuh TimelineTag -- query prep, seed scan, seed finalize, main scan, prefilter, re-score,
 * selection, BM25 prepare / first pass / finalize, RRF, push, wait, merge) records a start and an end event on its own
 * stream; orag_timeline_read synchronises the device and returns up to `cap` (at most 4096 are kept) records as
 * milliseconds since the enable call.  Both calls synchronise the device; disabled (the default) the marks cost nothing. */
int orag_timeline_enable(int on);
int orag_timeline_read(int *tags, float *begin_ms, float *end_ms, int cap);

/* ---------------------------------------------------------------------------
 * Synthetic inputs: bit-identical to optimized_rag_b200/synthetic.py
 * ------------------------------------------------------------------------- */
int orag_gen_embeddings(float *d_out, int64_t n_rows, int dim, int64_t row_start, uint64_t seed,
                        int dup_per_mille, void *stream);
int orag_gen_doc_lengths(int32_t *d_out, int64_t n_docs, int64_t doc_start, uint64_t seed, int lmin, int lmax,
                         void *stream);
int orag_gen_tokens(int32_t *d_out, const int64_t *d_doc_off, int64_t n_docs, int64_t doc_start, uint64_t seed,
                    const uint64_t *d_thresholds, int vocab, void *stream);

/* ---------------------------------------------------------------------------
 * Ingest-side helpers for the cosine scan
 * ------------------------------------------------------------------------- */
/* d_inv_norm[r] = 1/||corpus[r]|| as fp32 (0 for an all-zero row).  Used only by the
 * low-precision first pass; final scores never depend on it. */
int orag_row_inv_norms(const float *d_corpus, int64_t n_rows, int dim, float *d_inv_norm, void *stream);
/* d_row_sq[r] = sum(a*a for a in row r) in the reference's float64 arithmetic (sequential Neumaier sum,
 * rag/retrieval.py:366 under CPython >= 3.12): a per-row constant the final re-score would otherwise recompute. */
int orag_row_sq(const float *d_corpus, int64_t n_rows, int dim, double *d_row_sq, void *stream);
/* fp32 -> bf16 (round-to-nearest-even) shadow copy for ORAG_COS_BF16 */
int orag_f32_to_bf16(const float *d_src, void *d_dst_bf16, int64_t count, void *stream);
/* fp32 rows -> fp16 shadow rows for ORAG_COS_F16: row r is multiplied by s_r = 2^e (largest |x| lands in
 * [2^14, 2^15)) and rounded to nearest fp16; d_inv_norm_scaled[r] = 1 / (||row|| * s_r) is what
 * orag_cosine_topk takes as d_inv_norm in this mode (may be NULL); d_scale[r] = s_r (may be NULL). */
int orag_f32_to_f16_rows(const float *d_src, int64_t n_rows, int dim, void *d_dst_f16, float *d_inv_norm_scaled,
                         float *d_scale, void *stream);

/* ---------------------------------------------------------------------------
 * Cosine top-k.  Replaces the pgvector statement at rag/document_store.py:448-460
 * (`ORDER BY embedding <=> q LIMIT k`, exact instead of HNSW) and the Python loop
 * at rag/retrieval.py:252-256 + 362-371.  Scores are the reference's float64
 * arithmetic (Neumaier-compensated sums as CPython >= 3.12 `sum()` performs them,
 * sqrt, one multiply, one divide; 0.0 when either magnitude is 0).
 *
 *   d_corpus     fp32 [n_rows, dim] row-major (16-byte aligned, dim % 4 == 0;
 *                tensor-core modes additionally need dim % 32 == 0 (tf32) / % 64 (bf16))
 *   d_inv_norm   fp32 [n_rows] from orag_row_inv_norms (TF32, BF16) or orag_f32_to_f16_rows (F16); NULL for EXACT
 *   d_shadow     bf16 / fp16 [n_rows, dim] (ORAG_COS_BF16 / ORAG_COS_F16, else NULL)
 *   d_row_sq     fp64 [n_rows] from orag_row_sq, or NULL: the reference's sum(a*a) of every row; with the table the
 *                final re-score only runs the dot-product chain (same bits either way)
 *   d_queries    fp32 [n_queries, dim]
 *   d_out_ids    int64 [n_queries, k]; d_out_scores fp64 [n_queries, k]
 *   d_out_status int32 [n_queries] ORAG_STATUS_* (may be NULL)
 * ------------------------------------------------------------------------- */
size_t orag_cosine_workspace_bytes(int64_t n_rows, int dim, int n_queries, int k, int mode);
int orag_cosine_topk(const float *d_corpus, const float *d_inv_norm, const void *d_shadow, const double *d_row_sq,
                     int64_t n_rows, int dim, int64_t row_id_base, const float *d_queries, int n_queries, int k, int mode,
                     int64_t *d_out_ids, double *d_out_scores, int32_t *d_out_status, void *d_workspace,
                     size_t workspace_bytes, void *stream);

/* The same search in separate calls, so that a caller can keep all scans back to back on ONE stream while the query
 * preparation of the next batch and the latency-bound tails of earlier batches run on others: ORAG_PHASE_PREP enqueues
 * the query conversion and norms, ORAG_PHASE_SCAN the seed pass and the main scan (candidate lists stay in
 * d_workspace), ORAG_PHASE_FINISH the candidate re-scores and the selection into d_out_*.  Same arguments in every call
 * (the same d_workspace; d_out_status is cleared by the preparation phase); the caller orders each phase's stream
 * after the previous phase's, and issues all calls of a batch before the preparation phase of the next batch on the
 * same device.  At most 256 queries, tensor-core modes only.  phases = ORAG_PHASE_ALL is orag_cosine_topk. */
#define ORAG_PHASE_SCAN 1
#define ORAG_PHASE_FINISH 2
#define ORAG_PHASE_PREP 4
#define ORAG_PHASE_ALL 7
int orag_cosine_topk_phase(const float *d_corpus, const float *d_inv_norm, const void *d_shadow, const double *d_row_sq,
                           int64_t n_rows, int dim, int64_t row_id_base, const float *d_queries, int n_queries, int k,
                           int mode, int64_t *d_out_ids, double *d_out_scores, int32_t *d_out_status, void *d_workspace,
                           size_t workspace_bytes, int phases, void *stream);

/* Diagnostic: per-query sizes of the candidate sets of the LAST tensor-core search (<= 256 queries) that used
 * d_workspace -- first-pass candidates (may exceed the 4096 slots: overflow) and survivors of the fp32 re-score --
 * copied to HOST arrays of n_queries words; synchronises `stream`. */
int orag_cosine_last_counts(const void *d_workspace, int dim, int n_queries, uint32_t *h_candidates,
                            uint32_t *h_survivors, void *stream);

/* Co-scheduling hook: with orag_cosine_mark_prescan(1), every orag_cosine_topk in a tensor-core mode records an
 * internal CUDA event on its stream right BEFORE launching the main scan kernel; orag_stream_wait_prescan(s) makes
 * stream s wait for that event (cudaStreamWaitEvent; no host sync).  A BM25 call with ORAG_BM25_BACKGROUND does that
 * wait itself, between its query preparation and its first-pass kernel: the first pass starts once the scan's CTAs are
 * (about to be) resident and fills the SM resources the scan leaves idle, instead of grabbing the SMs first and delaying
 * the scan.  (Issue the cosine call of a batch before its BM25 call.) */
int orag_cosine_mark_prescan(int enable);
int orag_stream_wait_prescan(void *stream);

/* Dense float64 cosine matrix, d_out[q * n_rows + r] (test / small-N helper; same arithmetic).
 * d_out must hold n_queries * n_rows + n_queries doubles (the tail receives sum(q*q) per query). */
int orag_cosine_dense(const float *d_corpus, int64_t n_rows, int dim, const float *d_queries, int n_queries,
                      double *d_out, void *stream);

/* The three float64 sums of the reference's cosine WITHOUT the final sqrt / divide, for callers whose reference
 * finishes differently -- `cosine_similarity` of rag/nodes/helpers.py:266-290 takes the magnitudes as `sum ** 0.5`
 * (libm pow), not sqrt: d_out_dots[q * n_rows + r] = sum(a*b) (Neumaier, reference order), d_out_row_sq[r] = sum(a*a)
 * of corpus row r, d_out_query_sq[q] likewise for the queries. */
int orag_dot_dense(const float *d_corpus, int64_t n_rows, int dim, const float *d_queries, int n_queries,
                   double *d_out_dots, double *d_out_row_sq, double *d_out_query_sq, void *stream);

/* First-pass debug/test hook: raw tensor-core similarities (dot * inv_norm[row]) for rows
 * [0, n_rows) as fp32 d_out[r * 256 + q]; n_rows is rounded up to 128 internally, d_out must
 * hold round_up(n_rows,128) * 256 floats.  mode = ORAG_COS_TF32 / ORAG_COS_BF16 / ORAG_COS_F16 (the 16-bit modes need a
 * workspace of n_queries * dim * 2 + 2048 bytes). */
int orag_cosine_firstpass_dense(const float *d_corpus, const float *d_inv_norm, const void *d_shadow, int64_t n_rows,
                                int dim, const float *d_queries, int n_queries, int mode, float *d_out,
                                void *d_workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------
 * BM25 top-k over a GPU-resident, doc-range-tiled inverted index.  Replaces
 * rank_bm25.BM25Okapi.get_scores + the normalisation/sort glue at
 * rag/retrieval.py:324-347, 320.  Bit-exact float64 (no FMA contraction, query
 * terms accumulated per document in query order).  Two candidate paths produce the same result:
 *   - float64 scatter of every posting (bm25.cu), and
 *   - when the index carries d_postings_r16: an fp32 MaxScore first pass (bm25_ms.cu) that only scatters the
 *     postings of a query's "essential" terms, looks the others up for the docs that could still reach the
 *     running threshold, and re-scores the surviving candidates in the float64 arithmetic above.
 * ------------------------------------------------------------------------- */
typedef struct orag_bm25_index {
    int64_t n_docs;                /* docs in this shard */
    int32_t vocab;                 /* term ids are [0, vocab) */
    int32_t tile_docs;             /* docs per tile (power of two, <= 65536) */
    int32_t n_tiles;               /* ceil(n_docs / tile_docs) */
    int32_t has_negative_idf;      /* 1 if a negative idf survives the epsilon floor (forces the dense path) */
    int32_t max_doc_len;           /* size of d_t4_table - 1 */
    int32_t reserved;              /* bit 0: first-pass runs are 16-byte aligned and padded to four postings
                                      (the layout orag_bm25_index_fill writes; required by the first-pass kernel) */
    const int64_t *d_tile_base;    /* [n_tiles + 1] first posting of each tile */
    const int32_t *d_tile_term_off;/* [n_tiles, vocab + 1] term offsets relative to the tile base */
    const uint32_t *d_postings;    /* [P] (doc_in_tile << 16) | tf, ascending doc within (tile, term) */
    const int32_t *d_doc_len;      /* [n_docs] tokens per doc (<= 65535) */
    const double *d_t4_table;      /* [max_doc_len + 1] k1 * (1 - b + b * dl / avgdl), global avgdl */
    const double *d_r_table;       /* [max_doc_len + 1, 4] tf*(k1+1) / (tf + t4[dl]) for tf = 1..4 */
    const double *d_idf;           /* [vocab] global idf incl. epsilon floor; 0 for unseen terms */
    /* Optional first-pass view (pointers NULL = float64 scatter kernel only): a second tiling of the same
     * postings with tiles of fp_tile_docs docs (power of two, 32..16384),
     * (doc_in_tile << 16) | fp16 bits of r = tf*(k1+1)/(tf + t4[dl]) rounded to nearest (every r must be a
     * NORMAL fp16 number), 16-byte aligned, ascending doc within (tile, term); runs padded as described at
     * orag_bm25_index_plan.
     * d_term_max_r[t] = max over this shard's postings of term t of that fp16 value (0 if none). */
    const uint32_t *d_postings_r16;
    const float *d_term_max_r;          /* [vocab] */
    int32_t fp_tile_docs;
    int32_t fp_n_tiles;                 /* ceil(n_docs / fp_tile_docs) */
    const int64_t *d_fp_tile_base;      /* [fp_n_tiles + 1] */
    const int32_t *d_fp_tile_term_off;  /* [fp_n_tiles, vocab + 1] */
    /* Optional threshold warm start of the first pass (NULL = start cold): d_term_kth_r[l * vocab + t] = the K_l-th
     * largest fp16 impact among this shard's postings of term t, K = {10, 16, 32, 64, 128}, 0 when the term has fewer
     * postings (orag_bm25_term_kth; for a term with 16384 or more postings inside a block of 256 first-pass tiles: the
     * K_l-th largest among the postings up to the end of that block -- any subset gives a valid bound).  K_l docs
     * score at least weight(t) * that value whatever else they contain (all
     * contributions are >= 0), so max over a query's terms is a lower bound of its K_l-th best score: the first pass
     * starts with the frequent, low-idf terms already non-essential instead of scoring whole tiles until the running
     * threshold has caught up. */
    const float *d_term_kth_r;          /* [ORAG_BM25_KTH_LEVELS, vocab] */
} orag_bm25_index_t;
#define ORAG_BM25_KTH_LEVELS 5

/* ---------------------------------------------------------------------------
 * Index build (ingest): token corpus -> the two posting tilings above.  Replaces what `BM25Okapi(tokenized_corpus)`
 * derives on every call (rag/retrieval.py:334-338; rank_bm25 0.2.2 BM25._initialize): per-document term
 * frequencies and lengths, document frequencies, the first-seen order of the vocabulary.
 *   d_doc_off int64 [n_docs + 1], d_tokens int32 [d_doc_off[n_docs]] (ids in [0, vocab)); tile_docs / fp_tile_docs
 *   powers of two (32..65536 / 32..16384).  The first-pass view built here has every (tile, term) run start on a
 *   16-byte boundary and padded to a multiple of four postings with copies of its last doc carrying impact +0.0; set
 *   bit 0 of orag_bm25_index_t.reserved to tell orag_bm25_topk so.
 * Phase 1, orag_bm25_index_plan (SYNCHRONISES: the array sizes are an output): writes d_doc_len [n_docs], ADDS this
 *   shard's document frequencies to d_df [vocab] and min-s the global token position of every term's first occurrence
 *   (token_pos_base + local position) into d_first_pos [vocab] (callers initialise them to 0 / INT64_MAX; summing /
 *   min-ing over shards gives the global statistics), writes the final d_tile_base [n_tiles + 1], d_tile_term_off
 *   [n_tiles, vocab + 1] and their fp_ twins, d_info[0..3] = {max tf, max doc length, error bits (1 token id out of
 *   range, 2 doc longer than 65535, 4 internal table overflow, 8 tile >= 2^31 postings, 16 set by the fill: an impact
 *   that is not a normal fp16 number -> first-pass view unusable), 0}, and h_totals[0..1] (HOST) = postings of the
 *   exact view, postings (padded) of the first-pass view.
 * Phase 2, orag_bm25_index_fill: scatters every posting to its slot (same workspace, untouched in between).
 *   d_t4_table [max_doc_len + 1] from the GLOBAL avgdl; d_postings [h_totals[0] + 4]; d_postings_r16 [h_totals[1] + 4]
 *   (16-byte aligned) and d_term_max_r [vocab], or both NULL for an index without first-pass view.
 * ------------------------------------------------------------------------- */
size_t orag_bm25_build_workspace_bytes(int64_t n_docs, int vocab, int tile_docs, int fp_tile_docs);
/* After the fill: d_term_kth_r [ORAG_BM25_KTH_LEVELS, vocab] from the first-pass view of `index` (see the struct). */
int orag_bm25_term_kth(const orag_bm25_index_t *index, float *d_term_kth_r, void *stream);
int orag_bm25_index_plan(const int64_t *d_doc_off, const int32_t *d_tokens, int64_t n_docs, int vocab, int tile_docs,
                         int fp_tile_docs, int64_t token_pos_base, int32_t *d_doc_len, int64_t *d_df,
                         int64_t *d_first_pos, int64_t *d_tile_base, int32_t *d_tile_term_off, int64_t *d_fp_tile_base,
                         int32_t *d_fp_tile_term_off, int32_t *d_info, void *d_workspace, size_t workspace_bytes,
                         int64_t *h_totals, void *stream);
int orag_bm25_index_fill(const int64_t *d_doc_off, const int32_t *d_tokens, int64_t n_docs, int vocab, int tile_docs,
                         int fp_tile_docs, const double *d_t4_table, int max_doc_len, const int64_t *d_tile_base,
                         const int32_t *d_tile_term_off, uint32_t *d_postings, const int64_t *d_fp_tile_base,
                         const int32_t *d_fp_tile_term_off, uint32_t *d_postings_r16, float *d_term_max_r,
                         int32_t *d_info, void *d_workspace, size_t workspace_bytes, void *stream);

/*   d_query_terms int32 [n_queries, max_terms], entries < 0 or >= vocab are OOV / padding; max_terms <= 64 on the
 *   candidate paths, any length with ORAG_BM25_FORCE_DENSE and in orag_bm25_dense (per-document evaluation)
 *   d_query_lens  int32 [n_queries]
 *   outputs: top-k by (normalised score desc, id asc) with ORAG_BM25_NORMALIZE, and
 *   d_out_max[q] = the divisor the reference uses (max raw score, or 1.0 when that max is <= 0);
 *   without it the raw scores are returned (multi-GPU: normalise after the gather) and
 *   d_out_max[q] is the shard's max raw score. */
#define ORAG_BM25_NORMALIZE 1    /* divide by the max raw score (rag/retrieval.py:343-345) */
#define ORAG_BM25_FORCE_SPARSE 2 /* candidate path even for small corpora (tests) */
#define ORAG_BM25_FORCE_DENSE 4  /* dense accumulate + exact select (small N, negative idf, fallback) */
#define ORAG_BM25_EXACT_TILES 8  /* candidate path through the float64 scatter kernel even when the index carries
                                    the fp16 first-pass view (A/B tests; queries longer than 32 terms use it anyway) */
#define ORAG_BM25_BACKGROUND 16  /* size the first-pass launch (8-warp CTAs, < 31 KB shared memory) so that it runs NEXT TO a
                                    resident cosine scan CTA on every SM instead of after it; meant for a side stream:
                                    the first-pass launch waits for the latest pre-scan event (orag_cosine_mark_prescan) */
size_t orag_bm25_workspace_bytes(const orag_bm25_index_t *index, int n_queries, int k, int flags);
int orag_bm25_topk(const orag_bm25_index_t *index, int64_t doc_id_base, const int32_t *d_query_terms,
                   const int32_t *d_query_lens, int n_queries, int max_terms, int k, int flags,
                   int64_t *d_out_ids, double *d_out_scores, double *d_out_max, int32_t *d_out_status,
                   void *d_workspace, size_t workspace_bytes, void *stream);
/* Dense raw scores d_out[q * n_docs + d] (same arithmetic; test / fallback / small N). */
int orag_bm25_dense(const orag_bm25_index_t *index, const int32_t *d_query_terms, const int32_t *d_query_lens,
                    int n_queries, int max_terms, double *d_out, void *stream);

/* Top-k of dense fp64 score rows: d_scores[q * ld + i], i in [0, n).  With `normalize`,
 * rows are first divided by max(row) if that max is > 0 (rag/retrieval.py:343-345). */
int orag_dense_topk(const double *d_scores, int64_t n, int64_t ld, int n_queries, int k, int64_t id_base,
                    int normalize, int64_t *d_out_ids, double *d_out_scores, double *d_out_max, void *stream);

/* ---------------------------------------------------------------------------
 * Merge of per-shard candidate lists after the all-gather:
 * d_cand_ids/scores [n_queries, m] (id -1 = empty) -> top-k by (score desc, id asc).
 * With d_shard_max != NULL ([n_queries, n_shards] raw BM25 maxima) scores are divided
 * by the global max first (or 1.0 when it is <= 0) and *d_out_max receives it.
 * ------------------------------------------------------------------------- */
int orag_topk_merge(const int64_t *d_cand_ids, const double *d_cand_scores, int m, int n_queries, int k,
                    const double *d_shard_max, int n_shards, int64_t *d_out_ids, double *d_out_scores,
                    double *d_out_max, void *stream);

/* ---------------------------------------------------------------------------
 * Reciprocal Rank Fusion.  Replaces ReciprocalRankFusion.fuse (rag/reranker.py:224-271):
 * score[id] = sum over lists of 1/(rrf_k + rank), float64, lists walked in order;
 * stable descending sort over first-sighting order (tie_mode 0) or ascending id
 * (tie_mode 1).  d_list_ids [n_queries, n_lists, list_len] (id < 0 = padding, skipped
 * WITHOUT consuming a rank only at the tail of a list).
 *   d_out_src (optional, int32 [n_queries, top_k, n_lists]) 1-based rank of the fused
 *   item in each input list, 0 if absent.
 * ------------------------------------------------------------------------- */
int orag_rrf_fuse(const int64_t *d_list_ids, int n_queries, int n_lists, int list_len, int rrf_k, int top_k,
                  int tie_mode, int64_t *d_out_ids, double *d_out_scores, int32_t *d_out_src, void *stream);

/* Two-list form for the one-shard hybrid step: d_ids_a / d_ids_b [n_queries, list_len] are the ranked lists exactly as
 * orag_cosine_topk / orag_bm25_topk wrote them (no packing); same arithmetic and tie rule as orag_rrf_fuse with
 * n_lists = 2 (d_out_src int32 [n_queries, top_k, 2], optional).  d_out_status[q] (optional) = d_status_a[q] |
 * d_status_b[q] (either may be NULL): the OR of the two candidate-overflow words.  2 * list_len <= 128. */
int orag_rrf_fuse_pair(const int64_t *d_ids_a, const int64_t *d_ids_b, int n_queries, int list_len, int rrf_k, int top_k,
                       int tie_mode, const int32_t *d_status_a, const int32_t *d_status_b, int64_t *d_out_ids,
                       double *d_out_scores, int32_t *d_out_src, int32_t *d_out_status, void *stream);

/* Everything after the all-gather of a row-sharded hybrid search in one launch: d_gathered is
 * [n_shards, n_queries, W] int64 with W = 2*fetch_k + 2*kk + 2 holding, per shard and query,
 * cosine ids | cosine float64 score bits | BM25 ids | BM25 RAW float64 score bits | the shard's max raw
 * BM25 score bits | ORAG_STATUS_* bits.  Cosine winners are merged by (score desc, id asc); BM25 raw scores
 * are divided by the global max (or 1.0 if it is <= 0; rag/retrieval.py:343-345) and merged the same way;
 * the two fetch_k-long lists are fused as in orag_rrf_fuse.  Outputs: fused ids/scores [n_queries, top_k],
 * d_out_src int32 [n_queries, top_k, 2] (optional), the two merged lists [n_queries, fetch_k], the BM25
 * divisor [n_queries] and the OR of the shards' status bits [n_queries] (optional).
 * Limits: n_shards * kk <= 256, fetch_k <= 64, 2 * fetch_k <= 128. */
int orag_hybrid_merge(const int64_t *d_gathered, int n_shards, int n_queries, int fetch_k, int kk, int rrf_k,
                      int top_k, int tie_mode, int64_t *d_out_ids, double *d_out_scores, int32_t *d_out_src,
                      int64_t *d_cos_ids, double *d_cos_scores, int64_t *d_bm25_ids, double *d_bm25_scores,
                      double *d_bm25_max, int32_t *d_out_status, void *stream);

/* ---------------------------------------------------------------------------
 * Peer-memory exchange for the row-sharded search (one process per GPU, NVLink / NVSwitch): replaces
 * "pack + all-gather" in front of orag_hybrid_merge.  The reference has no counterpart (single Postgres);
 * the data exchanged is the merge input described above.
 *   setup, once per (n_shards, max_queries, fetch_k, kk):
 *     orag_exchange_bytes -> orag_exchange_alloc (cudaMalloc + zero fill: the ONE place the library allocates
 *     device memory, because an IPC handle names a whole allocation) -> orag_exchange_export (64-byte handle,
 *     sent to the peers by any host channel) -> orag_exchange_open on every peer's handle -> a DEVICE array
 *     d_peer_bufs[n_shards] of the buffers in rank order (own buffer at index rank) -> host barrier.
 *   per search, with the same seq = 1, 2, 3, ... on every rank:
 *     orag_hybrid_push: one launch packs this rank's lists (same arrays / layout as the gathered buffer of
 *       orag_hybrid_merge: d_cos_* [n_queries, fetch_k], d_bm25_* [n_queries, kk] RAW scores, d_bm25_max
 *       [n_queries]; d_status / d_status2 [n_queries] or NULL: the status words of the cosine and the BM25 call,
 *       OR-ed into the block's status column) and stores them into slot seq % 6 of EVERY peer, then publishes
 *       seq with a system-scope release store.
 *     orag_hybrid_wait: one tiny launch that acquires the n_shards sequence numbers in this rank's own buffer;
 *       *d_gathered (host out) is the [n_shards, n_queries, W] array to hand to orag_hybrid_merge on the same
 *       stream.  A block that has not arrived after timeout_ms gets ORAG_STATUS_EXCHANGE_TIMEOUT in its status
 *       words (the merge ORs them into d_out_status) instead of hanging the GPU.
 *   Every rank must call push and wait for every seq, in order, with the same n_queries / fetch_k / kk.  Three searches
 *   may be in flight at once on three streams (lane = seq % 3, each lane in stream order): the buffer has six slots.
 * ------------------------------------------------------------------------- */
size_t orag_exchange_bytes(int n_shards, int max_queries, int fetch_k, int kk);
int orag_exchange_alloc(size_t bytes, void **d_buf);
int orag_exchange_free(void *d_buf);
int orag_exchange_export(void *d_buf, unsigned char handle[64]);
int orag_exchange_open(const unsigned char handle[64], void **d_peer_buf);
int orag_exchange_close(void *d_peer_buf);
int orag_hybrid_push(const int64_t *d_cos_ids, const double *d_cos_scores, const int64_t *d_bm25_ids,
                     const double *d_bm25_scores, const double *d_bm25_max, const int32_t *d_status,
                     const int32_t *d_status2, int n_queries, int fetch_k, int kk, int rank, int n_shards,
                     int max_queries, void *const *d_peer_bufs, uint64_t seq, void *stream);
int orag_hybrid_wait(void *d_buf, int n_shards, int max_queries, int n_queries, int fetch_k, int kk, uint64_t seq,
                     int timeout_ms, const int64_t **d_gathered, void *stream);

/* Weighted hybrid score of HybridRetriever.hybrid_search (rag/retrieval.py:302):
 * out[i] = (alpha*sem[i] + beta*kw[i]) + gamma*temp[i] in float64 without contraction
 * (d_temp may be NULL = all zero).  Rank the result with orag_dense_topk. */
int orag_weighted_sum3(const double *d_sem, const double *d_kw, const double *d_temp, int64_t n, double alpha,
                       double beta, double gamma, double *d_out, void *stream);
/* d_out[i] = d_in[i] / divisor, one IEEE division each (`s / max_score`, rag/retrieval.py:345). */
int orag_div_scalar(const double *d_in, int64_t n, double divisor, double *d_out, void *stream);

/* ---------------------------------------------------------------------------
 * Pairwise cosine candidates (rag/consistency_checker.py:169-189): all i<j with
 * doc_idx[i] != doc_idx[j] and float64 cosine >= threshold.  Pairs are written
 * unordered; *d_out_count receives the total (may exceed cap -> truncated).
 * ------------------------------------------------------------------------- */
size_t orag_pairwise_workspace_bytes(int64_t m, int dim);
int orag_pairwise_cosine_threshold(const float *d_emb, int64_t m, int dim, const int32_t *d_doc_idx,
                                   double threshold, int64_t cap, int32_t *d_out_i, int32_t *d_out_j,
                                   double *d_out_sim, unsigned long long *d_out_count, void *d_workspace,
                                   size_t workspace_bytes, void *stream);

/* Tensor-core variant (BASELINE config 5, 64k x 1536): a tcgen05 first pass with a fixed threshold
 * (threshold - first-pass error bound) over the triangular space of (256-row block, 128-row tile) pairs -- fp16
 * shadow rows scaled by powers of two, kind::f16, CTA pairs with cta_group::2 when dim % 64 == 0, else tf32 off the
 * fp32 rows -- then the same float64 re-score and filter -> identical pair set.  Needs dim % 32 == 0 and
 * threshold > ~2.3e-3.  d_out_count has TWO elements: [0] = number of pairs, [1] = 1 if a per-row candidate
 * buffer overflowed (result incomplete: re-run with orag_pairwise_cosine_threshold).
 *   orag_pairwise_prepare   once per claim matrix: float64 sum(a*a) of every row, first-pass norms, fp16 shadow, into a
 *                           caller-owned buffer of orag_pairwise_prepared_bytes(m, dim) bytes
 *   orag_pairwise_pairs     the search over a prepared matrix (workspace: orag_pairwise_pairs_workspace_bytes(m))
 *   orag_pairwise_cosine_threshold_tc   both in one call (workspace: orag_pairwise_tc_workspace_bytes) */
size_t orag_pairwise_prepared_bytes(int64_t m, int dim);
size_t orag_pairwise_pairs_workspace_bytes(int64_t m);
int orag_pairwise_prepare(const float *d_emb, int64_t m, int dim, void *d_prepared, size_t prepared_bytes, void *stream);
int orag_pairwise_pairs(const float *d_emb, const void *d_prepared, int64_t m, int dim, const int32_t *d_doc_idx,
                        double threshold, int64_t cap, int32_t *d_out_i, int32_t *d_out_j, double *d_out_sim,
                        unsigned long long *d_out_count, void *d_workspace, size_t workspace_bytes, void *stream);
size_t orag_pairwise_tc_workspace_bytes(int64_t m, int dim);
int orag_pairwise_cosine_threshold_tc(const float *d_emb, int64_t m, int dim, const int32_t *d_doc_idx,
                                      double threshold, int64_t cap, int32_t *d_out_i, int32_t *d_out_j,
                                      double *d_out_sim, unsigned long long *d_out_count, void *d_workspace,
                                      size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ORAG_H */
