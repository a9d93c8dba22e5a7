/*
 * oracle.c -- CPU restatement of optimized-rag's hybrid retrieval arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under optimized_rag_b200/ may import, link
 * or execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and there only as the checker /
 * the reported CPU baseline.
 *
 * Parity pin: the reference ships no tests or golden vectors, so
 * this restatement is pinned against the reference's OWN Python functions run in
 * the build container (tests/golden/make_golden.py imports
 * /root/reference/rag/retrieval.py and rag/reranker.py by path and records their
 * outputs; tests/test_oracle_golden.py replays them against this file).
 * BM25 is the exception: the arithmetic lives in the third-party package
 * rank-bm25 (requirements.txt:22, ">=0.2.2", no lock file) which is absent from
 * /root/reference and from this image, so the BM25 functions below restate the
 * published algorithm of rank_bm25 0.2.2 `BM25Okapi` -> "BM25 parity unpinned
 * against the third-party package; pinned against oracle/rank_bm25.py driven
 * through the reference's own HybridRetriever._bm25_scores glue".
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -fno-fast-math -shared -fPIC
 * (no FMA contraction: every operation below is one IEEE-754 binary64 rounding,
 * in the order the Python source performs it).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* Cosine: rag/retrieval.py:362-371 (same body at rag/consistency_checker.py: */
/* 241-261, rag/reranker.py:92-101).                                          */
/*   dot = sum(a*b for a,b in zip(v1,v2)); m1 = sqrt(sum(a*a)); m2 = ...      */
/*   0.0 if m1 == 0 or m2 == 0 else dot / (m1*m2)                             */
/* `sum()` over Python floats is Neumaier-compensated since CPython 3.12       */
/* (Python/bltinmodule.c builtin_sum); mode 0 below is the pre-3.12 plain      */
/* left-to-right sum.  The container's interpreter is 3.12.3 -> mode 1 is the */
/* behaviour the golden vectors record.                                       */
/* ------------------------------------------------------------------------- */

static double sum_prod(const float *a, const float *b, int d, int neumaier)
{
    if (d <= 0) return 0.0;
    /* fp32 widened exactly to binary64; product of two widened fp32 is exact */
    /* sum() starts from int 0, so the first item enters as 0 + x0 (== 0.0 + x0 in binary64) */
    double s = 0.0 + (double)a[0] * (double)b[0];
    if (!neumaier) {
        for (int i = 1; i < d; ++i) s = s + (double)a[i] * (double)b[i];
        return s;
    }
    double c = 0.0;
    for (int i = 1; i < d; ++i) {
        double x = (double)a[i] * (double)b[i];
        double t = s + x;
        if (fabs(s) >= fabs(x)) c += (s - t) + x;
        else                    c += (x - t) + s;
        s = t;
    }
    if (c != 0.0 && isfinite(c)) s += c;
    return s;
}

/* Threads of the row loop below (a launcher such as torchrun exports OMP_NUM_THREADS=1, which would turn the
 * "all host cores" CPU baseline into a single-core one).  Returns the previous maximum. */
#ifdef _OPENMP
#include <omp.h>
#endif
int orc_set_threads(int n)
{
#ifdef _OPENMP
    int prev = omp_get_max_threads();
    if (n > 0) omp_set_num_threads(n);
    return prev;
#else
    (void)n;
    return 1;
#endif
}

/* sum(a*b for a,b in zip(v1,v2)) alone (callers that finish the cosine differently, e.g. rag/nodes/helpers.py:266-290
 * with `** 0.5`) */
double orc_dot(const float *a, const float *b, int d, int neumaier) { return sum_prod(a, b, d, neumaier); }

double orc_cosine(const float *a, const float *b, int d, int neumaier)
{
    double dot = sum_prod(a, b, d, neumaier);
    double m1 = sqrt(sum_prod(a, a, d, neumaier));
    double m2 = sqrt(sum_prod(b, b, d, neumaier));
    if (m1 == 0.0 || m2 == 0.0) return 0.0;
    return dot / (m1 * m2);
}

/* scores[r] = cosine(query, corpus[r]) for r in [0,n): the loop at
 * rag/retrieval.py:252-256 (argument order: query first). */
void orc_cosine_scores(const float *corpus, int64_t n, int d, const float *query,
                       int neumaier, double *scores)
{
    double m1 = sqrt(sum_prod(query, query, d, neumaier));
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r) {
        const float *row = corpus + r * (int64_t)d;
        double dot = sum_prod(query, row, d, neumaier);
        double m2 = sqrt(sum_prod(row, row, d, neumaier));
        scores[r] = (m1 == 0.0 || m2 == 0.0) ? 0.0 : dot / (m1 * m2);
    }
}

/* Top-k of a dense score vector, ordering = Python's stable
 * sorted(..., reverse=True)[:k] over index-ordered input (rag/retrieval.py:320):
 * descending score, ties -> ascending index.  Returns count = min(k, n). */
int orc_topk(const double *scores, int64_t n, int k, int64_t id_base,
             int64_t *out_ids, double *out_scores)
{
    int cnt = 0;
    for (int64_t i = 0; i < n; ++i) {
        double s = scores[i];
        if (cnt == k && !(s > out_scores[k - 1])) continue;
        int pos = (cnt < k) ? cnt : k - 1;
        /* strict '>' keeps earlier (lower) indices ahead on ties */
        while (pos > 0 && s > out_scores[pos - 1]) {
            out_scores[pos] = out_scores[pos - 1];
            out_ids[pos] = out_ids[pos - 1];
            --pos;
        }
        out_scores[pos] = s;
        out_ids[pos] = id_base + i;
        if (cnt < k) ++cnt;
    }
    return cnt;
}

/* ------------------------------------------------------------------------- */
/* BM25: rag/retrieval.py:324-347 glue + rank_bm25 0.2.2 BM25Okapi            */
/* (k1=1.5, b=0.75, epsilon=0.25).  Tokens are int32 ids in [0, vocab).       */
/* ------------------------------------------------------------------------- */

typedef struct {
    int64_t n_docs;
    int32_t vocab;
    double avgdl;
    double average_idf;
    double eps;
    int32_t n_terms_seen;  /* len(self.idf) */
    int32_t *dl;           /* [n_docs] */
    int64_t *df;           /* [vocab] */
    double *idf;           /* [vocab]; 0.0 for unseen terms (== `idf.get(q) or 0`) */
    int32_t *first_seen;   /* [n_terms_seen] term ids in dict-insertion order */
    int64_t *post_off;     /* [vocab+1] */
    int32_t *post_doc;     /* [P] ascending doc id within a term */
    int32_t *post_tf;      /* [P] */
} orc_bm25_t;

void orc_bm25_free(orc_bm25_t *ix)
{
    if (!ix) return;
    free(ix->dl); free(ix->df); free(ix->idf); free(ix->first_seen);
    free(ix->post_off); free(ix->post_doc); free(ix->post_tf); free(ix);
}

/* BM25._initialize + BM25Okapi._calc_idf.  doc_off[n_docs+1] indexes tokens[]. */
orc_bm25_t *orc_bm25_build(const int64_t *doc_off, const int32_t *tokens,
                           int64_t n_docs, int32_t vocab)
{
    orc_bm25_t *ix = (orc_bm25_t *)calloc(1, sizeof(*ix));
    ix->n_docs = n_docs; ix->vocab = vocab;
    ix->dl = (int32_t *)calloc((size_t)(n_docs > 0 ? n_docs : 1), sizeof(int32_t));
    ix->df = (int64_t *)calloc((size_t)vocab + 1, sizeof(int64_t));
    ix->idf = (double *)calloc((size_t)vocab + 1, sizeof(double));
    ix->first_seen = (int32_t *)calloc((size_t)vocab + 1, sizeof(int32_t));
    ix->post_off = (int64_t *)calloc((size_t)vocab + 2, sizeof(int64_t));

    int32_t *last_doc = (int32_t *)malloc(((size_t)vocab + 1) * sizeof(int32_t));
    for (int32_t t = 0; t < vocab; ++t) last_doc[t] = -1;

    /* pass 1: dl, df (insertion order of nd == first occurrence scanning docs in
     * order, tokens in order), total length */
    int64_t num_doc = 0;
    int32_t seen = 0;
    for (int64_t d = 0; d < n_docs; ++d) {
        int64_t lo = doc_off[d], hi = doc_off[d + 1];
        ix->dl[d] = (int32_t)(hi - lo);
        num_doc += hi - lo;
        for (int64_t p = lo; p < hi; ++p) {
            int32_t t = tokens[p];
            if (last_doc[t] != (int32_t)d) {
                if (ix->df[t] == 0) ix->first_seen[seen++] = t;
                last_doc[t] = (int32_t)d;
                ix->df[t] += 1;
            }
        }
    }
    ix->n_terms_seen = seen;
    ix->avgdl = (n_docs > 0) ? (double)num_doc / (double)n_docs : 0.0;

    /* idf in dict order; idf_sum is a plain `+=` on Python floats (no sum()) */
    double idf_sum = 0.0;
    for (int32_t i = 0; i < seen; ++i) {
        int32_t t = ix->first_seen[i];
        double freq = (double)ix->df[t];
        double v = log((double)n_docs - freq + 0.5) - log(freq + 0.5);
        ix->idf[t] = v;
        idf_sum += v;
    }
    ix->average_idf = (seen > 0) ? idf_sum / (double)seen : 0.0;
    ix->eps = 0.25 * ix->average_idf;
    for (int32_t i = 0; i < seen; ++i) {
        int32_t t = ix->first_seen[i];
        if (ix->idf[t] < 0.0) ix->idf[t] = ix->eps;
    }

    /* pass 2: postings (term -> ascending (doc, tf)) */
    for (int32_t t = 0; t < vocab; ++t) ix->post_off[t + 1] = ix->post_off[t] + ix->df[t];
    int64_t P = ix->post_off[vocab];
    ix->post_doc = (int32_t *)malloc((size_t)(P > 0 ? P : 1) * sizeof(int32_t));
    ix->post_tf = (int32_t *)malloc((size_t)(P > 0 ? P : 1) * sizeof(int32_t));
    int64_t *cursor = (int64_t *)malloc(((size_t)vocab + 1) * sizeof(int64_t));
    memcpy(cursor, ix->post_off, ((size_t)vocab + 1) * sizeof(int64_t));
    for (int32_t t = 0; t < vocab; ++t) last_doc[t] = -1;
    for (int64_t d = 0; d < n_docs; ++d) {
        for (int64_t p = doc_off[d]; p < doc_off[d + 1]; ++p) {
            int32_t t = tokens[p];
            if (last_doc[t] != (int32_t)d) {
                last_doc[t] = (int32_t)d;
                ix->post_doc[cursor[t]] = (int32_t)d;
                ix->post_tf[cursor[t]] = 1;
                cursor[t] += 1;
            } else {
                ix->post_tf[cursor[t] - 1] += 1;
            }
        }
    }
    free(cursor); free(last_doc);
    return ix;
}

/* accessors for the ctypes wrapper */
double orc_bm25_avgdl(const orc_bm25_t *ix) { return ix->avgdl; }
double orc_bm25_average_idf(const orc_bm25_t *ix) { return ix->average_idf; }
double orc_bm25_eps(const orc_bm25_t *ix) { return ix->eps; }
int32_t orc_bm25_n_terms(const orc_bm25_t *ix) { return ix->n_terms_seen; }
const double *orc_bm25_idf(const orc_bm25_t *ix) { return ix->idf; }
const int64_t *orc_bm25_df(const orc_bm25_t *ix) { return ix->df; }
const int32_t *orc_bm25_dl(const orc_bm25_t *ix) { return ix->dl; }
const int32_t *orc_bm25_first_seen(const orc_bm25_t *ix) { return ix->first_seen; }
int64_t orc_bm25_num_postings(const orc_bm25_t *ix) { return ix->post_off[ix->vocab]; }
const int64_t *orc_bm25_post_off(const orc_bm25_t *ix) { return ix->post_off; }
const int32_t *orc_bm25_post_doc(const orc_bm25_t *ix) { return ix->post_doc; }
const int32_t *orc_bm25_post_tf(const orc_bm25_t *ix) { return ix->post_tf; }

/* BM25Okapi.get_scores: score starts at zeros; for each query token IN QUERY
 * ORDER (duplicates repeated):
 *   score += idf * (tf*(k1+1) / (tf + k1*(1 - b + b*dl/avgdl)))
 * evaluated with numpy float64 elementwise ops in exactly this order.  For
 * tf == 0 the term is +-0.0 and the add is a no-op, so walking the posting list
 * is bit-identical to the dense evaluation.  Tokens outside [0,vocab) are OOV. */
void orc_bm25_scores_raw(const orc_bm25_t *ix, const int32_t *query, int lq, double *score)
{
    const double k1 = 1.5, b = 0.75;
    const double one_minus_b = 1 - b;  /* 0.25, formed first as a Python scalar */
    for (int64_t d = 0; d < ix->n_docs; ++d) score[d] = 0.0;
    for (int i = 0; i < lq; ++i) {
        int32_t t = query[i];
        if (t < 0 || t >= ix->vocab) continue;
        double idf = ix->idf[t];
        if (idf == 0.0) continue;  /* `self.idf.get(q) or 0` */
        for (int64_t p = ix->post_off[t]; p < ix->post_off[t + 1]; ++p) {
            int32_t d = ix->post_doc[p];
            double tf = (double)ix->post_tf[p];
            double t1 = b * (double)ix->dl[d];
            double t2 = t1 / ix->avgdl;
            double t3 = one_minus_b + t2;
            double t4 = k1 * t3;
            double den = tf + t4;
            double num = tf * (k1 + 1);
            double r = num / den;
            double c = idf * r;
            score[d] = score[d] + c;
        }
    }
}

/* rag/retrieval.py:343-345: max_score = max(scores) if len>0 and max>0 else 1.0;
 * normalized = s / max_score.  Returns max_score used. */
double orc_bm25_normalize(const double *raw, int64_t n, double *norm)
{
    double m = 1.0;
    if (n > 0) {
        double mx = raw[0];
        for (int64_t i = 1; i < n; ++i) if (raw[i] > mx) mx = raw[i];
        if (mx > 0.0) m = mx;
    }
    for (int64_t i = 0; i < n; ++i) norm[i] = raw[i] / m;
    return m;
}

/* ------------------------------------------------------------------------- */
/* RRF: rag/reranker.py:224-271.  Keys are int64 ids (the reference keys on    */
/* the content string; equivalent when contents are unique).  Tie rule = dict  */
/* insertion order under a stable descending sort.                            */
/* tie_mode 0 = reference (insertion order), 1 = ascending id.                */
/* ------------------------------------------------------------------------- */
int orc_rrf_fuse(const int64_t *ids_flat, const int32_t *list_len, int n_lists,
                 int rrf_k, int top_k, int tie_mode, int64_t *out_ids, double *out_scores)
{
    int total = 0;
    for (int l = 0; l < n_lists; ++l) total += list_len[l];
    int64_t *keys = (int64_t *)malloc((size_t)(total > 0 ? total : 1) * sizeof(int64_t));
    double *sc = (double *)malloc((size_t)(total > 0 ? total : 1) * sizeof(double));
    int m = 0, off = 0;
    for (int l = 0; l < n_lists; ++l) {
        for (int r = 0; r < list_len[l]; ++r) {
            int64_t key = ids_flat[off + r];
            double c = 1.0 / (double)(rrf_k + (r + 1));
            int j = 0;
            for (; j < m; ++j) if (keys[j] == key) break;
            if (j < m) sc[j] = sc[j] + c;
            else { keys[m] = key; sc[m] = c; ++m; }
        }
        off += list_len[l];
    }
    /* stable insertion sort, descending */
    int *ord = (int *)malloc((size_t)(m > 0 ? m : 1) * sizeof(int));
    for (int i = 0; i < m; ++i) {
        int p = i;
        while (p > 0) {
            int q = ord[p - 1];
            int before = (sc[i] > sc[q]) ||
                         (tie_mode == 1 && sc[i] == sc[q] && keys[i] < keys[q]);
            if (!before) break;
            ord[p] = q; --p;
        }
        ord[p] = i;
    }
    int cnt = m < top_k ? m : top_k;
    for (int i = 0; i < cnt; ++i) { out_ids[i] = keys[ord[i]]; out_scores[i] = sc[ord[i]]; }
    free(keys); free(sc); free(ord);
    return cnt;
}

/* ------------------------------------------------------------------------- */
/* Weighted hybrid (rag/retrieval.py:294-322, "next" row f1):                 */
/*   h = alpha*sem + beta*kw + gamma*temp   (Python left-to-right)            */
/* ------------------------------------------------------------------------- */
void orc_weighted_hybrid(const double *sem, const double *kw, const double *temp, int64_t n,
                         double alpha, double beta, double gamma, double *out)
{
    for (int64_t i = 0; i < n; ++i)
        out[i] = (alpha * sem[i] + beta * kw[i]) + gamma * (temp ? temp[i] : 0.0);
}

/* ------------------------------------------------------------------------- */
/* Pairwise consistency candidates (rag/consistency_checker.py:169-189):      */
/* all i<j with doc_idx[i] != doc_idx[j] and cosine >= thr.  Returns count;   */
/* writes at most cap pairs (i, j, sim) in (i, j) lexicographic order.        */
/* ------------------------------------------------------------------------- */
int64_t orc_pairwise_candidates(const float *emb, int64_t m, int d, const int32_t *doc_idx,
                                double thr, int neumaier, int64_t cap,
                                int32_t *out_i, int32_t *out_j, double *out_sim)
{
    int64_t cnt = 0;
    for (int64_t i = 0; i < m; ++i)
        for (int64_t j = i + 1; j < m; ++j) {
            if (doc_idx[i] == doc_idx[j]) continue;
            double s = orc_cosine(emb + i * d, emb + j * d, d, neumaier);
            if (s >= thr) {
                if (cnt < cap) { out_i[cnt] = (int32_t)i; out_j[cnt] = (int32_t)j; out_sim[cnt] = s; }
                ++cnt;
            }
        }
    return cnt;
}


/* ========================================================================= */
/* Streamed oracle for the BASELINE-size parity gates (SURVEY.md §8d "Parity  */
/* gates": the streamed oracle at 10M rows on a query subset).  The 10M x 1536 */
/* fp32 corpus (61.4 GB) and the 2e9-token corpus do not fit host memory, so    */
/* the functions below REGENERATE the synthetic inputs block by block from the  */
/* seeds (same counter-based hashes as optimized_rag_b200/synthetic.py and      */
/* csrc/gen.cu; tests/test_oracle_stream.py checks them against the numpy       */
/* generators) and apply exactly the arithmetic of the functions above:         */
/*   cosine  rag/retrieval.py:362-371 per (query, row), top-k by               */
/*           (score desc, id asc) as rag/retrieval.py:320                      */
/*   BM25    rank_bm25 0.2.2 BM25Okapi._initialize/_calc_idf/get_scores +      */
/*           rag/retrieval.py:324-347                                          */
/* Queries of a batch are evaluated side by side (one SIMD lane per query): each */
/* lane performs the same sequence of IEEE-754 binary64 operations as the        */
/* scalar code, so the results are bit-identical to orc_cosine_scores.           */
/* ========================================================================= */

static inline uint64_t orc_mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
#define ORC_K_ROW 0xD1B54A32D192ED03ull
#define ORC_K_DOC 0x9FB21C651E98DF25ull
#define ORC_K_DUP 0xA24BAED4963EE407ull

static inline uint64_t orc_row_key(uint64_t seed, uint64_t row) { return orc_mix64(seed ^ (row * ORC_K_ROW)); }

/* synthetic.source_rows: the row whose values row r carries */
static inline uint64_t orc_source_row(uint64_t seed, uint64_t row, int dup_per_mille)
{
    if (dup_per_mille <= 0 || row == 0) return row;
    uint64_t h = orc_mix64(seed ^ ORC_K_DUP ^ (row * ORC_K_DOC));
    if ((h % 1000ull) < (uint64_t)dup_per_mille) return orc_mix64(h) % row;
    return row;
}

/* synthetic.embeddings: one row */
static void orc_gen_row(uint64_t seed, uint64_t row, int dim, int dup_per_mille, float *out)
{
    uint64_t key = orc_row_key(seed, orc_source_row(seed, row, dup_per_mille));
    for (int c = 0; c < dim; ++c) {
        uint64_t h = orc_mix64(key + (uint64_t)c);
        int64_t v = (int64_t)(h >> 40) - ((int64_t)1 << 23);
        out[c] = (float)v * 0x1p-28f;
    }
}

void orc_gen_embeddings(uint64_t seed, int64_t row_start, int64_t n_rows, int dim, int dup_per_mille, float *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n_rows; ++r) orc_gen_row(seed, (uint64_t)(row_start + r), dim, dup_per_mille, out + r * dim);
}

/* Dot products of one corpus row with ORC_QB queries at once (queries transposed: qt[i * ORC_QB + j]).
 * Per lane j exactly sum_prod(query_j, row): the argument order of rag/retrieval.py:252-256 (query first) --
 * products of two widened fp32 values are exact, so the order of the factors does not matter. */
#define ORC_QB 16
__attribute__((target_clones("avx512f", "avx2", "default")))
static void orc_dot_block(const float *row, const double *qt, int d, int neumaier, double *out)
{
    double s[ORC_QB], c[ORC_QB];
    const double a0 = (double)row[0];
    for (int j = 0; j < ORC_QB; ++j) { s[j] = 0.0 + a0 * qt[j]; c[j] = 0.0; }
    if (!neumaier) {
        for (int i = 1; i < d; ++i) {
            const double a = (double)row[i];
            const double *q = qt + (size_t)i * ORC_QB;
            for (int j = 0; j < ORC_QB; ++j) s[j] = s[j] + a * q[j];
        }
        for (int j = 0; j < ORC_QB; ++j) out[j] = s[j];
        return;
    }
    for (int i = 1; i < d; ++i) {
        const double a = (double)row[i];
        const double *q = qt + (size_t)i * ORC_QB;
#pragma omp simd
        for (int j = 0; j < ORC_QB; ++j) {
            const double x = a * q[j];
            const double t = s[j] + x;
            const double big = fabs(s[j]) >= fabs(x) ? s[j] : x;
            const double small = fabs(s[j]) >= fabs(x) ? x : s[j];
            c[j] += (big - t) + small;
            s[j] = t;
        }
    }
    for (int j = 0; j < ORC_QB; ++j) {
        double r = s[j];
        if (c[j] != 0.0 && isfinite(c[j])) r += c[j];
        out[j] = r;
    }
}

/* (score desc, id asc) insertion into a k-long list */
static inline void orc_topk_insert(double s, int64_t id, int k, int *cnt, double *sc, int64_t *ids)
{
    if (*cnt == k && !(s > sc[k - 1] || (s == sc[k - 1] && id < ids[k - 1]))) return;
    int pos = (*cnt < k) ? *cnt : k - 1;
    while (pos > 0 && (s > sc[pos - 1] || (s == sc[pos - 1] && id < ids[pos - 1]))) {
        sc[pos] = sc[pos - 1];
        ids[pos] = ids[pos - 1];
        --pos;
    }
    sc[pos] = s;
    ids[pos] = id;
    if (*cnt < k) ++*cnt;
}

/* Exact cosine top-k of `nq` queries against the synthetic rows [row_start, row_start + n_rows) regenerated from
 * `seed` (or, with corpus != NULL, against that in-memory fp32 block: row r of it is global row row_start + r).
 * out_ids/out_scores [nq, k] (-1 / 0.0 padded); ids are global rows.  Returns 0. */
int orc_cosine_topk_stream(uint64_t seed, int64_t row_start, int64_t n_rows, int dim, int dup_per_mille,
                           const float *corpus, const float *queries, int nq, int k, int neumaier,
                           int64_t *out_ids, double *out_scores)
{
    const int nqb = (nq + ORC_QB - 1) / ORC_QB;
    double *qt = (double *)calloc((size_t)nqb * dim * ORC_QB, sizeof(double));
    double *m1 = (double *)calloc((size_t)nqb * ORC_QB, sizeof(double));
    for (int q = 0; q < nq; ++q) {
        for (int i = 0; i < dim; ++i)
            qt[((size_t)(q / ORC_QB) * dim + i) * ORC_QB + (q % ORC_QB)] = (double)queries[(size_t)q * dim + i];
        m1[q] = sqrt(sum_prod(queries + (size_t)q * dim, queries + (size_t)q * dim, dim, neumaier));
    }
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    int *cnts = (int *)calloc((size_t)nthreads * nq, sizeof(int));
    double *tsc = (double *)malloc((size_t)nthreads * nq * k * sizeof(double));
    int64_t *tid = (int64_t *)malloc((size_t)nthreads * nq * k * sizeof(int64_t));
#pragma omp parallel
    {
        int th = 0;
#ifdef _OPENMP
        th = omp_get_thread_num();
#endif
        float *rowbuf = (float *)malloc((size_t)dim * sizeof(float));
        double dots[ORC_QB];
        int *cnt = cnts + (size_t)th * nq;
        double *sc = tsc + (size_t)th * nq * k;
        int64_t *ids = tid + (size_t)th * nq * k;
#pragma omp for schedule(dynamic, 256)
        for (int64_t r = 0; r < n_rows; ++r) {
            const float *row;
            if (corpus) row = corpus + (size_t)r * dim;
            else { orc_gen_row(seed, (uint64_t)(row_start + r), dim, dup_per_mille, rowbuf); row = rowbuf; }
            const double m2 = sqrt(sum_prod(row, row, dim, neumaier));
            for (int b = 0; b < nqb; ++b) {
                orc_dot_block(row, qt + (size_t)b * dim * ORC_QB, dim, neumaier, dots);
                for (int j = 0; j < ORC_QB; ++j) {
                    const int q = b * ORC_QB + j;
                    if (q >= nq) break;
                    const double s = (m1[q] == 0.0 || m2 == 0.0) ? 0.0 : dots[j] / (m1[q] * m2);
                    orc_topk_insert(s, row_start + r, k, cnt + q, sc + (size_t)q * k, ids + (size_t)q * k);
                }
            }
        }
        free(rowbuf);
    }
    for (int q = 0; q < nq; ++q) {
        int cnt = 0;
        double *sc = out_scores + (size_t)q * k;
        int64_t *ids = out_ids + (size_t)q * k;
        for (int i = 0; i < k; ++i) { sc[i] = 0.0; ids[i] = -1; }
        for (int th = 0; th < nthreads; ++th)
            for (int i = 0; i < cnts[(size_t)th * nq + q]; ++i)
                orc_topk_insert(tsc[((size_t)th * nq + q) * k + i], tid[((size_t)th * nq + q) * k + i], k, &cnt, sc, ids);
    }
    free(qt); free(m1); free(cnts); free(tsc); free(tid);
    return 0;
}

/* ---- token corpus (synthetic.doc_lengths / token_corpus) ---- */
static inline int32_t orc_doc_len(uint64_t seed, uint64_t doc, int lmin, int lmax)
{
    uint64_t h = orc_mix64(seed ^ (doc * ORC_K_DOC));
    return (int32_t)(lmin + (int64_t)(h % (uint64_t)(lmax - lmin + 1)));
}

/* token rank = number of thresholds <= u (numpy searchsorted side='right'), clipped to vocab - 1.
 * `coarse` [4097]: coarse[b] = searchsorted(thr, b << 51, side='right') narrows the binary search. */
static inline int32_t orc_token_of(uint64_t u, const uint64_t *thr, int vocab, const int32_t *coarse)
{
    const uint32_t b = (uint32_t)(u >> 51);
    int lo = coarse[b], hi = coarse[b + 1];  /* answer in [lo, hi] */
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (thr[mid] <= u) lo = mid + 1; else hi = mid;
    }
    return lo < vocab - 1 ? lo : vocab - 1;
}

static int32_t *orc_coarse_table(const uint64_t *thr, int vocab)
{
    int32_t *coarse = (int32_t *)malloc(4098 * sizeof(int32_t));
    int pos = 0;
    for (uint32_t b = 0; b <= 4096; ++b) {
        const uint64_t u = (uint64_t)b << 51;  /* b = 4096 -> 2^63: one past the largest u */
        while (pos < vocab && thr[pos] <= u) ++pos;
        coarse[b] = pos;
    }
    coarse[4097] = vocab;
    return coarse;
}

static inline void orc_gen_doc_tokens(uint64_t seed, uint64_t doc, int len, const uint64_t *thr, int vocab,
                                      const int32_t *coarse, int32_t *out)
{
    const uint64_t key = orc_row_key(seed + 1, doc);
    for (int j = 0; j < len; ++j) out[j] = orc_token_of(orc_mix64(key + (uint64_t)j) >> 1, thr, vocab, coarse);
}

/* synthetic.token_corpus into caller-provided arrays (doc_off [n+1], tokens [total]); tokens may be NULL to get
 * the offsets only.  Returns the total number of tokens. */
int64_t orc_gen_token_corpus(uint64_t seed, int64_t doc_start, int64_t n_docs, int vocab, int lmin, int lmax,
                             const uint64_t *thr, int64_t *doc_off, int32_t *tokens)
{
    doc_off[0] = 0;
    for (int64_t d = 0; d < n_docs; ++d)
        doc_off[d + 1] = doc_off[d] + orc_doc_len(seed, (uint64_t)(doc_start + d), lmin, lmax);
    if (tokens) {
        int32_t *coarse = orc_coarse_table(thr, vocab);
#pragma omp parallel for schedule(dynamic, 1024)
        for (int64_t d = 0; d < n_docs; ++d)
            orc_gen_doc_tokens(seed, (uint64_t)(doc_start + d), (int)(doc_off[d + 1] - doc_off[d]), thr, vocab, coarse,
                               tokens + doc_off[d]);
        free(coarse);
    }
    return doc_off[n_docs];
}

/* BM25._initialize over the regenerated corpus: df[V], first_pos[V] (global token position of the first occurrence,
 * INT64_MAX if unseen: ascending first_pos == dict insertion order of `nd`), total token count. */
int orc_bm25_stream_stats(uint64_t seed, int64_t doc_start, int64_t n_docs, int vocab, int lmin, int lmax,
                          const uint64_t *thr, int64_t *df, int64_t *first_pos, int64_t *total_len)
{
    int32_t *coarse = orc_coarse_table(thr, vocab);
    const int64_t BLK = 4096;
    const int64_t n_blk = (n_docs + BLK - 1) / BLK;
    /* token offset of every block (doc lengths are cheap to recompute) */
    int64_t *blk_off = (int64_t *)malloc((size_t)(n_blk + 1) * sizeof(int64_t));
    blk_off[0] = 0;
    for (int64_t b = 0; b < n_blk; ++b) {
        int64_t s = 0;
        const int64_t d1 = (b + 1) * BLK < n_docs ? (b + 1) * BLK : n_docs;
        for (int64_t d = b * BLK; d < d1; ++d) s += orc_doc_len(seed, (uint64_t)(doc_start + d), lmin, lmax);
        blk_off[b + 1] = blk_off[b] + s;
    }
    *total_len = blk_off[n_blk];
    for (int t = 0; t < vocab; ++t) { df[t] = 0; first_pos[t] = INT64_MAX; }
#pragma omp parallel
    {
        int64_t *ldf = (int64_t *)calloc((size_t)vocab, sizeof(int64_t));
        int64_t *lfirst = (int64_t *)malloc((size_t)vocab * sizeof(int64_t));
        int64_t *last_doc = (int64_t *)malloc((size_t)vocab * sizeof(int64_t));
        int32_t *tok = (int32_t *)malloc((size_t)(lmax > 0 ? lmax : 1) * sizeof(int32_t));
        for (int t = 0; t < vocab; ++t) { lfirst[t] = INT64_MAX; last_doc[t] = -1; }
#pragma omp for schedule(dynamic, 1)
        for (int64_t b = 0; b < n_blk; ++b) {
            int64_t pos = blk_off[b];
            const int64_t d1 = (b + 1) * BLK < n_docs ? (b + 1) * BLK : n_docs;
            for (int64_t d = b * BLK; d < d1; ++d) {
                const int len = orc_doc_len(seed, (uint64_t)(doc_start + d), lmin, lmax);
                orc_gen_doc_tokens(seed, (uint64_t)(doc_start + d), len, thr, vocab, coarse, tok);
                for (int j = 0; j < len; ++j) {
                    const int32_t t = tok[j];
                    if (last_doc[t] != d) { last_doc[t] = d; ldf[t] += 1; }
                    if (pos + j < lfirst[t]) lfirst[t] = pos + j;
                }
                pos += len;
            }
        }
#pragma omp critical
        for (int t = 0; t < vocab; ++t) {
            df[t] += ldf[t];
            if (lfirst[t] < first_pos[t]) first_pos[t] = lfirst[t];
        }
        free(ldf); free(lfirst); free(last_doc); free(tok);
    }
    free(coarse); free(blk_off);
    return 0;
}

/* BM25Okapi._calc_idf from (df, first-seen order): same arithmetic as orc_bm25_build.  order [n_seen] = term ids in
 * dict insertion order.  Writes idf [vocab] (0 for unseen terms); returns eps, *average_idf. */
double orc_bm25_idf_from_df(int64_t n_docs, int vocab, const int64_t *df, const int32_t *order, int n_seen,
                            double *idf, double *average_idf)
{
    for (int t = 0; t < vocab; ++t) idf[t] = 0.0;
    double idf_sum = 0.0;
    for (int i = 0; i < n_seen; ++i) {
        const int32_t t = order[i];
        const double freq = (double)df[t];
        const double v = log((double)n_docs - freq + 0.5) - log(freq + 0.5);
        idf[t] = v;
        idf_sum += v;
    }
    const double avg = n_seen > 0 ? idf_sum / (double)n_seen : 0.0;
    const double eps = 0.25 * avg;
    for (int i = 0; i < n_seen; ++i)
        if (idf[order[i]] < 0.0) idf[order[i]] = eps;
    if (average_idf) *average_idf = avg;
    return eps;
}

/* BM25Okapi.get_scores for `nq` queries over the regenerated docs [doc_start, doc_start + n_docs): raw float64
 * scores raw[q * n_docs + d], same per-document arithmetic and query-token order as orc_bm25_scores_raw.
 * q_terms int32 [nq, lq_max] (entries outside [0, vocab) are OOV), q_lens [nq]; avgdl / idf are the GLOBAL values. */
int orc_bm25_stream_scores(uint64_t seed, int64_t doc_start, int64_t n_docs, int vocab, int lmin, int lmax,
                           const uint64_t *thr, double avgdl, const double *idf, const int32_t *q_terms,
                           const int32_t *q_lens, int nq, int lq_max, double *raw)
{
    const double k1 = 1.5, b = 0.75;
    const double one_minus_b = 1 - b;
    int32_t *coarse = orc_coarse_table(thr, vocab);
    /* slot of every term some query uses */
    int32_t *slot_of = (int32_t *)malloc((size_t)vocab * sizeof(int32_t));
    for (int t = 0; t < vocab; ++t) slot_of[t] = -1;
    int n_slots = 0;
    for (int q = 0; q < nq; ++q)
        for (int i = 0; i < q_lens[q] && i < lq_max; ++i) {
            const int32_t t = q_terms[(size_t)q * lq_max + i];
            if (t >= 0 && t < vocab && slot_of[t] < 0) slot_of[t] = n_slots++;
        }
#pragma omp parallel
    {
        int32_t *tok = (int32_t *)malloc((size_t)(lmax > 0 ? lmax : 1) * sizeof(int32_t));
        int32_t *tf = (int32_t *)calloc((size_t)(n_slots > 0 ? n_slots : 1), sizeof(int32_t));
#pragma omp for schedule(dynamic, 1024)
        for (int64_t d = 0; d < n_docs; ++d) {
            const int len = orc_doc_len(seed, (uint64_t)(doc_start + d), lmin, lmax);
            orc_gen_doc_tokens(seed, (uint64_t)(doc_start + d), len, thr, vocab, coarse, tok);
            for (int s = 0; s < n_slots; ++s) tf[s] = 0;
            for (int j = 0; j < len; ++j) {
                const int32_t s = slot_of[tok[j]];
                if (s >= 0) tf[s] += 1;
            }
            for (int q = 0; q < nq; ++q) {
                double score = 0.0;
                for (int i = 0; i < q_lens[q] && i < lq_max; ++i) {
                    const int32_t t = q_terms[(size_t)q * lq_max + i];
                    if (t < 0 || t >= vocab) continue;
                    const double w = idf[t];
                    if (w == 0.0) continue;
                    const int32_t f = tf[slot_of[t]];
                    if (f == 0) continue;  /* tf == 0 adds +-0.0: a no-op */
                    const double tfd = (double)f;
                    const double t1 = b * (double)len;
                    const double t2 = t1 / avgdl;
                    const double t3 = one_minus_b + t2;
                    const double t4 = k1 * t3;
                    const double den = tfd + t4;
                    const double num = tfd * (k1 + 1);
                    const double r = num / den;
                    const double c = w * r;
                    score = score + c;
                }
                raw[(size_t)q * n_docs + d] = score;
            }
        }
        free(tok); free(tf);
    }
    free(coarse); free(slot_of);
    return 0;
}


/* orc_pairwise_candidates on all cores, 16 partner rows per SIMD block (rag/consistency_checker.py:169-189, 241-261):
 * the same float64 operations per pair (products of widened fp32 values are exact and commute, so is m1 * m2), results
 * in (i, j) lexicographic order.  Returns the number of pairs found (writes at most cap). */
typedef struct { int32_t i, j; double s; } orc_pair_t;
static int orc_pair_cmp(const void *a, const void *b)
{
    const orc_pair_t *x = (const orc_pair_t *)a, *y = (const orc_pair_t *)b;
    if (x->i != y->i) return x->i < y->i ? -1 : 1;
    return x->j < y->j ? -1 : (x->j > y->j ? 1 : 0);
}
int64_t orc_pairwise_candidates_blocked(const float *emb, int64_t m, int d, const int32_t *doc_idx, double thr,
                                        int neumaier, int64_t cap, int32_t *out_i, int32_t *out_j, double *out_sim)
{
    const int64_t nb = (m + ORC_QB - 1) / ORC_QB;
    double *qt = (double *)calloc((size_t)nb * d * ORC_QB, sizeof(double));
    double *mag = (double *)malloc((size_t)(m > 0 ? m : 1) * sizeof(double));
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < m; ++r) {
        for (int i = 0; i < d; ++i) qt[((size_t)(r / ORC_QB) * d + i) * ORC_QB + (r % ORC_QB)] = (double)emb[r * d + i];
        mag[r] = sqrt(sum_prod(emb + r * d, emb + r * d, d, neumaier));
    }
    orc_pair_t *all = NULL;
    int64_t n_all = 0, cap_all = 0;
#pragma omp parallel
    {
        orc_pair_t *mine = NULL;
        int64_t n_mine = 0, cap_mine = 0;
        double dots[ORC_QB];
#pragma omp for schedule(dynamic, 16)
        for (int64_t i = 0; i < m; ++i) {
            for (int64_t b = (i + 1) / ORC_QB; b < nb; ++b) {
                orc_dot_block(emb + i * d, qt + (size_t)b * d * ORC_QB, d, neumaier, dots);
                for (int l = 0; l < ORC_QB; ++l) {
                    const int64_t j = b * ORC_QB + l;
                    if (j <= i || j >= m || doc_idx[i] == doc_idx[j]) continue;
                    const double s = (mag[i] == 0.0 || mag[j] == 0.0) ? 0.0 : dots[l] / (mag[i] * mag[j]);
                    if (s >= thr) {
                        if (n_mine == cap_mine) {
                            cap_mine = cap_mine ? 2 * cap_mine : 1024;
                            mine = (orc_pair_t *)realloc(mine, (size_t)cap_mine * sizeof(orc_pair_t));
                        }
                        mine[n_mine].i = (int32_t)i; mine[n_mine].j = (int32_t)j; mine[n_mine].s = s;
                        ++n_mine;
                    }
                }
            }
        }
#pragma omp critical
        {
            if (n_all + n_mine > cap_all) {
                cap_all = 2 * (n_all + n_mine) + 16;
                all = (orc_pair_t *)realloc(all, (size_t)cap_all * sizeof(orc_pair_t));
            }
            if (n_mine) memcpy(all + n_all, mine, (size_t)n_mine * sizeof(orc_pair_t));
            n_all += n_mine;
        }
        free(mine);
    }
    if (n_all) qsort(all, (size_t)n_all, sizeof(orc_pair_t), orc_pair_cmp);
    for (int64_t t = 0; t < n_all && t < cap; ++t) { out_i[t] = all[t].i; out_j[t] = all[t].j; out_sim[t] = all[t].s; }
    free(all); free(qt); free(mag);
    return n_all;
}

int orc_version(void) { return 1; }
