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

int orc_version(void) { return 1; }
