// bm25_shared.cuh -- pieces common to the two BM25 tile kernels (bm25.cu: exact float64 scatter;
// bm25_ms.cu: fp32 MaxScore first pass + exact re-score): the log-scale score histogram that drives the
// per-query running threshold, the exact per-document scorer, and the final (score desc, id asc) selection
// with the reference's max-normalisation (rag/retrieval.py:343-345) and zero-score fill.
#pragma once
#include "common.cuh"
#include "select.cuh"

namespace orag {
namespace bm25 {

constexpr int kHistBins = 8192;
constexpr int kBinBase = (1023 - 20) << 7;  // bins start at 2^-20, 128 bins per octave (0.54 % wide)

__device__ __forceinline__ int score_bin(double v)
{
    long long e = (__double_as_longlong(v) >> 45) - kBinBase;
    return e < 0 ? 0 : (e > kHistBins - 1 ? kHistBins - 1 : (int)e);
}
__device__ __forceinline__ unsigned long long bin_floor_bits(int b)
{
    return (unsigned long long)(b + kBinBase) << 45;
}

// Re-derives a query's threshold from its histogram `h`, scanning down from bin `hi` (the highest occupied
// bin) in batches of 16 bins fetched with four independent 16-byte loads, and publishes the lower edge of
// the k-th best's bin minus 4 ulps (keeps docs whose NORMALISED score could tie with the k-th best:
// x/m == y/m in float64 only for raw scores a couple of ulps apart).  The published value is always a
// lower bound of the k-th best score emitted so far, so no true top-k doc is ever dropped.
__device__ __forceinline__ void tighten_threshold(const uint32_t *h, int hi, int k, unsigned long long *thr_slot)
{
    const uint4 *h4 = reinterpret_cast<const uint4 *>(h);
    const uint4 zero = make_uint4(0, 0, 0, 0);
    uint32_t acc = 0;
    int found = -1;
    for (int c = hi >> 2; c >= 0 && found < 0; c -= 4) {
        const uint4 w0 = __ldcg(h4 + c);
        const uint4 w1 = c >= 1 ? __ldcg(h4 + c - 1) : zero;
        const uint4 w2 = c >= 2 ? __ldcg(h4 + c - 2) : zero;
        const uint4 w3 = c >= 3 ? __ldcg(h4 + c - 3) : zero;
        const uint32_t vals[16] = {w0.w, w0.z, w0.y, w0.x, w1.w, w1.z, w1.y, w1.x,
                                   w2.w, w2.z, w2.y, w2.x, w3.w, w3.z, w3.y, w3.x};
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            acc += vals[j];
            if (found < 0 && acc >= (uint32_t)k) found = 4 * c + 3 - j;
        }
    }
    if (found >= 1) atomicMax(thr_slot, bin_floor_bits(found) - 4ull);
}

// Contribution of one query token to one doc in the reference's arithmetic: idf * (tf*(k1+1) / (tf + t4[d])) if the
// doc contains the term (binary search in the term's run of the doc's tile), else 0.
__device__ inline double term_contribution(const orag_bm25_index_t &ix, int t, int64_t doc)
{
    if (t < 0 || t >= ix.vocab) return 0.0;
    const double idf = ix.d_idf[t];
    if (idf == 0.0) return 0.0;
    const int T = ix.tile_docs;
    const int tile = (int)(doc / T);
    const uint32_t want = (uint32_t)(doc - (int64_t)tile * T);
    const uint32_t *tile_post = ix.d_postings + ix.d_tile_base[tile];
    const int32_t *toff = ix.d_tile_term_off + (int64_t)tile * (ix.vocab + 1);
    int lo = toff[t], hi = toff[t + 1];
    const int end = hi;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if ((tile_post[mid] >> 16) < want) lo = mid + 1; else hi = mid;
    }
    if (lo < end && (tile_post[lo] >> 16) == want) {
        const double tf = (double)(tile_post[lo] & 0xFFFFu);
        const double t4 = ix.d_t4_table[ix.d_doc_len[doc]];
        return __dmul_rn(idf, __ddiv_rn(__dmul_rn(tf, 2.5), __dadd_rn(tf, t4)));
    }
    return 0.0;
}

// Exact score of one doc for one query in the reference's arithmetic and order: the contributions of the query
// tokens added in query order, duplicates twice (adding the exact 0.0 of an absent token changes nothing).
__device__ inline double score_doc(const orag_bm25_index_t &ix, const int32_t *terms, int nt, int64_t doc)
{
    double s = 0.0;
    for (int i = 0; i < nt; ++i) s = __dadd_rn(s, term_contribution(ix, terms[i], doc));
    return s;
}

// Block-wide: exact top-k of the list (cd[i] local doc, cs[i] raw float64 score), i < n, by
// (score/max desc, id asc), then zero-score fill (docs untouched by the query rank after all positive
// ones, in id order).  Every thread of the CTA must call it; scratch = 32 Picks + 32 doubles.
__device__ inline void select_from_list(const orag_bm25_index_t &ix, const int32_t *q_terms_row, int nt, int k,
                                        const int32_t *cd, const double *cs, uint32_t n, int64_t doc_id_base,
                                        int normalize, int64_t *out_ids, double *out_scores, double *out_max_slot,
                                        Pick *scratch, double *dscratch, bool zero_fill = true)
{
    double mx = -INFINITY;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) mx = fmax(mx, cs[i]);
    mx = block_max(mx, dscratch);
    const double raw_max = mx > 0.0 ? mx : 0.0;
    const double m = mx > 0.0 ? mx : 1.0;
    if (out_max_slot && threadIdx.x == 0) *out_max_slot = normalize ? m : raw_max;
    double prev_s = INFINITY;
    int64_t prev_id = -1;
    int found = 0;
    for (int r = 0; r < k; ++r) {
        Pick best;
        best.valid = 0; best.s = 0.0; best.id = 0;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            const int64_t id = doc_id_base + cd[i];
            const double v = normalize ? __ddiv_rn(cs[i], m) : cs[i];
            if (r > 0 && !ranks_before(prev_s, prev_id, v, id)) continue;
            Pick c;
            c.s = v; c.id = id; c.valid = 1;
            best = better(best, c);
        }
        best = block_best(best, scratch);
        if (!best.valid) break;
        if (threadIdx.x == 0) {
            out_ids[r] = best.id;
            out_scores[r] = best.s;
        }
        prev_s = best.s;
        prev_id = best.id;
        ++found;
    }
    if (found < k && threadIdx.x == 0) {
        // fewer than k docs with a positive score: the rest of the list is zero-score docs in id order
        int r = found;
        for (int64_t d = 0; zero_fill && d < ix.n_docs && r < k; ++d) {
            if (score_doc(ix, q_terms_row, nt, d) == 0.0) {
                out_ids[r] = doc_id_base + d;
                out_scores[r] = 0.0;
                ++r;
            }
        }
        for (; r < k; ++r) {
            out_ids[r] = -1;
            out_scores[r] = 0.0;
        }
    }
}

// bm25_ms.cu
bool ms_eligible(const orag_bm25_index_t *ix, int max_terms, int flags);
size_t ms_workspace_bytes(const orag_bm25_index_t *ix, int n_queries);
int ms_topk(const orag_bm25_index_t *ix, int64_t doc_id_base, const int32_t *d_query_terms,
            const int32_t *d_query_lens, int n_queries, int max_terms, int k, int normalize, bool background,
            int64_t *d_out_ids, double *d_out_scores, double *d_out_max, int32_t *d_out_status, void *d_workspace,
            cudaStream_t st);

}  // namespace bm25
}  // namespace orag
