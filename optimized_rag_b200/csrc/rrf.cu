// rrf.cu -- Reciprocal Rank Fusion, one thread per query.
//
// Restates ReciprocalRankFusion.fuse (rag/reranker.py:224-271): lists are walked in order, items
// in rank order; c = 1/(k + rank) in float64 (one IEEE division of 1.0 by the exactly converted
// integer); first sighting stores, later sightings add; the result is the stable descending sort
// of the union in first-sighting order (dict insertion order), truncated to top_k.  Keys are chunk
// ids (the reference keys on the content string; equivalent for unique contents).
// The union is at most n_lists * list_len <= 128 entries, so a register/local-memory insertion
// sort per thread is the right size; the kernel is latency-bound and fused after the final merge.
#include "common.cuh"

namespace orag {

constexpr int kRrfMaxUnion = 128;
constexpr int kRrfMaxLists = 8;

__global__ void __launch_bounds__(128) rrf_kernel(const int64_t *__restrict__ list_ids, int n_queries, int n_lists,
                                                 int list_len, int rrf_k, int top_k, int tie_mode,
                                                 int64_t *__restrict__ out_ids, double *__restrict__ out_scores,
                                                 int32_t *__restrict__ out_src)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_queries) return;
    int64_t keys[kRrfMaxUnion];
    double sc[kRrfMaxUnion];
    uint8_t src[kRrfMaxUnion][kRrfMaxLists];
    uint8_t ord[kRrfMaxUnion];
    int m = 0;
    const int64_t *base = list_ids + (int64_t)q * n_lists * list_len;
    for (int l = 0; l < n_lists; ++l) {
        for (int r = 0; r < list_len; ++r) {
            const int64_t id = base[l * list_len + r];
            if (id < 0) break;  // tail padding
            const double c = __ddiv_rn(1.0, (double)(rrf_k + r + 1));
            int j = 0;
            for (; j < m; ++j)
                if (keys[j] == id) break;
            if (j < m) {
                sc[j] = __dadd_rn(sc[j], c);
                if (src[j][l] == 0) src[j][l] = (uint8_t)(r + 1);
            } else {
                keys[m] = id;
                sc[m] = c;
                for (int t = 0; t < kRrfMaxLists; ++t) src[m][t] = 0;
                src[m][l] = (uint8_t)(r + 1);
                ++m;
            }
        }
    }
    // stable insertion sort of first-sighting order, descending score
    for (int i = 0; i < m; ++i) {
        int p = i;
        while (p > 0) {
            const int o = ord[p - 1];
            const bool before = (sc[i] > sc[o]) || (tie_mode == 1 && sc[i] == sc[o] && keys[i] < keys[o]);
            if (!before) break;
            ord[p] = (uint8_t)o;
            --p;
        }
        ord[p] = (uint8_t)i;
    }
    for (int r = 0; r < top_k; ++r) {
        const bool have = r < m;
        const int e = have ? ord[r] : 0;
        out_ids[(int64_t)q * top_k + r] = have ? keys[e] : -1;
        out_scores[(int64_t)q * top_k + r] = have ? sc[e] : 0.0;
        if (out_src)
            for (int l = 0; l < n_lists; ++l) out_src[((int64_t)q * top_k + r) * n_lists + l] = have ? src[e][l] : 0;
    }
}

}  // namespace orag

extern "C" int orag_rrf_fuse(const int64_t *d_list_ids, int n_queries, int n_lists, int list_len, int rrf_k, int top_k,
                             int tie_mode, int64_t *d_out_ids, double *d_out_scores, int32_t *d_out_src, void *stream)
{
    ORAG_REQUIRE(d_list_ids && d_out_ids && d_out_scores, "rrf_fuse pointers");
    ORAG_REQUIRE(n_queries >= 0 && n_lists >= 1 && n_lists <= orag::kRrfMaxLists && list_len >= 1 && top_k >= 1,
                 "rrf_fuse sizes");
    ORAG_REQUIRE(n_lists * list_len <= orag::kRrfMaxUnion, "n_lists * list_len <= 128");
    ORAG_REQUIRE(list_len <= 255 && rrf_k >= 0, "list_len <= 255");
    if (n_queries == 0) return ORAG_OK;
    orag::rrf_kernel<<<(n_queries + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        d_list_ids, n_queries, n_lists, list_len, rrf_k, top_k, tie_mode, d_out_ids, d_out_scores, d_out_src);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}

// ------------------------------------------------------------------------------------------------
// Weighted hybrid score of HybridRetriever.hybrid_search (rag/retrieval.py:302):
//   h = alpha*sem + beta*kw + gamma*temp        Python float64, left to right, no contraction
namespace orag {
__global__ void weighted_sum3_kernel(const double *__restrict__ a, const double *__restrict__ b,
                                     const double *__restrict__ c, int64_t n, double alpha, double beta, double gamma,
                                     double *__restrict__ out)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const double t = c ? c[i] : 0.0;
        out[i] = __dadd_rn(__dadd_rn(__dmul_rn(alpha, a[i]), __dmul_rn(beta, b[i])), __dmul_rn(gamma, t));
    }
}
}  // namespace orag

extern "C" int orag_weighted_sum3(const double *d_sem, const double *d_kw, const double *d_temp, int64_t n, double alpha,
                                  double beta, double gamma, double *d_out, void *stream)
{
    ORAG_REQUIRE(d_sem && d_kw && d_out && n >= 0, "weighted_sum3");
    if (n == 0) return ORAG_OK;
    int64_t blocks = (n + 255) / 256;
    int64_t lim = (int64_t)orag::sm_count() * 16;
    if (blocks > lim) blocks = lim;
    orag::weighted_sum3_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_sem, d_kw, d_temp, n, alpha, beta,
                                                                                   gamma, d_out);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}

// out[i] = in[i] / divisor (one IEEE division per element: `s / max_score`, rag/retrieval.py:345)
namespace orag {
__global__ void div_scalar_kernel(const double *__restrict__ in, int64_t n, double divisor, double *__restrict__ out)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = __ddiv_rn(in[i], divisor);
}
}  // namespace orag

extern "C" int orag_div_scalar(const double *d_in, int64_t n, double divisor, double *d_out, void *stream)
{
    ORAG_REQUIRE(d_in && d_out && n >= 0, "div_scalar");
    if (n == 0) return ORAG_OK;
    int64_t blocks = (n + 255) / 256;
    int64_t lim = (int64_t)orag::sm_count() * 16;
    if (blocks > lim) blocks = lim;
    orag::div_scalar_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_in, n, divisor, d_out);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}
