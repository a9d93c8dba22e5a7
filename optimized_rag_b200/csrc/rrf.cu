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

// Fuses the lists of ONE query.  base[l * list_stride + r] = id at rank r+1 of list l (id < 0 = tail padding).
__device__ inline void rrf_one(const int64_t *base, int n_lists, int list_len, int list_stride, int rrf_k, int top_k,
                               int tie_mode, int64_t *out_ids, double *out_scores, int32_t *out_src)
{
    int64_t keys[kRrfMaxUnion];
    double sc[kRrfMaxUnion];
    uint8_t src[kRrfMaxUnion][kRrfMaxLists];
    uint8_t ord[kRrfMaxUnion];
    int m = 0;
    for (int l = 0; l < n_lists; ++l) {
        for (int r = 0; r < list_len; ++r) {
            const int64_t id = base[l * list_stride + r];
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
        out_ids[r] = have ? keys[e] : -1;
        out_scores[r] = have ? sc[e] : 0.0;
        if (out_src)
            for (int l = 0; l < n_lists; ++l) out_src[r * n_lists + l] = have ? src[e][l] : 0;
    }
}

__global__ void __launch_bounds__(128) rrf_kernel(const int64_t *__restrict__ list_ids, int n_queries, int n_lists,
                                                 int list_len, int rrf_k, int top_k, int tie_mode,
                                                 int64_t *__restrict__ out_ids, double *__restrict__ out_scores,
                                                 int32_t *__restrict__ out_src)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_queries) return;
    rrf_one(list_ids + (int64_t)q * n_lists * list_len, n_lists, list_len, list_len, rrf_k, top_k, tie_mode,
            out_ids + (int64_t)q * top_k, out_scores + (int64_t)q * top_k,
            out_src ? out_src + (int64_t)q * top_k * n_lists : nullptr);
}

// The one-shard hybrid step: the two ranked lists live in two separate arrays (the outputs of orag_cosine_topk and
// orag_bm25_topk); fuse them in place and OR the two per-query status words -- no packing kernel in between.
__global__ void __launch_bounds__(128) rrf_pair_kernel(const int64_t *__restrict__ ids_a, const int64_t *__restrict__ ids_b,
                                                      int n_queries, int list_len, int rrf_k, int top_k, int tie_mode,
                                                      const int32_t *__restrict__ status_a,
                                                      const int32_t *__restrict__ status_b,
                                                      int64_t *__restrict__ out_ids, double *__restrict__ out_scores,
                                                      int32_t *__restrict__ out_src, int32_t *__restrict__ out_status)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_queries) return;
    int64_t both[kRrfMaxUnion];
    for (int r = 0; r < list_len; ++r) {
        both[r] = ids_a[(int64_t)q * list_len + r];
        both[list_len + r] = ids_b[(int64_t)q * list_len + r];
    }
    rrf_one(both, 2, list_len, list_len, rrf_k, top_k, tie_mode, out_ids + (int64_t)q * top_k,
            out_scores + (int64_t)q * top_k, out_src ? out_src + (int64_t)q * top_k * 2 : nullptr);
    if (out_status) out_status[q] = (status_a ? status_a[q] : 0) | (status_b ? status_b[q] : 0);
}

// ------------------------------------------------------------------------------------------------
// Everything that follows the all-gather of a row-sharded hybrid search, in ONE launch (one 64-thread
// CTA per query): merge the G shards' cosine winners by (score desc, id asc); divide the shards' raw
// BM25 winners by the GLOBAL max raw score (rag/retrieval.py:343-345) and merge them the same way;
// fuse the two lists with RRF.  The gathered buffer is [G, B, W] int64 with
// W = 2*fetch_k + 2*kk + 2: cosine ids | cosine score bits | BM25 ids | BM25 raw score bits |
// shard max raw BM25 score bits | status bits (optimized_rag_b200/dist.py pack_local).
constexpr int kMergeMax = 256;  // candidates of one kind per query (G * kk)

__global__ void __launch_bounds__(64) hybrid_merge_kernel(const int64_t *__restrict__ gathered, int G, int B, int fetch_k,
                                                         int kk, int rrf_k, int top_k, int tie_mode,
                                                         int64_t *__restrict__ out_ids, double *__restrict__ out_scores,
                                                         int32_t *__restrict__ out_src, int64_t *__restrict__ cos_ids,
                                                         double *__restrict__ cos_scores, int64_t *__restrict__ bm_ids,
                                                         double *__restrict__ bm_scores, double *__restrict__ bm_max,
                                                         int32_t *__restrict__ out_status)
{
    __shared__ int64_t c_id[kMergeMax];
    __shared__ double c_sc[kMergeMax];
    __shared__ int64_t lists[2 * 128];  // [2][fetch_k] merged lists for the RRF step
    __shared__ double l_sc[2 * 128];
    __shared__ double s_max;
    __shared__ int s_status;
    const int q = blockIdx.x;
    const int W = 2 * fetch_k + 2 * kk + 2;
    if (threadIdx.x == 0) {
        double mx = -INFINITY;
        int st = 0;
        for (int g = 0; g < G; ++g) {
            const int64_t *row = gathered + ((int64_t)g * B + q) * W;
            mx = fmax(mx, __longlong_as_double(row[2 * fetch_k + 2 * kk]));
            st |= (int)row[2 * fetch_k + 2 * kk + 1];
        }
        s_max = mx > 0.0 ? mx : 1.0;
        s_status = st;
    }
    for (int kind = 0; kind < 2; ++kind) {
        const int per = kind == 0 ? fetch_k : kk;
        const int off = kind == 0 ? 0 : 2 * fetch_k;
        const int n = G * per;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int g = i / per, r = i - g * per;
            const int64_t *row = gathered + ((int64_t)g * B + q) * W + off;
            c_id[i] = row[r];
            const double v = __longlong_as_double(row[per + r]);
            c_sc[i] = kind == 0 ? v : __ddiv_rn(v, s_max);
        }
        for (int i = threadIdx.x; i < fetch_k; i += blockDim.x) {
            lists[kind * fetch_k + i] = -1;
            l_sc[kind * fetch_k + i] = 0.0;
        }
        __syncthreads();
        // rank of every candidate = number of candidates that rank strictly before it (ids are unique)
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int64_t id = c_id[i];
            if (id < 0) continue;
            const double v = c_sc[i];
            int rank = 0;
            for (int j = 0; j < n; ++j)
                if (c_id[j] >= 0 && ranks_before(c_sc[j], c_id[j], v, id)) ++rank;
            if (rank < fetch_k) {
                lists[kind * fetch_k + rank] = id;
                l_sc[kind * fetch_k + rank] = v;
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < fetch_k; i += blockDim.x) {
        cos_ids[(int64_t)q * fetch_k + i] = lists[i];
        cos_scores[(int64_t)q * fetch_k + i] = l_sc[i];
        bm_ids[(int64_t)q * fetch_k + i] = lists[fetch_k + i];
        bm_scores[(int64_t)q * fetch_k + i] = l_sc[fetch_k + i];
    }
    if (threadIdx.x == 0) {
        // Truncation check of the BM25 lists.  Shards rank by RAW score (the divisor is only known after the
        // exchange); x -> x / max can map two or three neighbouring raw values to the same normalised double, and
        // the global order inside such a group is by id.  A shard whose FULL list ends inside the group of the
        // k-th normalised value may have cut off a member with a lower id than one it kept -- possible only if the
        // group mixes raw values (or a raw value just below the last kept one still lands in it).  Equal raw values
        // are harmless: the shard kept its lowest ids, and fewer than kk are ever needed from one shard.  Flag such
        // queries; the caller repairs them through the exhaustive path (max first, then ranking).
        const int64_t kth_id = lists[fetch_k + fetch_k - 1];
        if (kth_id >= 0) {
            const double vk = l_sc[fetch_k + fetch_k - 1];
            for (int g = 0; g < G; ++g) {
                const int64_t *row = gathered + ((int64_t)g * B + q) * W + 2 * fetch_k;
                if (row[kk - 1] < 0) continue;  // shard list not full: nothing was cut off
                const double raw_last = __longlong_as_double(row[kk + kk - 1]);
                if (__ddiv_rn(raw_last, s_max) != vk) continue;
                bool mixed = raw_last > 0.0 &&
                             __ddiv_rn(__longlong_as_double(__double_as_longlong(raw_last) - 1), s_max) == vk;
                for (int r = 0; r < kk - 1 && !mixed; ++r) {
                    const double raw = __longlong_as_double(row[kk + r]);
                    mixed = raw != raw_last && __ddiv_rn(raw, s_max) == vk;
                }
                if (mixed) s_status |= ORAG_STATUS_OVERFLOW;
            }
        }
        bm_max[q] = s_max;
        if (out_status) out_status[q] = s_status;
        rrf_one(lists, 2, fetch_k, fetch_k, rrf_k, top_k, tie_mode, out_ids + (int64_t)q * top_k,
                out_scores + (int64_t)q * top_k, out_src ? out_src + (int64_t)q * top_k * 2 : nullptr);
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

extern "C" int orag_rrf_fuse_pair(const int64_t *d_ids_a, const int64_t *d_ids_b, int n_queries, int list_len, int rrf_k,
                                  int top_k, int tie_mode, const int32_t *d_status_a, const int32_t *d_status_b,
                                  int64_t *d_out_ids, double *d_out_scores, int32_t *d_out_src, int32_t *d_out_status,
                                  void *stream)
{
    ORAG_REQUIRE(d_ids_a && d_ids_b && d_out_ids && d_out_scores, "rrf_fuse_pair pointers");
    ORAG_REQUIRE(n_queries >= 0 && list_len >= 1 && 2 * list_len <= orag::kRrfMaxUnion && top_k >= 1 && rrf_k >= 0,
                 "rrf_fuse_pair sizes (2 * list_len <= 128)");
    if (n_queries == 0) return ORAG_OK;
    orag::TimelineScope tl(orag::TL_RRF, (cudaStream_t)stream);
    orag::rrf_pair_kernel<<<(n_queries + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        d_ids_a, d_ids_b, n_queries, list_len, rrf_k, top_k, tie_mode, d_status_a, d_status_b, d_out_ids, d_out_scores,
        d_out_src, d_out_status);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}

extern "C" int orag_hybrid_merge(const int64_t *d_gathered, int n_shards, int n_queries, int fetch_k, int kk, int rrf_k,
                                 int top_k, int tie_mode, int64_t *d_out_ids, double *d_out_scores, int32_t *d_out_src,
                                 int64_t *d_cos_ids, double *d_cos_scores, int64_t *d_bm25_ids, double *d_bm25_scores,
                                 double *d_bm25_max, int32_t *d_out_status, void *stream)
{
    ORAG_REQUIRE(d_gathered && d_out_ids && d_out_scores && d_cos_ids && d_cos_scores && d_bm25_ids && d_bm25_scores &&
                     d_bm25_max,
                 "hybrid_merge pointers");
    ORAG_REQUIRE(n_shards >= 1 && n_queries >= 0 && fetch_k >= 1 && kk >= fetch_k && top_k >= 1 && rrf_k >= 0,
                 "hybrid_merge sizes");
    ORAG_REQUIRE(n_shards * kk <= orag::kMergeMax && fetch_k <= 64, "n_shards * kk <= 256 and fetch_k <= 64");
    if (n_queries == 0) return ORAG_OK;
    orag::TimelineScope tl(orag::TL_MERGE, (cudaStream_t)stream);
    orag::hybrid_merge_kernel<<<n_queries, 64, 0, (cudaStream_t)stream>>>(
        d_gathered, n_shards, n_queries, fetch_k, kk, rrf_k, top_k, tie_mode, d_out_ids, d_out_scores, d_out_src,
        d_cos_ids, d_cos_scores, d_bm25_ids, d_bm25_scores, d_bm25_max, d_out_status);
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
