// pairwise.cu -- all-pairs cosine candidates for the consistency checker
// (rag/consistency_checker.py:169-189): every i < j with doc_idx[i] != doc_idx[j] and
// float64 cosine >= threshold.  Exact CUDA-core version: lane <-> row j, a block of 4 "query"
// rows i per pass, same float64 arithmetic as the retrieval cosine (exact.cuh).
#include "common.cuh"
#include "exact.cuh"

namespace orag {

constexpr int kPairQB = 4;

// sq[r] = Neumaier sum of emb[r]^2
__global__ void __launch_bounds__(256) pair_sq_kernel(const float *__restrict__ emb, int64_t m, int dim,
                                                     double *__restrict__ sq)
{
    __shared__ float stage_all[8][32 * 33];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int64_t warp = blockIdx.x * (int64_t)(blockDim.x >> 5) + wib;
    int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t rb = warp; rb < (m + 31) / 32; rb += n_warps) {
        int64_t r = rb * 32 + lane;
        const float *rowptr = r < m ? emb + r * (int64_t)dim : nullptr;
        const float *qp[1] = {nullptr};
        NeuSum dot[1], s;
        warp_score_rows<1>(rowptr, qp, dim, stage_all[wib], dot, s, true);
        if (r < m) sq[r] = s.result();
    }
}

__global__ void __launch_bounds__(256) pairwise_kernel(const float *__restrict__ emb, int64_t m, int dim,
                                                      const int32_t *__restrict__ doc_idx, const double *__restrict__ sq,
                                                      double thr, int64_t cap, int32_t *__restrict__ out_i,
                                                      int32_t *__restrict__ out_j, double *__restrict__ out_sim,
                                                      unsigned long long *__restrict__ out_count)
{
    __shared__ float stage_all[8][32 * 33];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t n_jb = (m + 31) / 32;
    const int64_t n_ib = (m + kPairQB - 1) / kPairQB;
    int64_t warp = blockIdx.x * (int64_t)(blockDim.x >> 5) + wib;
    int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t w = warp; w < n_jb * n_ib; w += n_warps) {
        const int64_t jb = w % n_jb, ib = w / n_jb;
        const int64_t i0 = ib * kPairQB;
        if (jb * 32 + 31 <= i0) continue;  // whole j block is <= every i of the group: no i < j pair
        const int64_t j = jb * 32 + lane;
        const float *rowptr = j < m ? emb + j * (int64_t)dim : nullptr;
        const float *qp[kPairQB];
#pragma unroll
        for (int b = 0; b < kPairQB; ++b) qp[b] = (i0 + b < m) ? emb + (i0 + b) * (int64_t)dim : nullptr;
        NeuSum dot[kPairQB], s;
        warp_score_rows<kPairQB>(rowptr, qp, dim, stage_all[wib], dot, s, false);
        if (j >= m) continue;
#pragma unroll
        for (int b = 0; b < kPairQB; ++b) {
            const int64_t i = i0 + b;
            if (i >= m || i >= j || doc_idx[i] == doc_idx[j]) continue;
            // argument order of the reference: cosine(emb[i], emb[j]) -> products emb[i][c] * emb[j][c]
            const double c = cosine_from_sums(dot[b].result(), sq[i], sq[j]);
            if (c >= thr) {
                unsigned long long slot = atomicAdd(out_count, 1ull);
                if ((int64_t)slot < cap) {
                    out_i[slot] = (int32_t)i;
                    out_j[slot] = (int32_t)j;
                    out_sim[slot] = c;
                }
            }
        }
    }
}

}  // namespace orag

extern "C" int orag_pairwise_cosine_threshold(const float *d_emb, int64_t m, int dim, const int32_t *d_doc_idx,
                                              double threshold, int64_t cap, int32_t *d_out_i, int32_t *d_out_j,
                                              double *d_out_sim, unsigned long long *d_out_count, void *d_workspace,
                                              size_t workspace_bytes, void *stream)
{
    ORAG_REQUIRE(d_emb && d_doc_idx && d_out_i && d_out_j && d_out_sim && d_out_count && m >= 0 && dim > 0 && cap >= 0,
                 "pairwise");
    ORAG_REQUIRE(m < ((int64_t)1 << 31), "m must fit int32");
    if (workspace_bytes < orag_pairwise_workspace_bytes(m, dim) || !d_workspace) {
        orag::set_error("pairwise: workspace too small");
        return ORAG_EWORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    ORAG_CUDA_CHECK(cudaMemsetAsync(d_out_count, 0, sizeof(unsigned long long), st));
    if (m < 2) return ORAG_OK;
    double *sq = (double *)d_workspace;
    int64_t blocks = ((m + 31) / 32 + 7) / 8;
    orag::pair_sq_kernel<<<(unsigned)(blocks < 1 ? 1 : blocks), 256, 0, st>>>(d_emb, m, dim, sq);
    ORAG_LAUNCH_CHECK();
    int64_t warps = ((m + 31) / 32) * ((m + orag::kPairQB - 1) / orag::kPairQB);
    int64_t grid = (warps + 7) / 8;
    int64_t lim = (int64_t)orag::sm_count() * 8;
    if (grid > lim) grid = lim;
    orag::pairwise_kernel<<<(unsigned)grid, 256, 0, st>>>(d_emb, m, dim, d_doc_idx, sq, threshold, cap, d_out_i, d_out_j,
                                                          d_out_sim, d_out_count);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}

// ================================================================================================
// Tensor-core path (BASELINE config 5: 64k claims x 1536): the similarity scan of cosine_tc.cu with a FIXED
// threshold, in pair mode: ONE launch walks the triangular space of (256-row query block j, 128-row tile i below
// the block's end).  Operands: the fp16 shadow of the embeddings, every row scaled by a power of two
// (orag_f32_to_f16_rows; kind::f16, CTA pairs with one tcgen05.mma.cta_group::2 per K step exactly like the
// retrieval scan) when dim % 64 == 0, else tf32 straight off the fp32 rows.  The threshold compare runs on the
// accumulators as they leave TMEM; every (i < j) whose first-pass cosine clears threshold - eps(mode, dim) is
// re-scored in the reference's float64 arithmetic and kept if doc_idx differ and the exact cosine >= threshold.
// |first pass - exact| <= eps, so no pair is lost.
//
// Ingest and search are separate calls: orag_pairwise_prepare derives, once per claim matrix, the fp16 shadow,
// the first-pass norms and the float64 sum(a*a) of every row; orag_pairwise_pairs runs the search over them.
#include "cosine_tc.cuh"

extern "C" int orag_f32_to_f16_rows(const float *d_src, int64_t n_rows, int dim, void *d_dst_f16,
                                    float *d_inv_norm_scaled, float *d_scale, void *stream);
extern "C" int orag_row_inv_norms(const float *d_corpus, int64_t n_rows, int dim, float *d_inv_norm, void *stream);

namespace orag {

constexpr int kPairCap = 64;  // first-pass candidates (i < j) per row j

// qnorm[r] = 1 / inv_norm[r]: the row's magnitude in the units of the first-pass accumulators (|x| * scale for the
// fp16 shadow, |x| for tf32); 0 for an all-zero row
__global__ void pair_qnorm_kernel(const float *__restrict__ inv_norm, int64_t n, float *__restrict__ qnorm)
{
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r < n) qnorm[r] = inv_norm[r] > 0.f ? 1.f / inv_norm[r] : 0.f;
}

__global__ void pair_thr_init_kernel(const float *__restrict__ qnorm, int64_t n, float thr,
                                     uint32_t *__restrict__ thr_key, uint32_t *__restrict__ cnt)
{
    int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= n) return;
    // a zero-norm row has cosine 0.0 with everything: it can only pair up when thr <= 0 (handled by the
    // exact path); here it must not flood the candidate lists
    thr_key[q] = float_to_ordered(qnorm[q] > 0.f ? thr : INFINITY);
    cnt[q] = 0;
}

// one warp per row j: keep the re-scored candidates that really clear the threshold
__global__ void __launch_bounds__(256) pair_filter_kernel(const double *__restrict__ scores, const int64_t *__restrict__ ids,
                                                         const uint32_t *__restrict__ cnt, int cap, int64_t m,
                                                         const int32_t *__restrict__ doc_idx, double thr, int64_t out_cap,
                                                         int32_t *__restrict__ out_i, int32_t *__restrict__ out_j,
                                                         double *__restrict__ out_sim,
                                                         unsigned long long *__restrict__ out_count)
{
    const int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = warp; j < m; j += n_warps) {
        uint32_t n = cnt[j];
        if (n > (uint32_t)cap) {
            if (lane == 0) atomicExch(out_count + 1, 1ull);  // overflow flag: caller falls back to the exact path
            n = cap;
        }
        for (uint32_t s = lane; s < n; s += 32) {
            const int64_t i = ids[j * cap + s];
            const double c = scores[j * cap + s];
            if (i < j && doc_idx[i] != doc_idx[j] && c >= thr) {
                unsigned long long slot = atomicAdd(out_count, 1ull);
                if ((int64_t)slot < out_cap) {
                    out_i[slot] = (int32_t)i;
                    out_j[slot] = (int32_t)j;
                    out_sim[slot] = c;
                }
            }
        }
    }
}

static inline bool pair_f16(int dim)
{
    static const int force_tf32 = getenv("ORAG_PAIR_TF32") ? atoi(getenv("ORAG_PAIR_TF32")) : 0;
    return dim % 64 == 0 && !force_tf32;
}

struct PairPrep {   // per claim matrix (orag_pairwise_prepare)
    double *sq;      // [m] float64 sum(a*a), reference arithmetic
    float *inv_norm; // [m] first-pass 1/|row| (fp16 mode: divided by the row's power-of-two scale as well)
    float *qnorm;    // [m] 1 / inv_norm
    void *shadow;    // [m, dim] fp16, rows scaled (fp16 mode only)
    size_t bytes;
};

static PairPrep carve_prep(void *base, int64_t m, int dim)
{
    PairPrep w{};
    uint8_t *p = (uint8_t *)base;
    auto take = [&](size_t n) {
        uint8_t *r = p;
        p += align_up(n, 256);
        return r;
    };
    const size_t mm = (size_t)(m > 0 ? m : 1);
    w.sq = (double *)take(mm * 8);
    w.inv_norm = (float *)take(mm * 4);
    w.qnorm = (float *)take(mm * 4);
    w.shadow = pair_f16(dim) ? take(mm * (size_t)dim * 2) : nullptr;
    w.bytes = (size_t)(p - (uint8_t *)base);
    return w;
}

struct PairWs {     // per search (orag_pairwise_pairs)
    uint32_t *thr_key, *cnt;
    int32_t *cand;
    double *scores;
    int64_t *ids;
    size_t bytes;
};

static PairWs carve_pair(void *base, int64_t m)
{
    PairWs w{};
    uint8_t *p = (uint8_t *)base;
    auto take = [&](size_t n) {
        uint8_t *r = p;
        p += align_up(n, 256);
        return r;
    };
    const size_t mm = (size_t)(m > 0 ? m : 1);
    w.thr_key = (uint32_t *)take(mm * 4);
    w.cnt = (uint32_t *)take(mm * 4);
    w.cand = (int32_t *)take(mm * kPairCap * 4);
    w.scores = (double *)take(mm * kPairCap * 8);
    w.ids = (int64_t *)take(mm * kPairCap * 8);
    w.bytes = (size_t)(p - (uint8_t *)base);
    return w;
}

}  // namespace orag

extern "C" size_t orag_pairwise_prepared_bytes(int64_t m, int dim) { return orag::carve_prep(nullptr, m, dim).bytes; }
extern "C" size_t orag_pairwise_pairs_workspace_bytes(int64_t m) { return orag::carve_pair(nullptr, m).bytes; }

extern "C" int orag_pairwise_prepare(const float *d_emb, int64_t m, int dim, void *d_prepared, size_t prepared_bytes,
                                     void *stream)
{
    using namespace orag;
    ORAG_REQUIRE(d_emb && d_prepared && m >= 0 && dim > 0 && dim % 32 == 0 && m < ((int64_t)1 << 31), "pairwise_prepare");
    ORAG_REQUIRE((reinterpret_cast<uintptr_t>(d_emb) & 15) == 0, "16-byte alignment");
    if (prepared_bytes < orag_pairwise_prepared_bytes(m, dim)) {
        set_error("pairwise_prepare: buffer too small");
        return ORAG_EWORKSPACE;
    }
    if (m == 0) return ORAG_OK;
    cudaStream_t st = (cudaStream_t)stream;
    PairPrep w = carve_prep(d_prepared, m, dim);
    int64_t blocks = ((m + 31) / 32 + 7) / 8;
    pair_sq_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_emb, m, dim, w.sq);
    ORAG_LAUNCH_CHECK();
    int rc = w.shadow ? orag_f32_to_f16_rows(d_emb, m, dim, w.shadow, w.inv_norm, nullptr, st)
                      : orag_row_inv_norms(d_emb, m, dim, w.inv_norm, st);
    if (rc) return rc;
    pair_qnorm_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(w.inv_norm, m, w.qnorm);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}

extern "C" int orag_pairwise_pairs(const float *d_emb, const void *d_prepared, int64_t m, int dim,
                                   const int32_t *d_doc_idx, double threshold, int64_t cap, int32_t *d_out_i,
                                   int32_t *d_out_j, double *d_out_sim, unsigned long long *d_out_count,
                                   void *d_workspace, size_t workspace_bytes, void *stream)
{
    using namespace orag;
    ORAG_REQUIRE(d_emb && d_prepared && d_doc_idx && d_out_i && d_out_j && d_out_sim && d_out_count && m >= 0 && dim > 0 &&
                     cap >= 0,
                 "pairwise_pairs");
    ORAG_REQUIRE(dim % 32 == 0 && m < ((int64_t)1 << 31), "dim % 32 == 0");
    ORAG_REQUIRE(m <= 256 * 1024, "pairwise_pairs: at most 262144 rows (triangular tile index)");
    const bool f16 = pair_f16(dim);
    const float eps = tc::first_pass_eps(f16 ? ORAG_COS_F16 : ORAG_COS_TF32, dim);
    ORAG_REQUIRE(threshold > (double)eps, "tensor-core path needs threshold > first-pass error bound");
    if (workspace_bytes < orag_pairwise_pairs_workspace_bytes(m) || !d_workspace) {
        set_error("pairwise_pairs: workspace too small");
        return ORAG_EWORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    ORAG_CUDA_CHECK(cudaMemsetAsync(d_out_count, 0, 2 * sizeof(unsigned long long), st));
    if (m < 2) return ORAG_OK;
    PairPrep pre = carve_prep(const_cast<void *>(d_prepared), m, dim);
    PairWs w = carve_pair(d_workspace, m);
    pair_thr_init_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(pre.qnorm, m, (float)threshold, w.thr_key, w.cnt);
    ORAG_LAUNCH_CHECK();
    // ONE tensor-core launch over the triangular (query block, row tile) space
    tc::ScanParams p{};
    p.f16 = f16 ? 1 : 0;
    p.n_queries = (int)m;
    p.inv_norm = pre.inv_norm;
    p.thr_key = w.thr_key;
    p.cnt = w.cnt;
    p.hist = nullptr;
    p.cand = w.cand;
    p.cap = kPairCap;
    p.qnorm = pre.qnorm;
    p.inv_qnorm = nullptr;
    p.margin = eps + 1e-6f;  // one-sided: cos >= t  =>  first pass >= t - eps (+ float(threshold) rounding)
    p.k = 0x7fffffff;
    p.fixed_thr = 1;
    p.pair_mode = 1;
    p.row_begin = 0;
    p.row_end = m;
    p.dense = 0;
    const void *op = f16 ? (const void *)pre.shadow : (const void *)d_emb;
    int rc = tc::launch_scan(f16, op, m, op, dim, p, st);
    if (rc) return rc;
    // float64 re-score of every surviving (i, j) in the reference's arithmetic, then the exact filter
    rc = launch_rescore(d_emb, dim, 0, d_emb, pre.sq, w.cand, w.cnt, kPairCap, (int)m, pre.sq, w.scores, w.ids, st);
    if (rc) return rc;
    pair_filter_kernel<<<sm_count() * 8, 256, 0, st>>>(w.scores, w.ids, w.cnt, kPairCap, m, d_doc_idx, threshold, cap,
                                                       d_out_i, d_out_j, d_out_sim, d_out_count);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}

// One-call form: prepare + pairs out of a single workspace.
extern "C" size_t orag_pairwise_tc_workspace_bytes(int64_t m, int dim)
{
    return orag_pairwise_prepared_bytes(m, dim) + orag_pairwise_pairs_workspace_bytes(m);
}

extern "C" int orag_pairwise_cosine_threshold_tc(const float *d_emb, int64_t m, int dim, const int32_t *d_doc_idx,
                                                 double threshold, int64_t cap, int32_t *d_out_i, int32_t *d_out_j,
                                                 double *d_out_sim, unsigned long long *d_out_count, void *d_workspace,
                                                 size_t workspace_bytes, void *stream)
{
    using namespace orag;
    ORAG_REQUIRE(d_workspace && dim > 0 && m >= 0, "pairwise_tc");
    if (workspace_bytes < orag_pairwise_tc_workspace_bytes(m, dim)) {
        set_error("pairwise_tc: workspace too small");
        return ORAG_EWORKSPACE;
    }
    const size_t pb = orag_pairwise_prepared_bytes(m, dim);
    int rc = orag_pairwise_prepare(d_emb, m, dim, d_workspace, pb, stream);
    if (rc) return rc;
    return orag_pairwise_pairs(d_emb, d_workspace, m, dim, d_doc_idx, threshold, cap, d_out_i, d_out_j, d_out_sim,
                               d_out_count, (uint8_t *)d_workspace + pb, workspace_bytes - pb, stream);
}
