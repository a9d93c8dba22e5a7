// cosine_exact.cu -- the reference's float64 cosine on CUDA cores, plus exact selection.
//
//   * warp_score_rows: lane <-> corpus row, rows staged through shared memory with coalesced
//     128-byte loads, every (row, query) pair accumulated sequentially in the exact order and
//     arithmetic of rag/retrieval.py:362-371 (Neumaier sums, see common.cuh) -> BIT-EXACT scores.
//   * cosine_dense_kernel: all rows x all queries (anchor, small-N path, overflow fallback)
//   * rescore_kernel: the candidate rows emitted by the tensor-core first pass (cosine_tc.cu)
//   * select_topk_kernel: generic exact top-k by (score desc, id asc), optional max-normalisation
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "select.cuh"
#include "exact.cuh"

namespace orag {

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) row_inv_norms_kernel(const float *__restrict__ corpus, int64_t n_rows, int dim,
                                                           float *__restrict__ inv_norm)
{
    int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        const float4 *p = reinterpret_cast<const float4 *>(corpus + r * (int64_t)dim);
        float acc = 0.f;
        int n4 = dim >> 2;
        for (int j = lane; j < n4; j += 32) {
            float4 v = __ldg(p + j);
            acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) inv_norm[r] = acc > 0.f ? rsqrtf(acc) : 0.f;
    }
}

__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float4 *__restrict__ src, uint2 *__restrict__ dst,
                                                         int64_t n4)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n4; i += stride) {
        float4 v = __ldg(src + i);
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
        __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t *>(&lo);
        o.y = *reinterpret_cast<uint32_t *>(&hi);
        dst[i] = o;
    }
}

// fp32 rows -> IEEE fp16 rows scaled by a per-row power of two (warp per row).  s = 2^e is chosen so that the
// largest |x| lands in [2^14, 2^15): x * s is exact, the only rounding is the final fp16 RN (relative 2^-11 for
// normal results; results below 2^-14 are >= 2^28 times smaller than the row maximum and their absolute error
// 2^-25 is irrelevant).  inv_norm_scaled[r] = 1 / (|x| * s) so that (fp16 dot) * inv_norm_scaled is in the same
// units as the fp32 path; scale_out[r] = s.
__global__ void __launch_bounds__(256) f32_to_f16_rows_kernel(const float *__restrict__ src, int64_t n_rows, int dim,
                                                             __half *__restrict__ dst, float *__restrict__ inv_norm_scaled,
                                                             float *__restrict__ scale_out)
{
    const int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int n4 = dim >> 2;
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        const float4 *p = reinterpret_cast<const float4 *>(src + r * (int64_t)dim);
        float mx = 0.f, sq = 0.f;
        for (int j = lane; j < n4; j += 32) {
            const float4 v = __ldg(p + j);
            mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
            sq += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            sq += __shfl_xor_sync(0xffffffffu, sq, o);
        }
        int e = 0;
        if (mx > 0.f && isfinite(mx)) {
            int ex;
            frexpf(mx, &ex);        // mx = m * 2^ex, m in [0.5, 1)  ->  mx * 2^(15 - ex) in [2^14, 2^15)
            e = 15 - ex;
            e = e > 100 ? 100 : (e < -100 ? -100 : e);
        }
        const float s = ldexpf(1.f, e);
        uint2 *q = reinterpret_cast<uint2 *>(dst + r * (int64_t)dim);
        for (int j = lane; j < n4; j += 32) {
            const float4 v = __ldg(p + j);
            const __half2 lo = __floats2half2_rn(v.x * s, v.y * s);
            const __half2 hi = __floats2half2_rn(v.z * s, v.w * s);
            uint2 o;
            o.x = *reinterpret_cast<const uint32_t *>(&lo);
            o.y = *reinterpret_cast<const uint32_t *>(&hi);
            q[j] = o;
        }
        if (lane == 0) {
            if (inv_norm_scaled) inv_norm_scaled[r] = sq > 0.f ? rsqrtf(sq) * ldexpf(1.f, -e) : 0.f;
            if (scale_out) scale_out[r] = s;
        }
    }
}

// sum(q*q) per query in the reference's order (sequential Neumaier per query): lane <-> query, the
// query rows staged through shared memory with coalesced loads (same scheme as the row scorer).
__global__ void __launch_bounds__(256) query_sq_kernel(const float *__restrict__ queries, int n_queries, int dim,
                                                      double *__restrict__ sq)
{
    __shared__ float stage_all[8][32 * 33];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int q = (blockIdx.x * (blockDim.x >> 5) + wib) * 32 + lane;
    const float *rowptr = q < n_queries ? queries + (int64_t)q * dim : nullptr;
    const float *qp[1] = {nullptr};
    NeuSum dot[1], s;
    warp_score_rows<1>(rowptr, qp, dim, stage_all[wib], dot, s, true);
    if (q < n_queries) sq[q] = s.result();
}

constexpr int kDenseQB = 4;

// out[q * n_rows + r] = cosine(query q, row r)
// raw_row_sq != nullptr: out receives the plain dot products and raw_row_sq[r] the row's sum of squares (callers that
// finish the cosine themselves, e.g. with `** 0.5` instead of sqrt: rag/nodes/helpers.py:266-290)
__global__ void __launch_bounds__(256) cosine_dense_kernel(const float *__restrict__ corpus, int64_t n_rows, int dim,
                                                          const float *__restrict__ queries, int n_queries,
                                                          const double *__restrict__ sq_q, double *__restrict__ out,
                                                          double *__restrict__ raw_row_sq)
{
    __shared__ float stage_all[8][32 * 33];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float *stage = stage_all[wib];
    int64_t warp = blockIdx.x * (int64_t)(blockDim.x >> 5) + wib;
    int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    int64_t n_blocks = (n_rows + 31) / 32;
    for (int64_t rb = warp; rb < n_blocks; rb += n_warps) {
        int64_t r = rb * 32 + lane;
        const float *rowptr = r < n_rows ? corpus + r * (int64_t)dim : nullptr;
        double sq_r = 0.0;
        for (int q0 = 0; q0 < n_queries; q0 += kDenseQB) {
            const float *qp[kDenseQB];
#pragma unroll
            for (int b = 0; b < kDenseQB; ++b) qp[b] = (q0 + b < n_queries) ? queries + (int64_t)(q0 + b) * dim : nullptr;
            NeuSum dot[kDenseQB], sq;
            warp_score_rows<kDenseQB>(rowptr, qp, dim, stage, dot, sq, q0 == 0);
            if (q0 == 0) {
                sq_r = sq.result();
                if (raw_row_sq && r < n_rows) raw_row_sq[r] = sq_r;
            }
            if (r < n_rows) {
#pragma unroll
                for (int b = 0; b < kDenseQB; ++b)
                    if (q0 + b < n_queries)
                        out[(int64_t)(q0 + b) * n_rows + r] =
                            raw_row_sq ? dot[b].result() : cosine_from_sums(dot[b].result(), sq_q[q0 + b], sq_r);
            }
        }
    }
}

// Re-score candidate rows: cand[q * cap + slot] (local row) -> scores / global ids.
// One warp per CTA, one (query, 32-slot chunk) task per warp: a task is one long dependent chain per lane (1536
// compensated float64 adds), so the few hundred non-empty tasks of a batch must be spread over ALL SMs with one
// warp per scheduler, not packed eight to a CTA.  Tasks past a query's count exit at once.
__global__ void __launch_bounds__(32) rescore_kernel(const float *__restrict__ corpus, int dim, int64_t row_id_base,
                                                    const float *__restrict__ queries, const double *__restrict__ sq_q,
                                                    const int32_t *__restrict__ cand, const uint32_t *__restrict__ cnt,
                                                    int cap, int n_queries, const double *__restrict__ row_sq,
                                                    double *__restrict__ out_scores, int64_t *__restrict__ out_ids)
{
    __shared__ float stage[32 * 33];
    const int lane = threadIdx.x;
    const int chunks_per_q = (cap + 31) / 32;
    for (int64_t w = blockIdx.x; w < (int64_t)n_queries * chunks_per_q; w += gridDim.x) {
        // chunk-major task order: the (few) non-empty chunks of all queries are the first n_queries tasks
        int q = (int)(w % n_queries);
        int chunk = (int)(w / n_queries);
        uint32_t n = min(cnt[q], (uint32_t)cap);
        if ((uint32_t)chunk * 32u >= n) continue;
        uint32_t slot = chunk * 32 + lane;
        int32_t row = slot < n ? cand[(int64_t)q * cap + slot] : -1;
        const float *rowptr = row >= 0 ? corpus + (int64_t)row * dim : nullptr;
        const float *qp[1] = {queries + (int64_t)q * dim};
        NeuSum dot[1], sq;
        // sum(x*x) of a row does not depend on the query: with the ingest-time table only the dot chain remains
        warp_score_rows<1>(rowptr, qp, dim, stage, dot, sq, row_sq == nullptr);
        if (row >= 0) {
            const double sq_r = row_sq ? row_sq[row] : sq.result();
            out_scores[(int64_t)q * cap + slot] = cosine_from_sums(dot[0].result(), sq_q[q], sq_r);
            out_ids[(int64_t)q * cap + slot] = row_id_base + row;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Generic exact top-k.  Entry i of query q: score = scores[q*ld + i]; id = ids ? ids[q*ld + i] : id_base + i
// (entries with id < 0 are skipped).  n_q = counts ? min(counts[q], cap_n) : n.
// normalize: divide by max (if > 0, else 1.0) before ranking; ext_max [q, n_ext] supplies extra
// maxima (per-shard BM25 maxima after the all-gather).
__global__ void __launch_bounds__(256) select_topk_kernel(const double *__restrict__ scores, const int64_t *__restrict__ ids,
                                                         const uint32_t *__restrict__ counts, int64_t n, int64_t ld,
                                                         int k, int64_t id_base, int normalize,
                                                         const double *__restrict__ ext_max, int n_ext,
                                                         int64_t *__restrict__ out_ids, double *__restrict__ out_scores,
                                                         double *__restrict__ out_max, int32_t *__restrict__ status)
{
    __shared__ Pick scratch[32];
    __shared__ double dscratch[32];
    const int q = blockIdx.x;
    const double *s = scores + (int64_t)q * ld;
    const int64_t *idp = ids ? ids + (int64_t)q * ld : nullptr;
    int64_t nq = n;
    if (counts) {
        uint32_t c = counts[q];
        if ((int64_t)c > n) {
            if (status && threadIdx.x == 0) status[q] |= ORAG_STATUS_OVERFLOW;
            c = (uint32_t)n;
        }
        nq = c;
    }
    // normalize: 0 = rank raw scores; 1 = divide by max (if > 0 else 1.0) and report that divisor;
    //            2 = rank raw scores but report max(raw max, 0) (per-shard maximum for a later merge)
    double m = 1.0;
    if (normalize) {
        double mx = -INFINITY;
        for (int64_t i = threadIdx.x; i < nq; i += blockDim.x)
            if (!idp || idp[i] >= 0) mx = fmax(mx, s[i]);
        for (int e = threadIdx.x; e < n_ext; e += blockDim.x) mx = fmax(mx, ext_max[(int64_t)q * n_ext + e]);
        mx = block_max(mx, dscratch);
        if (mx > 0.0) m = mx;
        if (out_max && threadIdx.x == 0) out_max[q] = (normalize == 2) ? (mx > 0.0 ? mx : 0.0) : m;
        if (normalize == 2) normalize = 0;
    }
    double prev_s = INFINITY;
    int64_t prev_id = -1;
    for (int r = 0; r < k; ++r) {
        Pick best;
        best.valid = 0; best.s = 0.0; best.id = 0;
        for (int64_t i = threadIdx.x; i < nq; i += blockDim.x) {
            int64_t id = idp ? idp[i] : id_base + i;
            if (id < 0) continue;
            double v = normalize ? __ddiv_rn(s[i], m) : s[i];
            // strictly after the previous pick
            if (r > 0 && !ranks_before(prev_s, prev_id, v, id)) continue;
            Pick c;
            c.s = v; c.id = id; c.valid = 1;
            best = better(best, c);
        }
        best = block_best(best, scratch);
        if (threadIdx.x == 0) {
            out_ids[(int64_t)q * k + r] = best.valid ? best.id : -1;
            out_scores[(int64_t)q * k + r] = best.valid ? best.s : 0.0;
        }
        if (!best.valid) {
            for (int r2 = r + 1 + threadIdx.x; r2 < k; r2 += blockDim.x) {
                out_ids[(int64_t)q * k + r2] = -1;
                out_scores[(int64_t)q * k + r2] = 0.0;
            }
            break;
        }
        prev_s = best.s;
        prev_id = best.id;
    }
}

int launch_select_topk(const double *scores, const int64_t *ids, const uint32_t *counts, int64_t n, int64_t ld,
                       int n_queries, int k, int64_t id_base, int normalize, const double *ext_max, int n_ext,
                       int64_t *out_ids, double *out_scores, double *out_max, int32_t *status, cudaStream_t st)
{
    if (n_queries <= 0 || k <= 0) return ORAG_OK;
    select_topk_kernel<<<n_queries, 256, 0, st>>>(scores, ids, counts, n, ld, k, id_base, normalize, ext_max, n_ext,
                                                   out_ids, out_scores, out_max, status);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}

int launch_query_sq(const float *queries, int n_queries, int dim, double *sq, cudaStream_t st)
{
    // a query batch is a handful of warps, each one long dependent chain: one warp per CTA spreads them over the
    // SMs; the ingest-time sweep over millions of rows (orag_row_sq) keeps eight warps per CTA
    if (n_queries <= 4096)
        query_sq_kernel<<<(unsigned)((n_queries + 31) / 32), 32, 0, st>>>(queries, n_queries, dim, sq);
    else
        query_sq_kernel<<<(unsigned)((n_queries + 255) / 256), 256, 0, st>>>(queries, n_queries, dim, sq);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}

int launch_cosine_dense(const float *corpus, int64_t n_rows, int dim, const float *queries, int n_queries,
                        const double *sq_q, double *out, cudaStream_t st, double *raw_row_sq)
{
    int64_t blocks = ((n_rows + 31) / 32 + 7) / 8;
    int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    cosine_dense_kernel<<<(unsigned)blocks, 256, 0, st>>>(corpus, n_rows, dim, queries, n_queries, sq_q, out, raw_row_sq);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}

int launch_rescore(const float *corpus, int dim, int64_t row_id_base, const float *queries, const double *sq_q,
                   const int32_t *cand, const uint32_t *cnt, int cap, int n_queries, const double *row_sq,
                   double *out_scores, int64_t *out_ids, cudaStream_t st)
{
    int64_t blocks = (int64_t)n_queries * ((cap + 31) / 32);
    int64_t lim = (int64_t)sm_count() * 32;
    if (blocks > lim) blocks = lim;
    if (blocks < 1) blocks = 1;
    rescore_kernel<<<(unsigned)blocks, 32, 0, st>>>(corpus, dim, row_id_base, queries, sq_q, cand, cnt, cap,
                                                     n_queries, row_sq, out_scores, out_ids);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}

// ------------------------------------------------------------------------------------------------
// fp32 re-score of the first-pass candidates (the "fp32 re-score of the candidate set" stage).
// One warp per candidate row: coalesced 16-byte loads, fp32 FMA dot product and row sum of squares,
// shuffle reduction.  Per lane 48 sequential FMAs (dim 1536) + 5 shuffle adds: the result is within
// ~53 * 2^-24 * |a||b| of the exact dot product, i.e. the cosine is within eps32(dim) (1e-5 at
// dim 1536: a 2x cushion that also covers the norm, rsqrt and the final multiplies) of the float64 value.
// Task t = slot * n_queries + q, so that the work of all queries interleaves evenly over the warps.
static float eps32(int dim) { return fmaxf(1e-5f, (float)(dim / 32 + 16) * 1.2e-7f); }

__global__ void __launch_bounds__(256) prefilter_kernel(const float *__restrict__ corpus, int dim,
                                                       const float *__restrict__ queries,
                                                       const float *__restrict__ inv_qnorm,
                                                       const int32_t *__restrict__ cand, const uint32_t *__restrict__ cnt,
                                                       int cap, int n_queries, float *__restrict__ out_cos,
                                                       const uint32_t *__restrict__ thr_key,
                                                       const float *__restrict__ qnorm_scan, float margin)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    uint32_t mx = 0;
    for (int q = lane; q < n_queries; q += 32) mx = max(mx, min(cnt[q], (uint32_t)cap));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const int64_t n_tasks = (int64_t)mx * n_queries;
    const int n4 = dim >> 2;
    for (int64_t t = warp; t < n_tasks; t += n_warps) {
        const int q = (int)(t % n_queries);
        const uint32_t slot = (uint32_t)(t / n_queries);
        if (slot >= min(cnt[q], (uint32_t)cap)) continue;
        if (thr_key) {
            // out_cos arrives holding the first-pass value the scan emitted this candidate with.  The threshold only ever
            // rose during the scan and every value it took admits all true top-k rows, so its FINAL value does too: a
            // candidate emitted early, under a looser threshold, that the final one would reject needs no re-score
            // (about four in five at k = 10; same expression as the scan's own test).
            const float tv = (ordered_to_float(thr_key[q]) - margin) * qnorm_scan[q];
            if (out_cos[(int64_t)q * cap + slot] < tv) {
                if (lane == 0) out_cos[(int64_t)q * cap + slot] = -INFINITY;
                continue;
            }
        }
        const int32_t row = cand[(int64_t)q * cap + slot];
        const float4 *rp = reinterpret_cast<const float4 *>(corpus + (int64_t)row * dim);
        const float4 *qp = reinterpret_cast<const float4 *>(queries + (int64_t)q * dim);
        float dot = 0.f, sq = 0.f;
        // the row is read once from HBM with no reuse: keep eight 16-byte loads per lane in flight
        for (int j0 = lane; j0 < n4; j0 += 32 * 4) {
            float4 a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + 32 * u;
                a[u] = j < n4 ? __ldg(rp + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                b[u] = j < n4 ? __ldg(qp + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                dot = fmaf(a[u].x, b[u].x, dot); dot = fmaf(a[u].y, b[u].y, dot);
                dot = fmaf(a[u].z, b[u].z, dot); dot = fmaf(a[u].w, b[u].w, dot);
                sq = fmaf(a[u].x, a[u].x, sq); sq = fmaf(a[u].y, a[u].y, sq);
                sq = fmaf(a[u].z, a[u].z, sq); sq = fmaf(a[u].w, a[u].w, sq);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            dot += __shfl_xor_sync(0xffffffffu, dot, o);
            sq += __shfl_xor_sync(0xffffffffu, sq, o);
        }
        if (lane == 0) out_cos[(int64_t)q * cap + slot] = sq > 0.f ? dot * rsqrtf(sq) * inv_qnorm[q] : 0.f;
    }
}

// One CTA per query: k-th best fp32 cosine among the candidates -> keep everything within 2*kEps32
// of it (every row whose float64 cosine can reach the true top-k), compacted into surv[q][*].
// Proof sketch: k rows have cos32 >= kth32, hence cos64 >= kth32 - eps, so kth64 >= kth32 - eps; a true
// top-k row r has cos32(r) >= cos64(r) - eps >= kth64 - eps >= kth32 - 2 eps.
__global__ void __launch_bounds__(256) prune_kernel(const float *__restrict__ cos32, const int32_t *__restrict__ cand,
                                                   const uint32_t *__restrict__ cnt, int cap, int k, int cap2,
                                                   float eps, int32_t *__restrict__ surv,
                                                   uint32_t *__restrict__ surv_cnt, int32_t *__restrict__ status)
{
    __shared__ Pick scratch[32];
    __shared__ uint32_t s_n;
    const int q = blockIdx.x;
    uint32_t n = cnt[q];
    if (n > (uint32_t)cap) {
        if (status && threadIdx.x == 0) status[q] |= ORAG_STATUS_OVERFLOW;
        n = cap;
    }
    const float *c = cos32 + (int64_t)q * cap;
    const int32_t *rows = cand + (int64_t)q * cap;
    if (threadIdx.x == 0) s_n = 0;
    float thr = -INFINITY;
    if (n > (uint32_t)k) {
        // k rounds of block arg-max under (value desc, slot asc); after round k-1 `prev` is the k-th best
        double prev_s = INFINITY;
        int64_t prev_id = -1;
        for (int r = 0; r < k; ++r) {
            Pick best;
            best.valid = 0; best.s = 0.0; best.id = 0;
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
                const double v = (double)c[i];
                if (r > 0 && !ranks_before(prev_s, prev_id, v, (int64_t)i)) continue;
                Pick p;
                p.s = v; p.id = i; p.valid = 1;
                best = better(best, p);
            }
            best = block_best(best, scratch);
            prev_s = best.s;
            prev_id = best.id;
        }
        thr = (float)prev_s - 2.f * eps;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        if (c[i] >= thr) {
            const uint32_t slot = atomicAdd(&s_n, 1u);
            if (slot < (uint32_t)cap2) surv[(int64_t)q * cap2 + slot] = rows[i];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        surv_cnt[q] = s_n;  // > cap2 is flagged as overflow by the final selection
    }
}

int launch_prefilter(const float *corpus, int dim, const float *queries, const float *inv_qnorm, const int32_t *cand,
                     const uint32_t *cnt, int cap, int n_queries, int k, float *cos32, int cap2, int32_t *surv,
                     uint32_t *surv_cnt, int32_t *status, cudaStream_t st, const uint32_t *thr_key, const float *qnorm_scan,
                     float margin)
{
    prefilter_kernel<<<sm_count() * 8, 256, 0, st>>>(corpus, dim, queries, inv_qnorm, cand, cnt, cap, n_queries, cos32,
                                                     thr_key, qnorm_scan, margin);
    ORAG_LAUNCH_CHECK();
    prune_kernel<<<n_queries, 256, 0, st>>>(cos32, cand, cnt, cap, k, cap2, eps32(dim), surv, surv_cnt, status);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}

}  // namespace orag

// ================================================================================================
extern "C" int orag_row_inv_norms(const float *d_corpus, int64_t n_rows, int dim, float *d_inv_norm, void *stream)
{
    ORAG_REQUIRE(d_corpus && d_inv_norm && n_rows >= 0 && dim > 0 && dim % 4 == 0, "row_inv_norms");
    ORAG_REQUIRE((reinterpret_cast<uintptr_t>(d_corpus) & 15) == 0, "corpus must be 16-byte aligned");
    if (n_rows == 0) return ORAG_OK;
    int64_t blocks = (n_rows + 7) / 8;
    int64_t cap = (int64_t)orag::sm_count() * 16;
    if (blocks > cap) blocks = cap;
    orag::row_inv_norms_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_corpus, n_rows, dim, d_inv_norm);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}

extern "C" int orag_f32_to_bf16(const float *d_src, void *d_dst, int64_t count, void *stream)
{
    ORAG_REQUIRE(d_src && d_dst && count >= 0 && count % 4 == 0, "f32_to_bf16: count % 4 == 0");
    ORAG_REQUIRE((reinterpret_cast<uintptr_t>(d_src) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_dst) & 7) == 0,
                 "alignment");
    if (count == 0) return ORAG_OK;
    int64_t n4 = count / 4;
    int64_t blocks = (n4 + 255) / 256;
    int64_t cap = (int64_t)orag::sm_count() * 32;
    if (blocks > cap) blocks = cap;
    orag::f32_to_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(d_src), reinterpret_cast<uint2 *>(d_dst), n4);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}

extern "C" int orag_row_sq(const float *d_corpus, int64_t n_rows, int dim, double *d_row_sq, void *stream)
{
    ORAG_REQUIRE(d_corpus && d_row_sq && n_rows >= 0 && dim > 0, "row_sq");
    if (n_rows == 0) return ORAG_OK;
    // the reference's sum(a * a for a in row): sequential Neumaier sum per row (lane <-> row), 2^31 rows at most
    ORAG_REQUIRE(n_rows < ((int64_t)1 << 31), "row_sq: n_rows < 2^31");
    const int chunk = 1 << 22;
    for (int64_t r0 = 0; r0 < n_rows; r0 += chunk) {
        const int n = (int)(n_rows - r0 < chunk ? n_rows - r0 : chunk);
        int rc = orag::launch_query_sq(d_corpus + r0 * dim, n, dim, d_row_sq + r0, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return ORAG_OK;
}

extern "C" int orag_f32_to_f16_rows(const float *d_src, int64_t n_rows, int dim, void *d_dst_f16,
                                    float *d_inv_norm_scaled, float *d_scale, void *stream)
{
    ORAG_REQUIRE(d_src && d_dst_f16 && n_rows >= 0 && dim > 0 && dim % 4 == 0, "f32_to_f16_rows: dim % 4 == 0");
    ORAG_REQUIRE((reinterpret_cast<uintptr_t>(d_src) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_dst_f16) & 7) == 0,
                 "alignment");
    if (n_rows == 0) return ORAG_OK;
    int64_t blocks = (n_rows + 7) / 8;
    int64_t cap = (int64_t)orag::sm_count() * 16;
    if (blocks > cap) blocks = cap;
    orag::f32_to_f16_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        d_src, n_rows, dim, reinterpret_cast<__half *>(d_dst_f16), d_inv_norm_scaled, d_scale);
    ORAG_LAUNCH_CHECK();
    return ORAG_OK;
}

extern "C" int orag_dense_topk(const double *d_scores, int64_t n, int64_t ld, int n_queries, int k, int64_t id_base,
                               int normalize, int64_t *d_out_ids, double *d_out_scores, double *d_out_max, void *stream)
{
    ORAG_REQUIRE(d_scores && d_out_ids && d_out_scores && n >= 0 && ld >= n && k > 0 && n_queries >= 0, "dense_topk");
    return orag::launch_select_topk(d_scores, nullptr, nullptr, n, ld, n_queries, k, id_base, normalize, nullptr, 0,
                                    d_out_ids, d_out_scores, d_out_max, nullptr, (cudaStream_t)stream);
}

extern "C" int orag_topk_merge(const int64_t *d_cand_ids, const double *d_cand_scores, int m, int n_queries, int k,
                               const double *d_shard_max, int n_shards, int64_t *d_out_ids, double *d_out_scores,
                               double *d_out_max, void *stream)
{
    ORAG_REQUIRE(d_cand_ids && d_cand_scores && d_out_ids && d_out_scores && m >= 0 && k > 0, "topk_merge");
    int normalize = d_shard_max != nullptr;
    return orag::launch_select_topk(d_cand_scores, d_cand_ids, nullptr, m, m, n_queries, k, 0, normalize, d_shard_max,
                                    normalize ? n_shards : 0, d_out_ids, d_out_scores, d_out_max, nullptr,
                                    (cudaStream_t)stream);
}

extern "C" int orag_dot_dense(const float *d_corpus, int64_t n_rows, int dim, const float *d_queries, int n_queries,
                              double *d_out_dots, double *d_out_row_sq, double *d_out_query_sq, void *stream)
{
    ORAG_REQUIRE(d_corpus && d_queries && d_out_dots && d_out_row_sq && d_out_query_sq && n_rows >= 0 && dim > 0 &&
                     n_queries > 0,
                 "dot_dense");
    int rc = orag::launch_query_sq(d_queries, n_queries, dim, d_out_query_sq, (cudaStream_t)stream);
    if (rc) return rc;
    if (n_rows == 0) return ORAG_OK;
    return orag::launch_cosine_dense(d_corpus, n_rows, dim, d_queries, n_queries, d_out_query_sq, d_out_dots,
                                     (cudaStream_t)stream, d_out_row_sq);
}

extern "C" int orag_cosine_dense(const float *d_corpus, int64_t n_rows, int dim, const float *d_queries, int n_queries,
                                 double *d_out, void *stream)
{
    ORAG_REQUIRE(d_corpus && d_queries && d_out && n_rows >= 0 && dim > 0 && n_queries > 0, "cosine_dense");
    // sq_q lives at the tail of d_out?  No: keep the ABI allocation-free by computing it into the
    // first n_queries doubles of a row that is overwritten afterwards is not possible either, so the
    // caller-visible contract is: d_out holds n_queries * n_rows + n_queries doubles.
    double *sq = d_out + (int64_t)n_queries * n_rows;
    int rc = orag::launch_query_sq(d_queries, n_queries, dim, sq, (cudaStream_t)stream);
    if (rc) return rc;
    if (n_rows == 0) return ORAG_OK;
    return orag::launch_cosine_dense(d_corpus, n_rows, dim, d_queries, n_queries, sq, d_out, (cudaStream_t)stream);
}
