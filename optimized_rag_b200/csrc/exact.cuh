// exact.cuh -- the float64 row scorer shared by cosine_exact.cu and pairwise.cu, and the launchers
// other translation units call.
#pragma once
#include "common.cuh"

namespace orag {

// ------------------------------------------------------------------------------------------------
// lane l scores row `rowptr` (nullptr = inactive lane) against QB queries.
// stage: warp-private shared memory, 32 x 33 floats.
template <int QB>
__device__ __forceinline__ void warp_score_rows(const float *rowptr, const float *(&qptr)[QB], int dim,
                                                float *stage, NeuSum (&dot)[QB], NeuSum &sq, bool want_sq)
{
    const int lane = threadIdx.x & 31;
    unsigned long long myp = (unsigned long long)(uintptr_t)rowptr;
#pragma unroll
    for (int b = 0; b < QB; ++b) dot[b].init();
    sq.init();
    // software pipeline: the 32x32 block of chunk c+1 is in flight (registers) while chunk c is summed
    float nxt[32], qnxt[QB];
    auto fetch = [&](int c0) {
        const int col = c0 + lane;
        const bool col_ok = col < dim;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float *p = (const float *)(uintptr_t)__shfl_sync(0xffffffffu, myp, i);
            nxt[i] = (p != nullptr && col_ok) ? __ldg(p + col) : 0.f;
        }
#pragma unroll
        for (int b = 0; b < QB; ++b) qnxt[b] = (qptr[b] != nullptr && col_ok) ? __ldg(qptr[b] + col) : 0.f;
    };
    fetch(0);
    for (int c0 = 0; c0 < dim; c0 += 32) {
        float qreg[QB];
#pragma unroll
        for (int i = 0; i < 32; ++i) stage[i * 33 + lane] = nxt[i];
#pragma unroll
        for (int b = 0; b < QB; ++b) qreg[b] = qnxt[b];
        __syncwarp();
        if (c0 + 32 < dim) fetch(c0 + 32);
        const int jmax = min(32, dim - c0);
        if (jmax == 32) {
            // products of 8 elements are formed ahead of the (inherently sequential) compensated sums so that
            // shared-memory reads, shuffles and multiplies overlap the dependent add chains
#pragma unroll
            for (int j0 = 0; j0 < 32; j0 += 8) {
                double pa[8], pq[QB][8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const double a = (double)stage[lane * 33 + j0 + u];
                    pa[u] = __dmul_rn(a, a);
#pragma unroll
                    for (int b = 0; b < QB; ++b)
                        pq[b][u] = __dmul_rn((double)__shfl_sync(0xffffffffu, qreg[b], j0 + u), a);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (want_sq) sq.add(pa[u]);
#pragma unroll
                    for (int b = 0; b < QB; ++b) dot[b].add(pq[b][u]);
                }
            }
        } else {
            for (int j = 0; j < jmax; ++j) {
                double a = (double)stage[lane * 33 + j];
                if (want_sq) sq.add(__dmul_rn(a, a));
#pragma unroll
                for (int b = 0; b < QB; ++b) {
                    double qv = (double)__shfl_sync(0xffffffffu, qreg[b], j);
                    dot[b].add(__dmul_rn(qv, a));
                }
            }
        }
        __syncwarp();
    }
}


int launch_select_topk(const double *scores, const int64_t *ids, const uint32_t *counts, int64_t n, int64_t ld,
                       int n_queries, int k, int64_t id_base, int normalize, const double *ext_max, int n_ext,
                       int64_t *out_ids, double *out_scores, double *out_max, int32_t *status, cudaStream_t st);
int launch_query_sq(const float *queries, int n_queries, int dim, double *sq, cudaStream_t st);
int launch_cosine_dense(const float *corpus, int64_t n_rows, int dim, const float *queries, int n_queries,
                        const double *sq_q, double *out, cudaStream_t st, double *raw_row_sq = nullptr);
int launch_prefilter(const float *corpus, int dim, const float *queries, const float *inv_qnorm, const int32_t *cand,
                     const uint32_t *cnt, int cap, int n_queries, int k, float *cos32, int cap2, int32_t *surv,
                     uint32_t *surv_cnt, int32_t *status, cudaStream_t st, const uint32_t *thr_key = nullptr,
                     const float *qnorm_scan = nullptr, float margin = 0.f);
int launch_rescore(const float *corpus, int dim, int64_t row_id_base, const float *queries, const double *sq_q,
                   const int32_t *cand, const uint32_t *cnt, int cap, int n_queries, const double *row_sq,
                   double *out_scores, int64_t *out_ids, cudaStream_t st);

}  // namespace orag
